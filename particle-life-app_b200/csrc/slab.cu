// Multi-GPU slab decomposition of the physics step: one process per GPU, rank g owns
// grid rows [g*ny/G, (g+1)*ny/G).  The sorted order is row-major by cell, so a slab is one contiguous
// index range and its first / last rows are contiguous sub-ranges: packing a halo is a plain copy.
//
// Per step (host drives the three phases and exchanges the messages between them, e.g. with
// torch.distributed / NCCL send-recv; see plife/slab.py):
//   phase SORT    cell-list build of the owned particles; pack first and last owned row into the halo
//                 messages {count | per-cell END offsets of the row | 16-byte particle records}
//   (exchange)    halo_send[0] -> down neighbour's halo_recv[1], halo_send[1] -> up neighbour's halo_recv[0]
//   phase FORCE   place the ghost rows around the owned block of the sorted array
//                 [ghost below | owned | ghost above], run the force/integrate kernel over the owned
//                 particles; its epilogue bins the new positions and appends particles whose new row
//                 belongs to a neighbour to the migration messages
//   (exchange)    mig_send[0] -> down neighbour's mig_recv[1], mig_send[1] -> up neighbour's mig_recv[0]
//   phase FINISH  append the arrivals (binning them), update the particle count
//
// Two ways to move the messages:
//   external  the host exchanges caller-provided buffers between the phases (torch.distributed/NCCL);
//   peer      (default) the library owns the buffers, ranks map each other's receive slots with CUDA IPC
//             and the kernels push a message straight into the neighbour's slot over NVLink, then raise a
//             flag (sequence number) there; the consumer's stream spins on its own flag in a 1-thread
//             kernel.  Slots are double-buffered by step parity: a rank can only be one message ahead of
//             its neighbour (it waits for the neighbour's flag of step k before it produces step k+1), so
//             the slot of step k+1 was consumed by the neighbour before its step-k flag was raised.
//
// Order: arrivals from below logically precede all residents and arrivals from above follow them
// (StableKey in cells.cu), so the concatenation of the slabs in rank order is exactly the
// single-GPU particle order.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "plife_internal.h"

using namespace plife;

namespace plife {
int slab_make_grid(plife_handle *h, Grid *g);
int slab_sort(plife_handle *h, const Grid &g);
int slab_fail(plife_handle *h, int code, const char *msg);
cudaError_t slab_force(plife_handle *h, const Grid &g, double dt);
int slab_reset_capacity(plife_handle *h);
} // namespace plife

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int offsets_records(int nx) { return (nx + 3) >> 2; }

// wall-clock nanoseconds (not SM cycles: the spin limits below are times, whatever the clock does)
__device__ __forceinline__ unsigned long long wall_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// dir 0: first owned row (local row 1) -> becomes the down neighbour's top ghost row
// dir 1: last owned row (local row nly-2) -> becomes the up neighbour's bottom ghost row
// Completion signal of a multi-CTA producer: every CTA fences its (remote) writes and takes a ticket; the CTA that draws
// the last one publishes `seq` in the consumer's flag and resets the ticket for the next step.
__device__ __forceinline__ void signal_when_all_done(unsigned int *ticket, volatile unsigned long long *flag, unsigned long long seq,
                                                     unsigned int nblocks)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == nblocks - 1) {
            *ticket = 0;
            __threadfence_system();
            *flag = seq;
        }
    }
}

// `msg0/msg1` are where the two messages are written: the local send buffers (external exchange), or - peer exchange -
// straight into the neighbours' receive slots over NVLink, followed by the flag (flag0/flag1 non-NULL).
__global__ void __launch_bounds__(kThreads) pack_halo(const float4 *__restrict__ pt_sorted, const int32_t *__restrict__ cell_end,
                                                      Grid g, int halo_cap, float4 *__restrict__ msg0, float4 *__restrict__ msg1,
                                                      float4 *__restrict__ mig0, float4 *__restrict__ mig1,
                                                      volatile unsigned long long *flag0, volatile unsigned long long *flag1,
                                                      unsigned int *tickets, unsigned long long seq)
{
    const int dir = blockIdx.y;
    float4 *msg = dir ? msg1 : msg0;
    volatile unsigned long long *flag = dir ? flag1 : flag0;
    if (blockIdx.x == 0 && threadIdx.x == 0) (dir ? mig1 : mig0)[0] = make_float4(0.f, 0.f, 0.f, 0.f); // this step's migration cursor
    const int row = dir ? g.nly - 2 : 1;
    const int start = __ldg(cell_end + row * g.nx - 1);
    const int end = __ldg(cell_end + (row + 1) * g.nx - 1);
    const int count = end - start;
    const int noff = offsets_records(g.nx);
    const int t = blockIdx.x * kThreads + threadIdx.x;
    if (t == 0) {
        int4 hd = make_int4(count <= halo_cap ? count : -count, g.nx, 0, 0);
        msg[0] = *reinterpret_cast<float4 *>(&hd);
    }
    if (count <= halo_cap) { // else overflow: reported by the receiver and by phase FINISH
        int32_t *off = reinterpret_cast<int32_t *>(msg + 1);
        for (int c = t; c < 4 * noff; c += gridDim.x * kThreads) off[c] = c < g.nx ? __ldg(cell_end + row * g.nx + c) - start : 0;
        for (int k = t; k < count; k += gridDim.x * kThreads) msg[1 + noff + k] = __ldg(pt_sorted + start + k);
    }
    if (flag) signal_when_all_done(tickets + dir, flag, seq, gridDim.x);
}

// which 0: ghost row below (local row 0), right-aligned before the owned block at `first`
// which 1: ghost row above (local row nly-1), placed after the owned block
__global__ void __launch_bounds__(kThreads) unpack_halo(float4 *__restrict__ pt_sorted, int32_t *__restrict__ cell_end, Grid g,
                                                        int first, int n, const float4 *msg0, const float4 *msg1, int *__restrict__ err,
                                                        const volatile unsigned long long *flag0, const volatile unsigned long long *flag1,
                                                        unsigned long long seq, unsigned long long spin_ns)
{
    const int which = blockIdx.y;
    const float4 *msg = which ? msg1 : msg0;
    if (!msg) return;
    // peer exchange: the message is complete once the neighbour's flag reaches this step's sequence number; every CTA
    // checks for itself (bounded spin, so a dead peer cannot hang the GPU).  The message was written during this
    // kernel's lifetime, so it is read with ld.cg, not through the read-only path.
    const volatile unsigned long long *flag = which ? flag1 : flag0;
    if (flag) {
        __shared__ int s_timeout;
        if (threadIdx.x == 0) {
            s_timeout = 0;
            const unsigned long long t0 = wall_ns();
            while (*flag < seq) {
                if (wall_ns() - t0 > spin_ns) {
                    s_timeout = 1;
                    break;
                }
                __nanosleep(100);
            }
            __threadfence_system();
        }
        __syncthreads();
        if (s_timeout) {
            if (threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(err, 1 << 16);
            return;
        }
    }
    const int4 hd = __ldcg(reinterpret_cast<const int4 *>(msg));
    const int t = blockIdx.x * kThreads + threadIdx.x;
    if (hd.x < 0 || hd.y != g.nx) {
        if (t == 0) atomicAdd(err, 1);
        return;
    }
    const int count = hd.x;
    const int noff = offsets_records(g.nx);
    const int base = which ? first + n : first - count;
    const int row = which ? g.nly - 1 : 0;
    const int32_t *off = reinterpret_cast<const int32_t *>(msg + 1);
    for (int c = t; c < g.nx; c += gridDim.x * kThreads) cell_end[row * g.nx + c] = base + __ldcg(off + c);
    if (which == 0 && t == 0) cell_end[-1] = base;
    for (int k = t; k < count; k += gridDim.x * kThreads) pt_sorted[base + k] = __ldcg(msg + 1 + noff + k);
}

// Arrivals: message records {x,y,type,id},{vx,vy,source slot,-}.  Each is placed at
// base + (rank of its source slot among the arrivals of the same message): the sender appended them
// with an atomic cursor, the rank restores the sender's array order.
__global__ void __launch_bounds__(kThreads) append_arrivals(const float4 *__restrict__ msg, int k, int base, Grid g,
                                                            float4 *__restrict__ pt, float2 *__restrict__ vel,
                                                            int32_t *__restrict__ cell, int32_t *__restrict__ count,
                                                            int *__restrict__ err)
{
    const int j = blockIdx.x * kThreads + threadIdx.x;
    if (j >= k) return;
    const float4 a = __ldg(msg + 1 + 2 * j), b = __ldg(msg + 2 + 2 * j);
    const int src = __float_as_int(b.z);
    int rank = 0;
    for (int q = 0; q < k; ++q) rank += (__float_as_int(__ldg(msg + 2 + 2 * q).z) < src) ? 1 : 0;
    const int dst = base + rank;
    pt[dst] = a;
    vel[dst] = make_float2(b.x, b.y);
    const int cxy = cell_coords((double)a.x, (double)a.y, g);
    const int c = container_of(cxy, g);
    cell[dst] = c < 0 ? -1 : cxy;
    if (c >= 0) atomicAdd(count + c, 1);
    else atomicAdd(err, 1); // sender and receiver disagree about ownership
}

// copy a message (its used records only) into the neighbour's receive slot
__global__ void __launch_bounds__(kThreads) push_msg(const float4 *__restrict__ local, float4 *__restrict__ peer, int is_halo, int nx,
                                                     int cap, volatile unsigned long long *flag, unsigned int *ticket,
                                                     unsigned long long seq)
{
    const int count = *reinterpret_cast<const int *>(local);
    int nrec;
    if (is_halo) nrec = count < 0 ? 1 : 1 + offsets_records(nx) + count;
    else nrec = 1 + 2 * min(count, cap);
    for (int k = blockIdx.x * kThreads + threadIdx.x; k < nrec; k += gridDim.x * kThreads) peer[k] = local[k];
    if (flag) signal_when_all_done(ticket, flag, seq, gridDim.x);
    else __threadfence_system();
}

// Phase FINISH: wait for the neighbours' migration messages (peer mode), then put the four message headers and the
// error word where the host can read them after one stream synchronisation - mapped pinned memory, no copies.
__global__ void finish_headers(const volatile unsigned long long *f0, const volatile unsigned long long *f1, unsigned long long seq,
                               int *err, const float4 *ms0, const float4 *ms1, const float4 *mi0, const float4 *mi1,
                               volatile int4 *out, unsigned long long spin_ns)
{
    const unsigned long long t0 = wall_ns();
    for (int k = 0; k < 2; ++k) {
        const volatile unsigned long long *f = k ? f1 : f0;
        if (!f) continue;
        while (*f < seq) {
            if (wall_ns() - t0 > spin_ns) {
                atomicAdd(err, 1 << 16);
                break;
            }
            __nanosleep(200);
        }
    }
    __threadfence_system();
    const float4 *src[4] = {ms0, ms1, mi0, mi1};
    for (int k = 0; k < 4; ++k) {
        int4 v = make_int4(0, 0, 0, 0);
        if (src[k]) {
            const volatile int *p = reinterpret_cast<const volatile int *>(src[k]);
            v = make_int4(p[0], p[1], p[2], p[3]);
        }
        out[k].x = v.x; out[k].y = v.y; out[k].z = v.z; out[k].w = v.w;
    }
    out[4].x = *reinterpret_cast<volatile int *>(err);
    __threadfence_system();
}

int fail(plife_handle *h, int code, const char *msg) { return slab_fail(h, code, msg); }

#define CUS(h, expr)                                                      \
    do {                                                                  \
        cudaError_t e_ = (expr);                                          \
        if (e_ != cudaSuccess) {                                          \
            h->poisoned = true;                                           \
            return fail(h, PLIFE_ERR_CUDA, cudaGetErrorString(e_));       \
        }                                                                 \
    } while (0)

} // namespace

// ---- peer-mode buffer layout (one allocation per rank, exported through CUDA IPC) ----
// records (16 B): [0,8) flags: 8 x uint64 (halo from dn, halo from up, mig from dn, mig from up, spare)
//                 then halo slots [parity][dir][hrec], then migration slots [parity][dir][mrec]
namespace {
constexpr int kFlagRecords = 8;
enum { F_HALO_DN = 0, F_HALO_UP = 1, F_MIG_DN = 2, F_MIG_UP = 3, F_TICKETS = 8 }; // u64 slots of the flag block; 8..9 hold four local u32 tickets

inline float4 *halo_slot(float4 *base, const SlabState &S, int parity, int dir)
{
    return base + kFlagRecords + (size_t)(parity * 2 + dir) * S.hrec;
}
inline float4 *mig_slot(float4 *base, const SlabState &S, int parity, int dir)
{
    return base + kFlagRecords + (size_t)4 * S.hrec + (size_t)(parity * 2 + dir) * S.mrec;
}
inline volatile unsigned long long *flag_of(float4 *base, int idx)
{
    return base ? reinterpret_cast<volatile unsigned long long *>(base) + idx : nullptr;
}

void slab_release(plife_handle *h)
{
    SlabState &S = h->slab;
    for (int d = 0; d < 2; d++) {
        if (S.peer_ipc[d] && S.peer_base[d] && !(d == 1 && S.peer_base[1] == S.peer_base[0] && S.peer_ipc[0]))
            cudaIpcCloseMemHandle(S.peer_base[d]);
        S.peer_base[d] = nullptr;
        S.peer_ipc[d] = false;
    }
    if (S.h_hdr) cudaFreeHost((void *)S.h_hdr);
    if (S.peer_mode) {
        cudaFree(S.xbuf);
        for (int d = 0; d < 2; d++) {
            cudaFree(S.halo_send[d]);
            cudaFree(S.mig_send[d]);
        }
    }
    S = SlabState{};
}
} // namespace

namespace plife {
void slab_destroy(plife_handle *h) { slab_release(h); }
} // namespace plife

extern "C" {

int64_t plife_slab_halo_records(int32_t nx, int64_t halo_cap) { return 1 + (nx + 3) / 4 + halo_cap; }
int64_t plife_slab_migrate_records(int64_t mig_cap) { return 1 + 2 * mig_cap; }

int plife_slab_configure(plife_handle *h, int32_t rank, int32_t world, int64_t halo_cap, int64_t mig_cap,
                         const plife_slab_buffers *bufs)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (h->precision != PLIFE_F32) return fail(h, PLIFE_ERR_INVALID, "slab mode is implemented for PLIFE_F32 handles");
    if (h->flags & PLIFE_FLAG_NO_FUSED_BIN) return fail(h, PLIFE_ERR_INVALID, "slab mode needs the fused binning (migration rides on it)");
    if (world < 1 || rank < 0 || rank >= world || halo_cap < 1 || mig_cap < 1) return fail(h, PLIFE_ERR_INVALID, "slab_configure: bad arguments");
    if (h->n > 0) return fail(h, PLIFE_ERR_STATE, "configure the slab before uploading particles");
    if (cudaSetDevice(h->device) != cudaSuccess) return fail(h, PLIFE_ERR_CUDA, "cudaSetDevice");
    slab_release(h);
    SlabState &S = h->slab;
    Grid g;
    S.on = false;
    int rc = slab_make_grid(h, &g);
    if (rc) return rc;
    S.on = true;
    S.rank = rank;
    S.world = world;
    S.halo_cap = halo_cap;
    S.mig_cap = mig_cap;
    S.nx_cfg = g.nx;
    S.hrec = plife_slab_halo_records(g.nx, halo_cap);
    S.mrec = plife_slab_migrate_records(mig_cap);
    S.seq = 1;
    // how long a kernel waits for a neighbour's message before it reports a dead peer (wall time; default 30 s)
    if (const char *ev = getenv("PLIFE_SLAB_TIMEOUT_MS")) {
        const double ms = atof(ev);
        if (ms > 0) S.spin_ns = (unsigned long long)(ms * 1e6);
    }
    if (bufs) { // external exchange: the host moves the messages between the phases
        for (int d = 0; d < 2; d++) {
            if (!bufs->halo_send[d] || !bufs->halo_recv[d] || !bufs->mig_send[d] || !bufs->mig_recv[d]) {
                S.on = false;
                return fail(h, PLIFE_ERR_INVALID, "slab_configure: NULL exchange buffer");
            }
            S.halo_send[d] = (float4 *)bufs->halo_send[d];
            S.halo_recv[d] = (float4 *)bufs->halo_recv[d];
            S.mig_send[d] = (float4 *)bufs->mig_send[d];
            S.mig_recv[d] = (float4 *)bufs->mig_recv[d];
        }
        S.peer_mode = false;
    } else { // peer exchange: library-owned buffers, neighbours connect with plife_slab_connect_*
        S.peer_mode = true;
        S.xrecords = kFlagRecords + 4 * S.hrec + 4 * S.mrec;
        cudaError_t e = cudaMalloc((void **)&S.xbuf, (size_t)S.xrecords * 16);
        for (int d = 0; d < 2 && e == cudaSuccess; d++) {
            e = cudaMalloc((void **)&S.halo_send[d], (size_t)S.hrec * 16);
            if (e == cudaSuccess) e = cudaMalloc((void **)&S.mig_send[d], (size_t)S.mrec * 16);
        }
        if (e == cudaSuccess) e = cudaMemset(S.xbuf, 0, (size_t)S.xrecords * 16);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            slab_release(h);
            return fail(h, PLIFE_ERR_OOM, "slab_configure: allocating exchange buffers failed");
        }
    }
    if (cudaHostAlloc((void **)&S.h_hdr, 5 * sizeof(int4), cudaHostAllocMapped) != cudaSuccess) {
        slab_release(h);
        return fail(h, PLIFE_ERR_OOM, "slab_configure: allocating the pinned header block failed");
    }
    S.phase = PLIFE_SLAB_SORT;
    h->prebinned = false;
    return slab_reset_capacity(h); // the particle buffers need room for the ghost rows
}

int plife_slab_export(plife_handle *h, void *ipc_handle_64_bytes)
{
    if (!h || !ipc_handle_64_bytes) return PLIFE_ERR_INVALID;
    if (!h->slab.on || !h->slab.peer_mode) return fail(h, PLIFE_ERR_STATE, "slab_export: not in peer mode");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (cudaSetDevice(h->device) != cudaSuccess) return fail(h, PLIFE_ERR_CUDA, "cudaSetDevice");
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, h->slab.xbuf);
    if (e != cudaSuccess) return fail(h, PLIFE_ERR_CUDA, cudaGetErrorString(e));
    memcpy(ipc_handle_64_bytes, &hd, 64);
    return PLIFE_OK;
}

// neighbours in other processes: the 64-byte handles they exported (NULL: no neighbour in that direction)
int plife_slab_connect_ipc(plife_handle *h, const void *down_handle, const void *up_handle)
{
    if (!h) return PLIFE_ERR_INVALID;
    SlabState &S = h->slab;
    if (!S.on || !S.peer_mode) return fail(h, PLIFE_ERR_STATE, "slab_connect: not in peer mode");
    if (cudaSetDevice(h->device) != cudaSuccess) return fail(h, PLIFE_ERR_CUDA, "cudaSetDevice");
    const void *hd[2] = {down_handle, up_handle};
    for (int d = 0; d < 2; d++) {
        if (!hd[d]) continue;
        if (d == 1 && hd[0] && memcmp(hd[0], hd[1], 64) == 0) { // two ranks: both neighbours are the same peer
            S.peer_base[1] = S.peer_base[0];
            S.peer_ipc[1] = true;
            continue;
        }
        cudaIpcMemHandle_t m;
        memcpy(&m, hd[d], 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, m, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(h, PLIFE_ERR_CUDA, cudaGetErrorString(e));
        S.peer_base[d] = (float4 *)p;
        S.peer_ipc[d] = true;
    }
    return PLIFE_OK;
}

// neighbours that are handles of this process (virtual ranks on one device, or one process driving several GPUs
// with peer access enabled)
int plife_slab_connect_local(plife_handle *h, plife_handle *down, plife_handle *up)
{
    if (!h) return PLIFE_ERR_INVALID;
    SlabState &S = h->slab;
    if (!S.on || !S.peer_mode) return fail(h, PLIFE_ERR_STATE, "slab_connect: not in peer mode");
    plife_handle *nb[2] = {down, up};
    for (int d = 0; d < 2; d++) {
        if (!nb[d]) continue;
        if (!nb[d]->slab.on || !nb[d]->slab.peer_mode || nb[d]->slab.hrec != S.hrec || nb[d]->slab.mrec != S.mrec)
            return fail(h, PLIFE_ERR_INVALID, "slab_connect_local: neighbour not configured identically");
        S.peer_base[d] = nb[d]->slab.xbuf;
        S.peer_ipc[d] = false;
    }
    return PLIFE_OK;
}

int plife_slab_rows(plife_handle *h, int32_t *row_lo, int32_t *row_hi, int32_t *nx)
{
    if (!h || !h->slab.on) return PLIFE_ERR_STATE;
    Grid g;
    int rc = slab_make_grid(h, &g);
    if (rc) return rc;
    if (row_lo) *row_lo = g.row_lo;
    if (row_hi) *row_hi = g.row_hi;
    if (nx) *nx = g.nx;
    return PLIFE_OK;
}

int plife_slab_phase(plife_handle *h, int32_t phase, double dt)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (h->poisoned) return fail(h, PLIFE_ERR_CUDA, "handle poisoned by an earlier CUDA error");
    if (!h->slab.on) return fail(h, PLIFE_ERR_STATE, "plife_slab_configure has not been called");
    if (phase != h->slab.phase) return fail(h, PLIFE_ERR_STATE, "slab phases must run in order SORT, FORCE, FINISH");
    if (cudaSetDevice(h->device) != cudaSuccess) return fail(h, PLIFE_ERR_CUDA, "cudaSetDevice");
    Grid g;
    int rc = slab_make_grid(h, &g);
    if (rc) return rc;
    SlabState &S = h->slab;
    if (g.nx > S.nx_cfg) return fail(h, PLIFE_ERR_STATE, "slab: rmax shrank since plife_slab_configure (halo messages would not fit): reconfigure");
    const bool wrap = h->settings.wrap != 0;
    const bool has_dn = S.world > 1 && (wrap || S.rank > 0);
    const bool has_up = S.world > 1 && (wrap || S.rank < S.world - 1);
    if (S.peer_mode && ((has_dn && !S.peer_base[0]) || (has_up && !S.peer_base[1])))
        return fail(h, PLIFE_ERR_STATE, "slab: neighbours not connected (plife_slab_connect_ipc / _local)");
    int *d_err = reinterpret_cast<int *>(h->d_scalar + 4);
    const int first = (int)S.halo_cap;
    const int parity = (int)(S.seq & 1);
    float4 *dn = S.peer_base[0], *up = S.peer_base[1];
    // where this step's messages arrive
    float4 *halo_in[2], *mig_in[2];
    for (int d = 0; d < 2; d++) {
        halo_in[d] = S.peer_mode ? halo_slot(S.xbuf, S, parity, d) : S.halo_recv[d];
        mig_in[d] = S.peer_mode ? mig_slot(S.xbuf, S, parity, d) : S.mig_recv[d];
    }

    if (phase == PLIFE_SLAB_SORT) {
        rc = slab_sort(h, g); // bin (if needed), scan, scatter, gather: owned block of the sorted array
        if (rc) return rc;
        const int sorted = h->cur ^ 1;
        dim3 grid(32, 2);
        unsigned int *tickets = reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned long long *>(S.xbuf) + F_TICKETS); // 4 local counters
        if (S.peer_mode) {
            // my first row is the down neighbour's ghost row ABOVE its slab (its slot dir 1), and vice versa; the pack kernel
            // writes it there and raises the neighbour's flag when its last CTA is done
            pack_halo<<<grid, kThreads, 0, h->stream>>>(h->s32[sorted].pt, h->d_cell_end, g, (int)S.halo_cap,
                                                        has_dn ? halo_slot(dn, S, parity, 1) : S.halo_send[0],
                                                        has_up ? halo_slot(up, S, parity, 0) : S.halo_send[1], S.mig_send[0], S.mig_send[1],
                                                        has_dn ? flag_of(dn, F_HALO_UP) : nullptr, has_up ? flag_of(up, F_HALO_DN) : nullptr,
                                                        tickets, S.seq);
        } else {
            pack_halo<<<grid, kThreads, 0, h->stream>>>(h->s32[sorted].pt, h->d_cell_end, g, (int)S.halo_cap, S.halo_send[0], S.halo_send[1],
                                                        S.mig_send[0], S.mig_send[1], nullptr, nullptr, nullptr, 0ull);
        }
        CUS(h, cudaGetLastError());
        S.phase = PLIFE_SLAB_FORCE;
        return PLIFE_OK;
    }
    if (phase == PLIFE_SLAB_FORCE) {
        const int sorted = h->cur ^ 1;
        dim3 grid(32, 2);
        const bool peer = S.peer_mode;
        unpack_halo<<<grid, kThreads, 0, h->stream>>>(h->s32[sorted].pt, h->d_cell_end, g, first, (int)h->n,
                                                      has_dn ? halo_in[0] : nullptr, has_up ? halo_in[1] : nullptr, d_err,
                                                      peer && has_dn ? flag_of(S.xbuf, F_HALO_DN) : nullptr,
                                                      peer && has_up ? flag_of(S.xbuf, F_HALO_UP) : nullptr, S.seq, S.spin_ns);
        CUS(h, cudaGetLastError());
        CUS(h, slab_force(h, g, dt));
        if (S.peer_mode) {
            unsigned int *tickets = reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned long long *>(S.xbuf) + F_TICKETS);
            if (has_dn) push_msg<<<8, kThreads, 0, h->stream>>>(S.mig_send[0], mig_slot(dn, S, parity, 1), 0, g.nx, (int)S.mig_cap,
                                                                flag_of(dn, F_MIG_UP), tickets + 2, S.seq);
            if (has_up) push_msg<<<8, kThreads, 0, h->stream>>>(S.mig_send[1], mig_slot(up, S, parity, 0), 0, g.nx, (int)S.mig_cap,
                                                                flag_of(up, F_MIG_DN), tickets + 3, S.seq);
            CUS(h, cudaGetLastError());
        }
        S.phase = PLIFE_SLAB_FINISH;
        return PLIFE_OK;
    }
    // PLIFE_SLAB_FINISH: headers back to the host (the only synchronisation of the step)
    {
        const bool waits = S.peer_mode && (has_dn || has_up);
        finish_headers<<<1, 1, 0, h->stream>>>(waits && has_dn ? flag_of(S.xbuf, F_MIG_DN) : nullptr, waits && has_up ? flag_of(S.xbuf, F_MIG_UP) : nullptr,
                                               S.seq, d_err, S.mig_send[0], S.mig_send[1], has_dn ? mig_in[0] : nullptr, has_up ? mig_in[1] : nullptr,
                                               S.h_hdr, S.spin_ns);
    }
    CUS(h, cudaGetLastError());
    CUS(h, cudaStreamSynchronize(h->stream));
    int4 hs[4];
    for (int k = 0; k < 4; k++) hs[k] = make_int4(S.h_hdr[k].x, S.h_hdr[k].y, S.h_hdr[k].z, S.h_hdr[k].w);
    const int err = S.h_hdr[4].x;
    S.phase = PLIFE_SLAB_SORT;
    S.seq++;
    const int sent_dn = hs[0].x, sent_up = hs[1].x;
    const int k_below = has_dn ? hs[2].x : 0, k_above = has_up ? hs[3].x : 0;
    if (err >> 16) return fail(h, PLIFE_ERR_STATE, "slab: timed out waiting for a neighbour's message");
    if (err) return fail(h, PLIFE_ERR_STATE, "slab: halo overflow / grid mismatch between ranks / ownership mismatch (raise halo_cap)");
    if (hs[0].y || hs[1].y) return fail(h, PLIFE_ERR_STATE, "slab: a particle crossed more than one slab in one step");
    if ((!has_dn && sent_dn) || (!has_up && sent_up)) return fail(h, PLIFE_ERR_STATE, "slab: particle left through a closed boundary");
    if (sent_dn > S.mig_cap || sent_up > S.mig_cap || k_below > S.mig_cap || k_above > S.mig_cap)
        return fail(h, PLIFE_ERR_STATE, "slab: migration message overflow (raise mig_cap)");
    const int64_t L = h->n; // residents before this step (the force pass wrote slots [0, L))
    if (L + k_below + k_above > h->cap) return fail(h, PLIFE_ERR_OOM, "slab: particle capacity exceeded by arrivals");
    const int cur = h->cur;
    CUS(h, cudaMemsetAsync(d_err, 0, sizeof(int), h->stream)); // errors raised from here on are reported by the next FINISH
    if (k_below > 0)
        append_arrivals<<<(k_below + kThreads - 1) / kThreads, kThreads, 0, h->stream>>>(mig_in[0], k_below, (int)L, g, h->s32[cur].pt,
                                                                                      h->s32[cur].vel, h->d_cell, h->d_count, d_err);
    if (k_above > 0)
        append_arrivals<<<(k_above + kThreads - 1) / kThreads, kThreads, 0, h->stream>>>(mig_in[1], k_above, (int)L + k_below, g,
                                                                                      h->s32[cur].pt, h->s32[cur].vel, h->d_cell, h->d_count, d_err);
    CUS(h, cudaGetLastError());
    S.n_old = L;
    S.k_below = k_below;
    S.k_above = k_above;
    h->n_phys = L + k_below + k_above;
    h->n = L - sent_dn - sent_up + k_below + k_above;
    h->steps++;
    return PLIFE_OK;
}

// all three phases back to back (peer mode only: nothing for the host to do in between)
int plife_slab_step(plife_handle *h, double dt, int32_t nsteps)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (!h->slab.on || !h->slab.peer_mode) return fail(h, PLIFE_ERR_STATE, "plife_slab_step needs the peer exchange mode");
    for (int s = 0; s < nsteps; s++) {
        if (h->stop_requested.exchange(0)) return fail(h, PLIFE_ERR_STOPPED, "stopped");
        for (int ph = PLIFE_SLAB_SORT; ph <= PLIFE_SLAB_FINISH; ph++) {
            int rc = plife_slab_phase(h, ph, dt);
            if (rc) return rc;
        }
    }
    return PLIFE_OK;
}

} // extern "C"
