// Multi-GPU slab decomposition of the physics step: one process per GPU, rank g owns
// grid rows [g*ny/G, (g+1)*ny/G).  The sorted order is row-major by cell, so a slab is one contiguous
// index range and its first / last rows are contiguous sub-ranges: packing a halo is a plain copy.
//
// Per step, all on the handle's stream and WITHOUT a host synchronisation - the particle counts of a rank change
// every step (migration) and live in device memory (SlabCounts); the host sizes grids by an upper bound and reads
// the true counts back lazily, a few steps late, from a ring in mapped pinned memory:
//   phase SORT    cell-list build of the owned particles; pack first and last owned row into the halo
//                 messages {count | per-bin END offsets of the row | 16-byte particle records} - peer exchange:
//                 written straight into the neighbours' receive slots over NVLink, then a flag is raised
//   phase FORCE   force/integrate over the INTERIOR rows (they need no ghost row: the halo is in flight
//                 meanwhile); wait for the neighbours' halo flags and place the ghost rows around the owned block
//                 [ghost below | owned | ghost above]; force/integrate over the first and last owned row.  The
//                 epilogue bins the new positions and appends particles whose new row belongs to a neighbour to
//                 the migration messages, which are then pushed to the neighbours
//   phase FINISH  wait for the neighbours' migration messages, validate, compute the new counts on the device,
//                 append the arrivals (binning them)
//
// Two ways to move the messages:
//   external  the host exchanges caller-provided buffers between the phases (torch.distributed/NCCL);
//   peer      (default) the library owns the buffers, ranks map each other's receive slots with CUDA IPC
//             and the kernels push a message straight into the neighbour's slot over NVLink, then raise a
//             flag (sequence number) there; consumers spin on their own flag (bounded by wall time).
//             Slots are double-buffered by step parity: a rank can only be one message ahead of
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
cudaError_t slab_force(plife_handle *h, const Grid &g, double dt, const int *d_tr, int nblocks, int bin_lo, int bin_hi, cudaStream_t stream,
                       bool first_part, bool last_part);
cudaError_t slab_force_done(plife_handle *h, const Grid &g);
int slab_reset_capacity(plife_handle *h);
} // namespace plife

namespace {

constexpr int kThreads = 256;

// sticky error bits of SlabCounts::err
enum {
    kErrHalo = 1,      // halo row did not fit / grid or bin mismatch between ranks
    kErrTimeout = 2,   // a neighbour's message did not arrive within the wait limit
    kErrFar = 4,       // a particle crossed more than one slab in one step, or left the slab from an interior row (>= 1 row in one step)
    kErrClosed = 8,    // a particle left through a closed boundary
    kErrMigCap = 16,   // migration message overflow
    kErrCapacity = 32, // arrivals exceed the particle capacity
    kErrOwner = 64,    // sender and receiver disagree about ownership
    kErrBound = 128,   // the host's grid-size bound was below the true particle count
};

__device__ __forceinline__ int offsets_records(int nx) { return (nx + 3) >> 2; }

// wall-clock nanoseconds (not SM cycles: the spin limits below are times, whatever the clock does)
__device__ __forceinline__ unsigned long long wall_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag >= seq; false on timeout
__device__ __forceinline__ bool wait_flag(const volatile unsigned long long *flag, unsigned long long seq, unsigned long long spin_ns)
{
    const unsigned long long t0 = wall_ns();
    while (*flag < seq) {
        if (wall_ns() - t0 > spin_ns) return false;
        __nanosleep(100);
    }
    __threadfence_system();
    return true;
}

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

// dir 0: first owned row (local row 1) -> becomes the down neighbour's top ghost row
// dir 1: last owned row (local row nly-2) -> becomes the up neighbour's bottom ghost row
// `msg0/msg1` are where the two messages are written: the local send buffers (external exchange), or - peer exchange -
// straight into the neighbours' receive slots over NVLink, followed by the flag (flag0/flag1 non-NULL).
// (The target ranges of the force launches and the migration cursors are prepared by the gather kernel, cells.cu.)
__global__ void __launch_bounds__(kThreads) pack_halo(const float4 *__restrict__ pt_sorted, const int32_t *__restrict__ cell_end,
                                                      Grid g, int halo_cap, float4 *__restrict__ msg0, float4 *__restrict__ msg1,
                                                      volatile unsigned long long *flag0, volatile unsigned long long *flag1,
                                                      unsigned int *tickets, unsigned long long seq)
{
    const int dir = blockIdx.y;
    float4 *msg = dir ? msg1 : msg0;
    volatile unsigned long long *flag = dir ? flag1 : flag0;
    const int nxk = g.nxk(); // the message carries one END offset per fine bin of the row
    const int row = dir ? g.nly - 2 : 1;
    const int start = __ldg(cell_end + row * nxk - 1);
    const int end = __ldg(cell_end + (row + 1) * nxk - 1);
    const int count = end - start;
    const int noff = offsets_records(nxk);
    const int t = blockIdx.x * kThreads + threadIdx.x;
    if (t == 0) {
        int4 hd = make_int4(count <= halo_cap ? count : -count - 1, nxk, 0, 0);
        msg[0] = *reinterpret_cast<float4 *>(&hd);
    }
    if (count <= halo_cap) { // else overflow: reported by the receiver
        int32_t *off = reinterpret_cast<int32_t *>(msg + 1);
        for (int c = t; c < 4 * noff; c += gridDim.x * kThreads) off[c] = c < nxk ? __ldg(cell_end + row * nxk + c) - start : 0;
        for (int k = t; k < count; k += gridDim.x * kThreads) msg[1 + noff + k] = __ldg(pt_sorted + start + k);
    }
    if (flag) signal_when_all_done(tickets + dir, flag, seq, gridDim.x);
}

// which 0: ghost row below (local row 0), right-aligned before the owned block at `first`
// which 1: ghost row above (local row nly-1), placed after the owned block
__global__ void __launch_bounds__(kThreads) unpack_halo(float4 *__restrict__ pt_sorted, int32_t *__restrict__ cell_end, Grid g,
                                                        int first, const SlabCounts *__restrict__ cnt, int *__restrict__ err, const float4 *msg0,
                                                        const float4 *msg1, const volatile unsigned long long *flag0,
                                                        const volatile unsigned long long *flag1, unsigned long long seq, unsigned long long spin_ns)
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
        if (threadIdx.x == 0) s_timeout = wait_flag(flag, seq, spin_ns) ? 0 : 1;
        __syncthreads();
        if (s_timeout) {
            if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(err, kErrTimeout);
            return;
        }
    }
    const int4 hd = __ldcg(reinterpret_cast<const int4 *>(msg));
    const int t = blockIdx.x * kThreads + threadIdx.x;
    const int nxk = g.nxk();
    if (hd.x < 0 || hd.y != nxk) { // overflow at the sender, or the ranks disagree about the grid / the bins per cell
        if (t == 0) atomicOr(err, kErrHalo);
        return;
    }
    const int n = cnt->n;
    const int count = hd.x;
    const int noff = offsets_records(nxk);
    const int base = which ? first + n : first - count;
    const int row = which ? g.nly - 1 : 0;
    const int32_t *off = reinterpret_cast<const int32_t *>(msg + 1);
    for (int c = t; c < nxk; c += gridDim.x * kThreads) cell_end[row * nxk + c] = base + __ldcg(off + c);
    if (which == 0 && t == 0) cell_end[-1] = base;
    for (int k = t; k < count; k += gridDim.x * kThreads) pt_sorted[base + k] = __ldcg(msg + 1 + noff + k);
}

// Phase FINISH as ONE kernel of a few co-resident CTAs (gridDim = (kMigCtas, 2), blockIdx.y = direction):
//   push    every CTA copies its share of the two migration messages (their used records only) into the neighbours'
//           receive slots over NVLink; the last CTA per direction raises the neighbour's flag (peer mode)
//   finish  CTA (0,0): wait for the neighbours' messages, validate the four headers, compute the counts of the next
//           step, publish them (device struct + the host's ring in mapped pinned memory), then release the other CTAs
//   append  every CTA places its share of the arrivals.  Message records {x,y,type,id},{vx,vy,source slot,-}: each
//           goes to base + (rank of its source slot among the arrivals of the same message) - the sender appended them
//           with an atomic cursor, the rank restores the sender's array order - and is binned.
// The CTAs wait for each other inside the kernel, so all of them must be resident: the grid is 16 CTAs.
constexpr int kMigCtas = 8;

__global__ void __launch_bounds__(kThreads) slab_migrate(const float4 *__restrict__ ms0, const float4 *__restrict__ ms1, float4 *peer0, float4 *peer1,
                                                         volatile unsigned long long *pflag0, volatile unsigned long long *pflag1,
                                                         unsigned int *tickets, const volatile unsigned long long *f0,
                                                         const volatile unsigned long long *f1, unsigned long long seq, const SlabCounts *cnt,
                                                         SlabCounts *nxt, int *errw, const float4 *mi0, const float4 *mi1, int has_dn, int has_up, int mig_cap, int cap, int bound,
                                                         volatile SlabCounts *ring, volatile unsigned long long *go, unsigned long long spin_ns, Grid g,
                                                         float4 *__restrict__ pt, float2 *__restrict__ vel, int32_t *__restrict__ cell,
                                                         int32_t *__restrict__ count)
{
    const int dir = blockIdx.y;
    // ---- push ----
    {
        const float4 *local = dir ? ms1 : ms0;
        float4 *peer = dir ? peer1 : peer0;
        if (peer) {
            const int c = *reinterpret_cast<const int *>(local);
            const int nrec = 1 + 2 * min(c, mig_cap);
            for (int k = blockIdx.x * kThreads + threadIdx.x; k < nrec; k += gridDim.x * kThreads) peer[k] = local[k];
            signal_when_all_done(tickets + dir, dir ? pflag1 : pflag0, seq, gridDim.x);
        }
    }
    // ---- finish ----
    if (blockIdx.x == 0 && dir == 0) {
        if (threadIdx.x == 0) {
            int err = 0;
            if (f0 && !wait_flag(f0, seq, spin_ns)) err |= kErrTimeout;
            if (f1 && !wait_flag(f1, seq, spin_ns)) err |= kErrTimeout;
            const volatile int *s0 = reinterpret_cast<const volatile int *>(ms0), *s1 = reinterpret_cast<const volatile int *>(ms1);
            const int sent_dn = s0[0], sent_up = s1[0];
            if (s0[1] || s1[1]) err |= kErrFar;
            if ((!has_dn && sent_dn) || (!has_up && sent_up)) err |= kErrClosed;
            int k_below = 0, k_above = 0;
            if (!(err & kErrTimeout)) {
                if (has_dn && mi0) k_below = reinterpret_cast<const volatile int *>(mi0)[0];
                if (has_up && mi1) k_above = reinterpret_cast<const volatile int *>(mi1)[0];
            }
            if (sent_dn > mig_cap || sent_up > mig_cap || k_below > mig_cap || k_above > mig_cap) {
                err |= kErrMigCap;
                k_below = min(k_below, mig_cap);
                k_above = min(k_above, mig_cap);
            }
            const int L = cnt->n; // residents before this step (the force pass wrote slots [0, L))
            if (L + k_below + k_above > cap) {
                err |= kErrCapacity;
                k_below = k_above = 0;
            }
            if (L > bound || cnt->n_phys > bound) err |= kErrBound;
            if (err) atomicOr(errw, err);
            err = *reinterpret_cast<volatile int *>(errw); // everything found so far, by any kernel (sticky)
            SlabCounts c;
            c.n_old = L;
            c.k_below = k_below;
            c.k_above = k_above;
            c.n_phys = L + k_below + k_above;
            c.n = L - min(sent_dn, mig_cap) - min(sent_up, mig_cap) + k_below + k_above;
            c.err = err;
            c.sent_dn = sent_dn;
            c.sent_up = sent_up;
            c.seq = seq;
            *nxt = c; // the next step reads the other struct: this kernel may run next to this step's interior force launch
            __threadfence();
            *go = seq; // release the other CTAs of this kernel before the (slow) host-visible copy
            volatile SlabCounts *r = ring + (seq & 7);
            r->seq = 0;
            __threadfence_system();
            r->n = c.n; r->n_phys = c.n_phys; r->n_old = c.n_old; r->k_below = c.k_below; r->k_above = c.k_above;
            r->err = c.err; r->sent_dn = c.sent_dn; r->sent_up = c.sent_up;
            __threadfence_system();
            r->seq = seq;
            __threadfence_system();
        }
    } else if (threadIdx.x == 0) {
        const unsigned long long t0 = wall_ns();
        while (*go < seq && wall_ns() - t0 < 2 * spin_ns) __nanosleep(50);
        __threadfence();
    }
    __syncthreads();
    // ---- append ----
    const float4 *msg = dir ? mi1 : mi0;
    const volatile SlabCounts *vc = nxt;
    const int k = dir ? vc->k_above : vc->k_below;
    if (!msg || k == 0) return;
    const int base = vc->n_old + (dir ? vc->k_below : 0);
    for (int j = blockIdx.x * kThreads + threadIdx.x; j < k; j += gridDim.x * kThreads) {
        const float4 a = __ldcg(msg + 1 + 2 * j), b = __ldcg(msg + 2 + 2 * j);
        const int src = __float_as_int(b.z);
        int rank = 0;
        for (int q = 0; q < k; ++q) rank += (__float_as_int(__ldcg(msg + 2 + 2 * q).z) < src) ? 1 : 0;
        const int dst = base + rank;
        pt[dst] = a;
        vel[dst] = make_float2(b.x, b.y);
        const int cxy = cell_coords((double)a.x, (double)a.y, g);
        const int c = container_of(cxy, g);
        cell[dst] = c < 0 ? -1 : cxy;
        if (c >= 0) atomicAdd(count + c, 1);
        else atomicOr(errw, kErrOwner); // sender and receiver disagree about ownership
    }
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

int report(plife_handle *h, int err)
{
    if (!err) return PLIFE_OK;
    h->slab.err_seen = err;
    if (err & kErrTimeout) return fail(h, PLIFE_ERR_STATE, "slab: timed out waiting for a neighbour's message");
    if (err & kErrFar) return fail(h, PLIFE_ERR_STATE, "slab: a particle moved too far in one step (left its slab from an interior row, or crossed more than one slab)");
    if (err & kErrClosed) return fail(h, PLIFE_ERR_STATE, "slab: particle left through a closed boundary");
    if (err & kErrMigCap) return fail(h, PLIFE_ERR_STATE, "slab: migration message overflow (raise mig_cap)");
    if (err & kErrCapacity) return fail(h, PLIFE_ERR_OOM, "slab: particle capacity exceeded by arrivals");
    if (err & kErrBound) return fail(h, PLIFE_ERR_STATE, "slab: particle count outran the host's launch bound (internal)");
    return fail(h, PLIFE_ERR_STATE, "slab: halo overflow / grid mismatch between ranks / ownership mismatch (raise halo_cap)");
}

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
    if (S.h_ring) cudaFreeHost((void *)S.h_ring);
    cudaFree(S.counts);
    cudaFree(S.d_err);
    cudaFree(S.d_tr);
    cudaFree((void *)S.d_go);
    for (int k = 0; k < 4; k++)
        if (S.step_done[k]) cudaEventDestroy(S.step_done[k]);
    if (S.ev_sorted) cudaEventDestroy(S.ev_sorted);
    if (S.ev_edge) cudaEventDestroy(S.ev_edge);
    if (S.side) cudaStreamDestroy(S.side);
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

// The host's view of the counts.  block = false: take the newest entry the finish kernels have published so far (the
// host may be a few steps ahead of the device) and derive the launch bound from it; block = true: drain the stream
// first, so the counts are exact (plife_count, download, snapshot, upload ...).
int slab_refresh(plife_handle *h, bool block)
{
    SlabState &S = h->slab;
    if (!S.on || !S.counts) return PLIFE_OK;
    if (block) {
        cudaError_t e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) {
            h->poisoned = true;
            return slab_fail(h, PLIFE_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    const unsigned long long done = S.seq - 1; // last step whose FINISH has been queued
    for (unsigned long long s = done; s > S.seq_known && s + 8 > done; --s) {
        volatile SlabCounts *r = S.h_ring + (s & 7);
        if (r->seq != s) continue;
        SlabCounts c;
        c.n = r->n; c.n_phys = r->n_phys; c.n_old = r->n_old; c.k_below = r->k_below; c.k_above = r->k_above;
        c.err = r->err; c.sent_dn = r->sent_dn; c.sent_up = r->sent_up;
        if (r->seq != s) continue; // overwritten while reading (cannot happen with <= 4 steps in flight)
        S.seq_known = s;
        h->n = c.n;
        h->n_phys = c.n_phys;
        S.n_old = c.n_old;
        S.k_below = c.k_below;
        S.k_above = c.k_above;
        S.max_arrivals = S.max_arrivals > c.k_below + c.k_above ? S.max_arrivals : c.k_below + c.k_above;
        if (c.err && !S.err_seen) S.err_pending = c.err;
        break;
    }
    // growth a step can bring: what has been seen, with a generous margin; the device reports kErrBound if it is ever wrong
    const int64_t lag = (int64_t)(done - S.seq_known);
    const int64_t per_step = 4 * S.max_arrivals + 4096;
    int64_t b = h->n_phys + (lag + 1) * per_step;
    S.n_bound = b < h->cap ? b : h->cap;
    if (block && (S.err_pending || S.err_seen)) return report(h, S.err_pending | S.err_seen); // a synchronising call reports what the device found
    return PLIFE_OK;
}

// counts known exactly on the host (after upload / init): write them to the device
int slab_set_counts(plife_handle *h)
{
    SlabState &S = h->slab;
    if (!S.on || !S.counts) return PLIFE_OK;
    SlabCounts c{};
    c.n = (int)h->n;
    c.n_phys = (int)h->n_phys;
    c.n_old = (int)h->n_phys;
    c.seq = S.seq - 1;
    cudaError_t e = cudaMemcpyAsync(S.cnt(), &c, sizeof c, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(S.d_err, 0, sizeof(int), h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return slab_fail(h, PLIFE_ERR_CUDA, cudaGetErrorString(e));
    S.seq_known = S.seq - 1;
    S.n_old = h->n_phys;
    S.k_below = S.k_above = 0;
    S.n_bound = h->n_phys;
    S.err_pending = S.err_seen = 0;
    return PLIFE_OK;
}
} // namespace plife

extern "C" {

// header + one END offset per fine bin of a row (at most 8 bins per cell) + the records
int64_t plife_slab_halo_records(int32_t nx, int64_t halo_cap) { return 1 + (8 * (int64_t)nx + 3) / 4 + halo_cap; }
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
        if (e != cudaSuccess) {
            slab_release(h);
            return fail(h, PLIFE_ERR_OOM, "slab_configure: allocating exchange buffers failed");
        }
    }
    cudaError_t e = cudaMalloc((void **)&S.counts, 2 * sizeof(SlabCounts));
    if (e == cudaSuccess) e = cudaMemset(S.counts, 0, 2 * sizeof(SlabCounts));
    if (e == cudaSuccess) e = cudaMalloc((void **)&S.d_err, sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(S.d_err, 0, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&S.d_go, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset((void *)S.d_go, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc((void **)&S.d_tr, 12 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(S.d_tr, 0, 12 * sizeof(int));
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&S.h_ring, 8 * sizeof(SlabCounts), cudaHostAllocMapped);
    for (int k = 0; k < 4 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&S.step_done[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.ev_sorted, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.ev_edge, cudaEventDisableTiming);
    if (e == cudaSuccess && S.peer_mode) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi); // hi = greatest priority (numerically lowest)
        e = cudaStreamCreateWithPriority(&S.side, cudaStreamNonBlocking, hi);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        slab_release(h);
        return fail(h, PLIFE_ERR_OOM, "slab_configure: allocating the count blocks failed");
    }
    memset((void *)S.h_ring, 0, 8 * sizeof(SlabCounts));
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
    SlabState &S = h->slab;
    if (S.err_seen) return report(h, S.err_seen);
    Grid g;
    int rc = slab_make_grid(h, &g);
    if (rc) return rc;
    if (g.nx > S.nx_cfg) return fail(h, PLIFE_ERR_STATE, "slab: rmax shrank since plife_slab_configure (halo messages would not fit): reconfigure");
    const bool wrap = h->settings.wrap != 0;
    const bool has_dn = S.world > 1 && (wrap || S.rank > 0);
    const bool has_up = S.world > 1 && (wrap || S.rank < S.world - 1);
    if (S.peer_mode && ((has_dn && !S.peer_base[0]) || (has_up && !S.peer_base[1])))
        return fail(h, PLIFE_ERR_STATE, "slab: neighbours not connected (plife_slab_connect_ipc / _local)");
    const int first = (int)S.halo_cap;
    const int parity = (int)(S.seq & 1);
    float4 *dn = S.peer_base[0], *up = S.peer_base[1];
    // where this step's messages arrive
    float4 *halo_in[2], *mig_in[2];
    for (int d = 0; d < 2; d++) {
        halo_in[d] = S.peer_mode ? halo_slot(S.xbuf, S, parity, d) : S.halo_recv[d];
        mig_in[d] = S.peer_mode ? mig_slot(S.xbuf, S, parity, d) : S.mig_recv[d];
    }
    unsigned int *tickets = S.peer_mode ? reinterpret_cast<unsigned int *>(reinterpret_cast<unsigned long long *>(S.xbuf) + F_TICKETS) : nullptr; // 4 local counters

    // Peer exchange: the halo traffic and the two edge rows run on a second, high-priority stream next to the interior
    // force launch, so neither the wait for the neighbours nor the tail of the small edge launch is on the critical path:
    //   main:  scan, scatter, gather -> [sorted]      force(interior) ................ wait [edge] -> push migrants, finish
    //   side:              wait [sorted] -> pack halo (push + signal), wait + unpack halo, force(edge rows) -> [edge]
    // External exchange (the host moves the messages between the phases): everything stays on the main stream.
    cudaStream_t side = S.peer_mode ? S.side : h->stream;
    if (phase == PLIFE_SLAB_SORT) {
        // no more than 3 steps queued ahead of the device: bounds the slack of the launch bound and the count ring
        if (S.seq > 4) CUS(h, cudaEventSynchronize(S.step_done[(S.seq - 4) & 3]));
        rc = slab_refresh(h, false);
        if (rc) return rc;
        if (S.err_pending) return report(h, S.err_pending);
        rc = slab_sort(h, g); // bin (if needed), scan, scatter, gather: owned block of the sorted array
        if (rc) return rc;
        const int sorted = h->cur ^ 1;
        dim3 grid(32, 2);
        if (S.peer_mode) {
            CUS(h, cudaEventRecord(S.ev_sorted, h->stream));
            CUS(h, cudaStreamWaitEvent(side, S.ev_sorted, 0));
            // my first row is the down neighbour's ghost row ABOVE its slab (its slot dir 1), and vice versa; the pack kernel
            // writes it there and raises the neighbour's flag when its last CTA is done
            pack_halo<<<grid, kThreads, 0, side>>>(h->s32[sorted].pt, h->d_cell_end, g, (int)S.halo_cap,
                                                   has_dn ? halo_slot(dn, S, parity, 1) : S.halo_send[0],
                                                   has_up ? halo_slot(up, S, parity, 0) : S.halo_send[1],
                                                   has_dn ? flag_of(dn, F_HALO_UP) : nullptr, has_up ? flag_of(up, F_HALO_DN) : nullptr, tickets, S.seq);
        } else {
            pack_halo<<<grid, kThreads, 0, side>>>(h->s32[sorted].pt, h->d_cell_end, g, (int)S.halo_cap, S.halo_send[0], S.halo_send[1], nullptr,
                                                   nullptr, nullptr, 0ull);
        }
        CUS(h, cudaGetLastError());
        S.phase = PLIFE_SLAB_FORCE;
        return PLIFE_OK;
    }
    // phase FINISH as one launch: push the migrants, wait for the neighbours', new counts, append the arrivals.
    // `vel_new` = the velocity buffer this step's force launches write (the arrivals' velocities join it).
    auto launch_migrate = [&](cudaStream_t st, float2 *vel_new) -> cudaError_t {
        const bool peer = S.peer_mode;
        dim3 mg(kMigCtas, 2);
        slab_migrate<<<mg, kThreads, 0, st>>>(
            S.mig_send[0], S.mig_send[1], peer && has_dn ? mig_slot(dn, S, parity, 1) : nullptr, peer && has_up ? mig_slot(up, S, parity, 0) : nullptr,
            peer && has_dn ? flag_of(dn, F_MIG_UP) : nullptr, peer && has_up ? flag_of(up, F_MIG_DN) : nullptr, peer ? tickets + 2 : nullptr,
            peer && has_dn ? flag_of(S.xbuf, F_MIG_DN) : nullptr, peer && has_up ? flag_of(S.xbuf, F_MIG_UP) : nullptr, S.seq, S.cnt(), S.nxt(), S.d_err,
            has_dn ? mig_in[0] : nullptr, has_up ? mig_in[1] : nullptr, has_dn ? 1 : 0, has_up ? 1 : 0, (int)S.mig_cap, (int)h->cap, (int)S.n_bound,
            S.h_ring, S.d_go, S.spin_ns, g, h->s32[h->cur].pt, vel_new, h->d_cell, h->d_count);
        return cudaGetLastError();
    };
    if (phase == PLIFE_SLAB_FORCE) {
        const int sorted = h->cur ^ 1;
        const bool peer = S.peer_mode;
        const int nxk = g.nxk();
        // interior rows: they read owned rows only, so the halo may still be in flight.  Their particles must stay in the slab
        // (a particle may enter a neighbour slab only from the first / last owned row, i.e. move less than a row per step
        // towards it): the migration exchange runs next to this launch, a late leaver raises kErrFar.
        const int nb_all = (int)((S.n_bound + 127) / 128) + 1;
        const int nb_edge = (int)((2 * S.halo_cap + 127) / 128) + 2;
        CUS(h, slab_force(h, g, dt, S.d_tr, nb_all, nxk, (g.nly - 1) * nxk - 1, h->stream, true, false));
        dim3 grid(32, 2);
        unpack_halo<<<grid, kThreads, 0, side>>>(h->s32[sorted].pt, h->d_cell_end, g, first, S.cnt(), S.d_err,
                                                 has_dn ? halo_in[0] : nullptr, has_up ? halo_in[1] : nullptr,
                                                 peer && has_dn ? flag_of(S.xbuf, F_HALO_DN) : nullptr,
                                                 peer && has_up ? flag_of(S.xbuf, F_HALO_UP) : nullptr, S.seq, S.spin_ns);
        CUS(h, cudaGetLastError());
        CUS(h, slab_force(h, g, dt, S.d_tr + 4, nb_edge, 0, nxk * g.nly - 1, side, false, true));
        if (S.peer_mode) {
            // the whole migration exchange also runs on the side stream, next to the interior launch: nothing of it is left on
            // the critical path (the arrivals' velocities go into the buffer the force launches are writing)
            CUS(h, launch_migrate(side, h->s32[h->cur ^ 1].vel));
            CUS(h, cudaEventRecord(S.ev_edge, side));
            CUS(h, cudaStreamWaitEvent(h->stream, S.ev_edge, 0));
        }
        CUS(h, slab_force_done(h, g));
        S.phase = PLIFE_SLAB_FINISH;
        return PLIFE_OK;
    }
    // PLIFE_SLAB_FINISH.  Peer exchange: everything is queued already.  External exchange: the host has moved the migration
    // messages between the phases; consume them now.
    if (!S.peer_mode) CUS(h, launch_migrate(h->stream, h->s32[h->cur].vel));
    CUS(h, cudaEventRecord(S.step_done[S.seq & 3], h->stream));
    S.phase = PLIFE_SLAB_SORT;
    S.seq++;
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
