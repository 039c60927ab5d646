// Per-step cell-list build: the reference's serial counting sort
// (Physics.makeContainers, B/Physics.java:309-354) as four data-parallel passes
//   K_BIN      cell id per particle + per-cell histogram      (:329-332)
//   K_SCAN     exclusive prefix over the nx*ny cells          (:335-340)
//   K_SCATTER  cursor scatter of source indices               (:343-348)
//   K_GATHER   stable in-cell rank + reorder of the SoA state (:343-353)
// plus the small utility kernels (type histogram, seeded generator, snapshot).
//
// Stability: the reference's scatter is stable (array order inside a cell is
// the previous array order).  The atomic cursor scatter is not, so K_GATHER
// ranks every slot's source index among the source indices of its cell and
// writes to start+rank; the result is exactly the reference's permutation.
#include <stdlib.h>

#include "plife_internal.h"

namespace plife {

namespace {

constexpr int kThreads = 256;


__global__ void __launch_bounds__(kThreads) bin_f32(const float4 *__restrict__ pt, int n, Grid g,
                                                    int32_t *__restrict__ cell, int32_t *__restrict__ count)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float4 p = __ldg(&pt[i]);
    int cxy = cell_coords_fast((double)p.x, (double)p.y, g);
    int c = container_of(cxy, g);
    cell[i] = c < 0 ? -1 : cxy; // slab mode: uploads hold owned particles only; anything else is dropped
    if (c >= 0) atomicAdd(&count[c], 1);
}

__global__ void __launch_bounds__(kThreads) bin_f64(const double2 *__restrict__ pos, int n, Grid g,
                                                    int32_t *__restrict__ cell, int32_t *__restrict__ count)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    double2 p = __ldg(&pos[i]);
    int cxy = cell_coords_fast(p.x, p.y, g);
    cell[i] = cxy;
    atomicAdd(&count[container_of(cxy, g)], 1);
}

// ---- exclusive scan over the (fine) bins: tile sums -> scan of sums -> apply ----
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems; // 4096 bins per CTA
using u64 = unsigned long long;

__device__ __forceinline__ int warp_inclusive_scan(int v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix and the block total
__device__ __forceinline__ int block_exclusive_scan(int v, int &total, int *smem /* >= 33 */)
{
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = warp_inclusive_scan(v, lane);
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int w = lane < nw ? smem[lane] : 0;
        int winc = warp_inclusive_scan(w, lane);
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    int r = smem[warp] + inc - v;
    total = smem[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int32_t *__restrict__ count, int64_t ncell,
                                                               int32_t *__restrict__ tile_sums)
{
    __shared__ int sm[33];
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int s = 0;
    // vectorised: each thread sums 16 consecutive bins (4 x int4)
    int64_t first = base + (int64_t)threadIdx.x * kScanItems;
    if (first + kScanItems <= ncell) {
        const int4 *p = reinterpret_cast<const int4 *>(count + first);
#pragma unroll
        for (int k = 0; k < kScanItems / 4; k++) {
            int4 v = __ldg(p + k);
            s += v.x + v.y + v.z + v.w;
        }
    } else {
        for (int k = 0; k < kScanItems; k++)
            if (first + k < ncell) s += count[first + k];
    }
    int total;
    block_exclusive_scan(s, total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_sums(int32_t *__restrict__ tile_sums, int ntiles, int carry0,
                                                  int32_t *__restrict__ cell_end)
{
    __shared__ int sm[33];
    int carry = carry0; // particle offsets start after the ghost-below capacity
    if (threadIdx.x == 0) cell_end[-1] = carry0; // start of local bin 0
    for (int base = 0; base < ntiles; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < ntiles ? tile_sums[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, total, sm);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}

// writes the exclusive prefixes (the cursor start of every bin) and zeroes count for the next step
__global__ void __launch_bounds__(kScanThreads) scan_apply(int32_t *__restrict__ count, int64_t ncell,
                                                           const int32_t *__restrict__ tile_sums,
                                                           int32_t *__restrict__ cell_end)
{
    __shared__ int sm[33];
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t first = base + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
    bool full = first + kScanItems <= ncell;
    if (full) {
        int4 *p = reinterpret_cast<int4 *>(count + first);
#pragma unroll
        for (int k = 0; k < kScanItems / 4; k++) {
            int4 q = p[k];
            v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
            p[k] = make_int4(0, 0, 0, 0);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            v[k] = 0;
            if (first + k < ncell) {
                v[k] = count[first + k];
                count[first + k] = 0;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < kScanItems; k++) s += v[k];
    int total;
    int ex = block_exclusive_scan(s, total, sm) + tile_sums[blockIdx.x];
    if (full) {
        int4 *o = reinterpret_cast<int4 *>(cell_end + first);
#pragma unroll
        for (int k = 0; k < kScanItems / 4; k++) {
            int4 q;
            q.x = ex; ex += v[4 * k];
            q.y = ex; ex += v[4 * k + 1];
            q.z = ex; ex += v[4 * k + 2];
            q.w = ex; ex += v[4 * k + 3];
            o[k] = q;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            if (first + k < ncell) cell_end[first + k] = ex;
            ex += v[k];
        }
    }
}

// ---- the same scan in ONE launch: persistent CTAs, reduce - look back - scan ----
// A few hundred CTAs (all resident) each own a contiguous chunk of tiles.  Pass 1 sums the chunk and publishes the sum
// (AGG); the CTA then looks back over its predecessors' status words - 32 at a time, until it meets one that already
// carries an inclusive prefix - and publishes its own inclusive prefix (PREFIX); pass 2 re-reads the chunk (L2-hot: a
// chunk is ~100 KB), scans it with the carry, writes the offsets and zeroes the histogram.  Chunks are handed out by a
// ticket, so a chunk's predecessors are always running or done.  Status words are tagged with the launch's epoch, which
// the last CTA to finish advances (together with re-arming the ticket): nothing has to be cleared between launches and a
// captured CUDA graph can replay the kernel.  (A tile-per-CTA chained scan was measured first: with ~2000 tiles in flight
// at once the look-back walks of the first wave made it slower than the three-launch scan.)
struct ScanState {
    unsigned int ticket, done, epoch, pad;
};
constexpr unsigned long long kScanAgg = 1ull << 32, kScanPrefix = 2ull << 32;
constexpr int kScanCtasPerSm = 4;

__global__ void __launch_bounds__(kScanThreads) scan_onepass(int32_t *__restrict__ count, int64_t ncell, int ntiles, int tiles_per_cta,
                                                             int carry0, int32_t *__restrict__ cell_end, ScanState *st,
                                                             volatile unsigned long long *status)
{
    __shared__ int sm[33];
    __shared__ unsigned int s_chunk, s_epoch;
    __shared__ int s_prefix;
    if (threadIdx.x == 0) {
        s_chunk = atomicAdd(&st->ticket, 1u);
        s_epoch = *reinterpret_cast<volatile unsigned int *>(&st->epoch);
    }
    __syncthreads();
    const int chunk = (int)s_chunk;
    const unsigned long long tag = (unsigned long long)(s_epoch & 0x3fffffffu) << 34;
    const int t0 = chunk * tiles_per_cta, t1 = min(t0 + tiles_per_cta, ntiles);
    // pass 1: the chunk's sum
    int s = 0;
    for (int tile = t0; tile < t1; ++tile) {
        const int64_t first = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
        if (first + kScanItems <= ncell) {
            const int4 *p = reinterpret_cast<const int4 *>(count + first);
#pragma unroll
            for (int k = 0; k < kScanItems / 4; k++) {
                const int4 q = p[k];
                s += q.x + q.y + q.z + q.w;
            }
        } else {
            for (int k = 0; k < kScanItems; k++)
                if (first + k < ncell) s += count[first + k];
        }
    }
    int total;
    block_exclusive_scan(s, total, sm);
    if (threadIdx.x < 32) { // warp 0 publishes and looks back
        const int lane = threadIdx.x;
        if (lane == 0) {
            __threadfence();
            status[chunk] = tag | (chunk == 0 ? kScanPrefix : kScanAgg) | (unsigned int)total;
        }
        int prefix = 0;
        for (int base = chunk - 1; base >= 0; base -= 32) {
            const int t = base - lane;
            unsigned long long w = 0;
            for (;;) { // wait until every word of this window is from this launch
                w = t >= 0 ? status[t] : (tag | kScanPrefix);
                const bool ready = (w >> 34) == (tag >> 34) && ((w >> 32) & 3ull) != 0;
                if (__all_sync(0xffffffffu, ready)) break;
            }
            const unsigned pmask = __ballot_sync(0xffffffffu, ((w >> 32) & 3ull) == 2ull);
            const int stop = pmask ? __ffs(pmask) - 1 : 31; // nearest predecessor that carries a prefix
            int val = lane <= stop ? (int)(unsigned int)w : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            prefix += val;
            if (pmask) break;
        }
        if (lane == 0) {
            if (chunk > 0) {
                __threadfence();
                status[chunk] = tag | kScanPrefix | (unsigned int)(prefix + total);
            }
            s_prefix = prefix;
        }
    }
    __syncthreads();
    int carry = s_prefix + carry0;
    if (chunk == 0 && threadIdx.x == 0) cell_end[-1] = carry0; // start of local bin 0
    // pass 2: scan the chunk tile by tile with the running carry, zero the histogram
    for (int tile = t0; tile < t1; ++tile) {
        const int64_t first = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
        int v[kScanItems];
        int ts = 0;
        const bool full = first + kScanItems <= ncell;
        if (full) {
            int4 *p = reinterpret_cast<int4 *>(count + first);
#pragma unroll
            for (int k = 0; k < kScanItems / 4; k++) {
                int4 q = p[k];
                v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
                p[k] = make_int4(0, 0, 0, 0);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kScanItems; k++) {
                v[k] = 0;
                if (first + k < ncell) {
                    v[k] = count[first + k];
                    count[first + k] = 0;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kScanItems; k++) ts += v[k];
        int tile_total;
        int ex = block_exclusive_scan(ts, tile_total, sm) + carry;
        carry += tile_total;
        if (full) {
            int4 *o = reinterpret_cast<int4 *>(cell_end + first);
#pragma unroll
            for (int k = 0; k < kScanItems / 4; k++) {
                int4 q;
                q.x = ex; ex += v[4 * k];
                q.y = ex; ex += v[4 * k + 1];
                q.z = ex; ex += v[4 * k + 2];
                q.w = ex; ex += v[4 * k + 3];
                o[k] = q;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kScanItems; k++) {
                if (first + k < ncell) cell_end[first + k] = ex;
                ex += v[k];
            }
        }
    }
    if (threadIdx.x == 0) { // the last CTA to finish re-arms the state for the next launch
        __threadfence();
        if (atomicAdd(&st->done, 1u) == gridDim.x - 1u) {
            st->ticket = 0;
            st->done = 0;
            __threadfence();
            st->epoch = s_epoch + 1u;
        }
    }
}

// ---- small grids: histogram, scan and cursor scatter in ONE CTA ----
// Up to kSmallBins bins and kSmallN particles (BASELINE config 1: 10 000 particles, 25 x 25 cells) the step is bound by
// launch latency, not by work: the five launches of K_BIN's consumers (3 scan + scatter) become one.  The histogram lives
// in shared memory and is rebuilt from the cached bin words, so the force pass does not count in this mode.
constexpr int kSmallThreads = 1024;

__global__ void __launch_bounds__(kSmallThreads) small_sort(const int32_t *__restrict__ cell, int n_phys, Grid g, int nbins,
                                                            int32_t *__restrict__ cell_end, int32_t *__restrict__ perm,
                                                            volatile int *__restrict__ max_occ)
{
    extern __shared__ int s_cnt[]; // [nbins]
    __shared__ int sm[33];
    __shared__ int s_max;
    const int tid = threadIdx.x;
    if (tid == 0) s_max = 0;
    for (int b = tid; b < nbins; b += kSmallThreads) s_cnt[b] = 0;
    __syncthreads();
    for (int i = tid; i < n_phys; i += kSmallThreads) {
        const int c = container_of(__ldg(&cell[i]), g);
        if (c >= 0) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    // exclusive scan: thread t owns the bins [t*B, (t+1)*B)
    const int B = (nbins + kSmallThreads - 1) / kSmallThreads;
    const int b0 = tid * B, b1 = min(b0 + B, nbins);
    int local = 0, mx = 0;
    for (int b = b0; b < b1; ++b) {
        local += s_cnt[b];
        mx = max(mx, s_cnt[b]);
    }
    // fullest bin of this step -> host (mapped pinned memory, read without synchronising: it picks the force kernel of
    // the following steps - one warp per cell only pays while no cell is much fuller than the average)
    mx = __reduce_max_sync(0xffffffffu, mx);
    if ((tid & 31) == 0) atomicMax(&s_max, mx);
    int total;
    int ex = block_exclusive_scan(local, total, sm);
    for (int b = b0; b < b1; ++b) {
        const int c = s_cnt[b];
        s_cnt[b] = ex; // cursor = start of the bin
        ex += c;
    }
    if (tid == 0) {
        cell_end[-1] = 0;
        if (max_occ) *max_occ = s_max; // (block_exclusive_scan synchronised since the atomicMax above)
    }
    __syncthreads();
    for (int i = tid; i < n_phys; i += kSmallThreads) { // B/Physics.java:343-348 (order inside a bin is fixed by K_GATHER)
        const int c = container_of(__ldg(&cell[i]), g);
        if (c >= 0) perm[atomicAdd(&s_cnt[c], 1)] = i;
    }
    __syncthreads();
    for (int b = tid; b < nbins; b += kSmallThreads) cell_end[b] = s_cnt[b]; // END offsets, like `containers`
}

// B/Physics.java:343-348: `i = containers[ci]; buffer[i] = p; containers[ci]++`.
// Afterwards cell_end[b] is the END offset of bin b, as in the reference (per cell: every K-th entry).
constexpr int kScatterUnroll = 4; // measured: 1 -> 0.078 ms, 4 -> 0.056 ms at 16M particles

template <bool AGG>
__global__ void __launch_bounds__(kThreads) scatter_perm(const int32_t *__restrict__ cell, DevInt n_phys_, Grid g, int first,
                                                         int32_t *__restrict__ cell_end, int32_t *__restrict__ perm)
{
    // Several slots per thread: the kernel is load -> atomic round trip -> store, so what counts is how many atomics are
    // in flight.  All loads first, then all atomics, then the stores.
    constexpr int U = kScatterUnroll;
    const int n_phys = n_phys_.get();
    const int i0 = blockIdx.x * (kThreads * U) + threadIdx.x;
    if (blockIdx.x * (kThreads * U) >= n_phys) return;
    int c[U], slot[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int i = i0 + u * kThreads;
        c[u] = i < n_phys ? container_of(__ldg(&cell[i]), g) : -1; // -1: dead slot (the particle migrated to another slab)
    }
    if (!AGG) { // sparse grids: every lane has its own bin, aggregation only costs
#pragma unroll
        for (int u = 0; u < U; ++u) slot[u] = c[u] >= 0 ? atomicAdd(&cell_end[c[u]], 1) : -1;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (c[u] >= 0) perm[slot[u] - first] = i0 + u * kThreads;
        return;
    }
    // Warp-aggregated cursor: after the first step the array is almost cell-sorted, so the 32 lanes of a warp hit
    // few distinct bins; one atomic per distinct bin instead of one per particle, and neighbouring slots for
    // lanes of the same bin.  (Order inside a bin is still arbitrary across warps; K_GATHER ranks it.)
    const int lane = threadIdx.x & 31;
    unsigned peers[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        peers[u] = __match_any_sync(0xffffffffu, c[u]);
        slot[u] = 0;
        if (c[u] >= 0 && lane == __ffs(peers[u]) - 1) slot[u] = atomicAdd(&cell_end[c[u]], __popc(peers[u]));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int base = __shfl_sync(0xffffffffu, slot[u], __ffs(peers[u]) - 1);
        if (c[u] >= 0) perm[base + __popc(peers[u] & ((1u << lane) - 1u)) - first] = i0 + u * kThreads;
    }
}

constexpr int kGatherUnroll = 2; // measured: 1 -> 0.186 ms, 2 -> 0.166 ms, 4 -> 0.190 ms at 16M particles

// fp32: only the 16-byte candidate record moves.  Velocities stay where they are - each is needed once, by its own
// particle - and the force pass reads them through src_sorted.
//
// The record goes to its slot in the COMPUTE order: sorted by fine bin, stable inside a bin (deterministic, and the
// same on one GPU and on slabs).  The REFERENCE order (sorted by cell, stable inside a cell - exactly the permutation
// of B/Physics.java:343-348) is where the force pass writes its result; a particle's slot there is its rank among
// the pre-sort slots of its CELL, which the force pass computes from src_sorted (reference_slot, plife_internal.h).
// Bins nest in cells, so both slots lie in the cell's own index range; with ks = 0 they coincide.
// The record's type field is stored shifted (kTypeShift): it is the byte offset of a row of the force kernel's per-lane
// matrix table, so the staged key is the table offset itself.
// Block 0 also prepares the slab step (tr: target ranges of the two force launches; mig0/mig1: migration cursors).
template <bool STABLE>
__global__ void __launch_bounds__(kThreads) gather_f32(const float4 *__restrict__ pt_in, float4 *__restrict__ pt_out, DevInt n_, Grid g,
                                                       int first, StableKey key, const int32_t *__restrict__ cell,
                                                       int32_t *__restrict__ cell_sorted, int32_t *__restrict__ src_sorted,
                                                       const int32_t *__restrict__ cell_end, const int32_t *__restrict__ perm,
                                                       int *__restrict__ tr, float4 *__restrict__ mig0, float4 *__restrict__ mig1)
{
    // Two slots per thread: the kernel is a chain of four dependent memory round trips (perm -> record and cell ->
    // bin offsets -> keys of the bin), so its speed is the number of chains in flight; two per thread instead of one.
    constexpr int U = kGatherUnroll;
    const int n = n_.get();
    if (tr && blockIdx.x == 0 && threadIdx.x == 0) {
        // slab step: rows 1 and nly-2 are the first / last owned row; the force pass runs over the interior rows while the
        // halo is in flight, then over those two.  cell_end holds END offsets since the scatter.
        const int nxk = g.nxk();
        const int e1 = __ldg(cell_end + 2 * nxk - 1) - first;           // end of the first owned row
        const int sl = __ldg(cell_end + (g.nly - 2) * nxk - 1) - first; // start of the last owned row
        tr[0] = e1; tr[1] = sl; tr[2] = 0; tr[3] = 0;  // interior rows
        tr[4] = 0; tr[5] = e1; tr[6] = sl; tr[7] = n;  // first and last owned row
        tr[8] = 0; tr[9] = n; tr[10] = 0; tr[11] = 0;  // the whole sorted block (plife_get_step_stats)
        mig0[0] = make_float4(0.f, 0.f, 0.f, 0.f);     // this step's migration cursors
        mig1[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int d0 = blockIdx.x * (kThreads * U) + threadIdx.x;
    if (blockIdx.x * (kThreads * U) >= n) return;
    key.load();
    int src[U], cxy[U], c[U];
    float4 p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int d = d0 + u * kThreads;
        src[u] = d < n ? __ldg(&perm[d]) : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (src[u] >= 0) {
            p[u] = __ldg(&pt_in[src[u]]);
            cxy[u] = __ldg(&cell[src[u]]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = src[u] >= 0 ? container_of(cxy[u], g) : 0;
    const int32_t *pp = perm - first;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (src[u] < 0) continue;
        int dst = d0 + u * kThreads + first; // PLIFE_FLAG_UNSTABLE_SORT: the cursor order is the order
        if (STABLE) {
            const int bs = __ldg(&cell_end[c[u] - 1]), be = __ldg(&cell_end[c[u]]);
            dst = bs + stable_rank(pp, bs, be, src[u], first, key);
        }
        p[u].z = __int_as_float(__float_as_int(p[u].z) << kTypeShift);
        pt_out[dst] = p[u]; // sorted positions carry the ghost-below offset `first`; per-target arrays do not
        cell_sorted[dst - first] = cxy[u];
        src_sorted[dst - first] = src[u];
    }
}

template <bool STABLE>
__global__ void __launch_bounds__(kThreads) gather_f64(StateF64 in, StateF64 out, int n, Grid g, const int32_t *__restrict__ cell,
                                                       int32_t *__restrict__ cell_sorted, const int32_t *__restrict__ cell_end,
                                                       const int32_t *__restrict__ perm)
{
    int d = blockIdx.x * kThreads + threadIdx.x;
    if (d >= n) return;
    int src = __ldg(&perm[d]);
    double2 p = __ldg(&in.pos[src]);
    double2 v = __ldg(&in.vel[src]);
    int t = __ldg(&in.type[src]);
    uint32_t id = __ldg(&in.id[src]);
    int cxy = __ldg(&cell[src]);
    const int c = container_of(cxy, g);
    const int s = __ldg(&cell_end[c - 1]);
    int dst = d;
    if (STABLE) dst = s + stable_rank(perm, s, __ldg(&cell_end[c]), src, 0, StableKey{});
    cell_sorted[dst] = cxy;
    out.pos[dst] = p;
    out.vel[dst] = v;
    out.type[dst] = t;
    out.id[dst] = id;
}

// After plife_debug_neighbors on an fp32 handle the sort becomes the visible state, like makeContainers' swap
// (B/Physics.java:351-353): records go back to plain types at their reference slots, velocities follow them.
__global__ void __launch_bounds__(kThreads) apply_sort_f32(const float4 *__restrict__ pt_sorted, const float2 *__restrict__ vel_in,
                                                           const int32_t *__restrict__ src_sorted, const int32_t *__restrict__ cell_sorted,
                                                           const int32_t *__restrict__ cell_end, Grid g, int stable, int n,
                                                           float4 *__restrict__ pt_out, float2 *__restrict__ vel_out)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float4 p = __ldg(pt_sorted + i);
    p.z = __int_as_float(__float_as_int(p.z) >> kTypeShift);
    const int r = stable ? reference_slot(i, __ldg(cell_sorted + i), src_sorted, cell_end, g, 0, StableKey{}) : i;
    pt_out[r] = p;
    vel_out[r] = __ldg(vel_in + __ldg(src_sorted + i));
}

// END offset of every CELL = every K-th bin END offset (plife_get_containers)
__global__ void __launch_bounds__(kThreads) containers_from_bins(const int32_t *__restrict__ cell_end, int64_t ncell, int ks,
                                                                 int32_t *__restrict__ out)
{
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c < ncell) out[c] = __ldg(cell_end + ((c + 1) << ks) - 1);
}

__global__ void __launch_bounds__(kThreads) type_hist_f32(const float4 *__restrict__ pt, int n, int m,
                                                          unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int sh[256];
    for (int k = threadIdx.x; k < 256; k += kThreads) sh[k] = 0;
    __syncthreads();
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        int t = __float_as_int(__ldg(&pt[i]).z);
        if (t >= 0 && t < m) atomicAdd(&sh[t], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += kThreads)
        if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

__global__ void __launch_bounds__(kThreads) type_hist_f64(const int32_t *__restrict__ type, int n, int m,
                                                          unsigned long long *__restrict__ hist)
{
    __shared__ unsigned int sh[256];
    for (int k = threadIdx.x; k < 256; k += kThreads) sh[k] = 0;
    __syncthreads();
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        int t = __ldg(&type[i]);
        if (t >= 0 && t < m) atomicAdd(&sh[t], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += kThreads)
        if (sh[k]) atomicAdd(&hist[k], (unsigned long long)sh[k]);
}

// SplitMix64 counter stream, bit-identical to plife/synth.py
__device__ __forceinline__ double splitmix_u01(uint64_t seed, uint64_t counter)
{
    uint64_t z = seed + (counter + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

__global__ void __launch_bounds__(kThreads) init_uniform_f32(float4 *pt, float2 *vel, int n, int m, uint64_t seed)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    uint64_t c = 4ull * (uint64_t)i;
    double x = splitmix_u01(seed, c), y = splitmix_u01(seed, c + 1);
    int t = min((int)floor(splitmix_u01(seed, c + 2) * m), m - 1);
    pt[i] = make_float4((float)x, (float)y, __int_as_float(t), __uint_as_float((uint32_t)i));
    vel[i] = make_float2(0.f, 0.f);
}

// slab mode: every rank scans the global stream and keeps the particles of its own rows
__global__ void __launch_bounds__(kThreads) init_uniform_owned_f32(float4 *pt, float2 *vel, long long n_global, int m,
                                                                   uint64_t seed, Grid g, int cap, int *counter)
{
    long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (i >= n_global) return;
    uint64_t c = 4ull * (uint64_t)i;
    float x = (float)splitmix_u01(seed, c), y = (float)splitmix_u01(seed, c + 1);
    if (container_of(cell_coords((double)x, (double)y, g), g) < 0) return;
    int t = min((int)floor(splitmix_u01(seed, c + 2) * m), m - 1);
    int k = atomicAdd(counter, 1);
    if (k >= cap) return;
    pt[k] = make_float4(x, y, __int_as_float(t), __uint_as_float((uint32_t)i));
    vel[k] = make_float2(0.f, 0.f);
}

__global__ void __launch_bounds__(kThreads) init_uniform_f64(StateF64 s, int n, int m, uint64_t seed)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    uint64_t c = 4ull * (uint64_t)i;
    s.pos[i] = make_double2(splitmix_u01(seed, c), splitmix_u01(seed, c + 1));
    s.vel[i] = make_double2(0.0, 0.0);
    s.type[i] = min((int)floor(splitmix_u01(seed, c + 2) * m), m - 1);
    s.id[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kThreads) snapshot_from_f32(const float4 *__restrict__ pt, const float2 *__restrict__ vel,
                                                              int n, float2 *pos_out, float2 *vel_out, int32_t *type_out,
                                                              uint8_t *type8_out)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    float4 p = __ldg(&pt[i]);
    if (pos_out) pos_out[i] = make_float2(p.x, p.y);
    if (vel_out) vel_out[i] = __ldg(&vel[i]);
    if (type_out) type_out[i] = __float_as_int(p.z);
    if (type8_out) type8_out[i] = (uint8_t)__float_as_int(p.z); // m <= 256
}

__global__ void __launch_bounds__(kThreads) snapshot_from_f64(StateF64 s, int n, float2 *pos_out, float2 *vel_out,
                                                              int32_t *type_out, uint8_t *type8_out)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    if (pos_out) {
        double2 p = s.pos[i];
        pos_out[i] = make_float2((float)p.x, (float)p.y);
    }
    if (vel_out) {
        double2 v = s.vel[i];
        vel_out[i] = make_float2((float)v.x, (float)v.y);
    }
    if (type_out) type_out[i] = s.type[i];
    if (type8_out) type8_out[i] = (uint8_t)s.type[i];
}

// Slab mode: the array holds dead slots (particles that migrated away) until the next cell-list build; the
// snapshot skips them, keeping the order: per-block live counts -> scan -> compacting write.
__global__ void __launch_bounds__(kThreads) live_counts(const int32_t *__restrict__ cell, int n_phys, int *block_counts)
{
    int i = blockIdx.x * kThreads + threadIdx.x;
    int live = (i < n_phys && __ldg(&cell[i]) >= 0) ? 1 : 0;
    int cnt = __syncthreads_count(live);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(1024) scan_block_counts(int *block_counts, int nblocks)
{
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        int i = base + threadIdx.x;
        int v = i < nblocks ? block_counts[i] : 0;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[w] = inc;
        __syncthreads();
        if (w == 0) {
            int ws = warp_sums[lane], winc = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sums[lane] = winc - ws;
        }
        __syncthreads();
        int ex = carry_s + warp_sums[w] + inc - v;
        if (i < nblocks) block_counts[i] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = ex + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) snapshot_live_f32(const float4 *__restrict__ pt, const float2 *__restrict__ vel,
                                                              const int32_t *__restrict__ cell, int n_phys,
                                                              const int *__restrict__ block_offsets, float2 *pos_out,
                                                              float2 *vel_out, int32_t *type_out, uint8_t *type8_out)
{
    __shared__ int warp_sums[kThreads / 32];
    int i = blockIdx.x * kThreads + threadIdx.x;
    int live = (i < n_phys && __ldg(&cell[i]) >= 0) ? 1 : 0;
    unsigned m = __ballot_sync(0xffffffffu, live);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_sums[w] = __popc(m);
    __syncthreads();
    if (!live) return;
    int off = block_offsets[blockIdx.x];
    for (int k = 0; k < w; k++) off += warp_sums[k];
    off += __popc(m & ((1u << lane) - 1u));
    float4 p = __ldg(&pt[i]);
    if (pos_out) pos_out[off] = make_float2(p.x, p.y);
    if (vel_out) vel_out[off] = __ldg(&vel[i]);
    if (type_out) type_out[off] = __float_as_int(p.z);
    if (type8_out) type8_out[off] = (uint8_t)__float_as_int(p.z);
}

inline int blocks_for(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }
inline int first_index(const plife_handle *h) { return h->slab.on ? (int)h->slab.halo_cap : 0; }

} // namespace

cudaError_t launch_bin(plife_handle *h, const Grid &g)
{
    int n = (int)h->n_phys;
    if (n == 0) return cudaSuccess;
    if (h->precision == PLIFE_F32)
        bin_f32<<<blocks_for(n, kThreads), kThreads, 0, h->stream>>>(h->s32[h->cur].pt, n, g, h->d_cell, h->d_count);
    else
        bin_f64<<<blocks_for(n, kThreads), kThreads, 0, h->stream>>>(h->s64[h->cur].pos, n, g, h->d_cell, h->d_count);
    return cudaGetLastError();
}

cudaError_t launch_small_sort(plife_handle *h, const Grid &g)
{
    const int nbins = g.nxk() * g.nly;
    small_sort<<<1, kSmallThreads, sizeof(int) * nbins, h->stream>>>(h->d_cell, (int)h->n_phys, g, nbins, h->d_cell_end, h->d_perm, h->h_maxocc);
    return cudaGetLastError();
}

cudaError_t launch_scan(plife_handle *h, const Grid &g)
{
    int64_t ncell = (int64_t)g.nxk() * g.nly;
    int ntiles = (int)((ncell + kScanTile - 1) / kScanTile);
    if (!(h->flags & PLIFE_FLAG_SCAN3)) { // one launch (decoupled look-back); d_tile_sums = ScanState + one status word per tile
        ScanState *st = reinterpret_cast<ScanState *>(h->d_tile_sums);
        volatile unsigned long long *status = reinterpret_cast<unsigned long long *>(st + 1);
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
        static const int per_sm_env = getenv("PLIFE_SCAN_CTAS") ? atoi(getenv("PLIFE_SCAN_CTAS")) : 0;
        int nctas = nsm * (per_sm_env > 0 && per_sm_env <= 8 ? per_sm_env : kScanCtasPerSm); // all resident: the CTAs wait for each other
        if (nctas > ntiles) nctas = ntiles;
        const int tpc = (ntiles + nctas - 1) / nctas;
        nctas = (ntiles + tpc - 1) / tpc;
        scan_onepass<<<nctas, kScanThreads, 0, h->stream>>>(h->d_count, ncell, ntiles, tpc, first_index(h), h->d_cell_end, st, status);
        return cudaGetLastError();
    }
    int32_t *ts = reinterpret_cast<int32_t *>(h->d_tile_sums);
    scan_tile_sums<<<ntiles, kScanThreads, 0, h->stream>>>(h->d_count, ncell, ts);
    scan_sums<<<1, 1024, 0, h->stream>>>(ts, ntiles, first_index(h), h->d_cell_end);
    scan_apply<<<ntiles, kScanThreads, 0, h->stream>>>(h->d_count, ncell, ts, h->d_cell_end);
    return cudaGetLastError();
}

// Slab mode with device-resident counts: the host sizes the grids by an upper bound (h->n_bound) and the kernels read
// the true counts from SlabCounts; CTAs beyond the count exit at once.
cudaError_t launch_scatter(plife_handle *h, const Grid &g)
{
    const bool dev = h->slab.on && h->slab.counts;
    const int n = dev ? (int)h->slab.n_bound : (int)h->n_phys; // physical pre-sort length (dead slots included)
    if (n == 0) return cudaSuccess;
    const DevInt np{(int)h->n_phys, dev ? &h->slab.cnt()->n_phys : nullptr};
    const double rho = (double)h->n / ((double)g.nxk() * (g.row_hi - g.row_lo)); // particles per bin
    if (rho >= 4.0) scatter_perm<true><<<blocks_for(n, kThreads * kScatterUnroll), kThreads, 0, h->stream>>>(h->d_cell, np, g, first_index(h), h->d_cell_end, h->d_perm);
    else scatter_perm<false><<<blocks_for(n, kThreads * kScatterUnroll), kThreads, 0, h->stream>>>(h->d_cell, np, g, first_index(h), h->d_cell_end, h->d_perm);
    return cudaGetLastError();
}

cudaError_t launch_gather(plife_handle *h, const Grid &g)
{
    const bool dev = h->slab.on && h->slab.counts;
    const int n = dev ? (int)h->slab.n_bound : (int)h->n;
    if (n == 0) return cudaSuccess;
    int a = h->cur, b = h->cur ^ 1;
    bool stable = !(h->flags & PLIFE_FLAG_UNSTABLE_SORT);
    int nb = h->precision == PLIFE_F32 ? blocks_for(n, kThreads * kGatherUnroll) : blocks_for(n, kThreads);
    if (h->precision == PLIFE_F32) {
        const DevInt nn{(int)h->n, dev ? &h->slab.cnt()->n : nullptr};
        const StableKey key = stable_key_of(h); // slabs: arrivals are ordered by their previous global position
        int *tr = h->slab.on ? h->slab.d_tr : nullptr;
        if (stable)
            gather_f32<true><<<nb, kThreads, 0, h->stream>>>(h->s32[a].pt, h->s32[b].pt, nn, g, first_index(h), key, h->d_cell, h->d_cell_sorted,
                                                             h->d_src_sorted, h->d_cell_end, h->d_perm, tr, h->slab.mig_send[0], h->slab.mig_send[1]);
        else
            gather_f32<false><<<nb, kThreads, 0, h->stream>>>(h->s32[a].pt, h->s32[b].pt, nn, g, first_index(h), key, h->d_cell, h->d_cell_sorted,
                                                              h->d_src_sorted, h->d_cell_end, h->d_perm, tr, h->slab.mig_send[0], h->slab.mig_send[1]);
    } else {
        if (stable)
            gather_f64<true><<<nb, kThreads, 0, h->stream>>>(h->s64[a], h->s64[b], n, g, h->d_cell, h->d_cell_sorted, h->d_cell_end, h->d_perm);
        else
            gather_f64<false><<<nb, kThreads, 0, h->stream>>>(h->s64[a], h->s64[b], n, g, h->d_cell, h->d_cell_sorted, h->d_cell_end, h->d_perm);
    }
    return cudaGetLastError();
}

// fp32: make the sorted scratch (compute order, shifted types) the visible state in reference order
StableKey stable_key_of(const plife_handle *h)
{
    StableKey key;
    if (h->slab.on) {
        const SlabState &S = h->slab;
        const bool wrap = h->settings.wrap != 0 && S.world > 1;
        key.cnt = S.cnt();
        if (wrap && S.rank == 0 && S.rank != S.world - 1) key.order = 1;  // residents, above, below(seam)
        else if (wrap && S.rank == S.world - 1) key.order = 2;            // above(seam), below, residents
    }
    return key;
}

cudaError_t launch_apply_sort_f32(plife_handle *h, const Grid &g)
{
    const int n = (int)h->n;
    if (n == 0) return cudaSuccess;
    const int a = h->cur, b = h->cur ^ 1;
    apply_sort_f32<<<blocks_for(n, kThreads), kThreads, 0, h->stream>>>(h->s32[b].pt, h->s32[a].vel, h->d_src_sorted, h->d_cell_sorted, h->d_cell_end,
                                                                      g, (h->flags & PLIFE_FLAG_UNSTABLE_SORT) ? 0 : 1, n, h->s32[a].pt,
                                                                      h->s32[b].vel);
    return cudaGetLastError();
}

cudaError_t launch_containers(plife_handle *h, const Grid &g, int32_t *d_out)
{
    const int64_t ncell = (int64_t)g.nx * g.nly;
    containers_from_bins<<<blocks_for(ncell, kThreads), kThreads, 0, h->stream>>>(h->d_cell_end, ncell, g.ks, d_out);
    return cudaGetLastError();
}

cudaError_t launch_type_histogram(plife_handle *h, unsigned long long *d_hist)
{
    int n = (int)h->n;
    if (n == 0) return cudaSuccess;
    int nb = blocks_for(n, kThreads);
    if (nb > 148 * 8) nb = 148 * 8;
    if (h->precision == PLIFE_F32)
        type_hist_f32<<<nb, kThreads, 0, h->stream>>>(h->s32[h->cur].pt, n, h->m, d_hist);
    else
        type_hist_f64<<<nb, kThreads, 0, h->stream>>>(h->s64[h->cur].type, n, h->m, d_hist);
    return cudaGetLastError();
}

cudaError_t launch_init_uniform(plife_handle *h, int64_t n64, uint64_t seed)
{
    int n = (int)n64;
    if (n == 0) return cudaSuccess;
    if (h->precision == PLIFE_F32)
        init_uniform_f32<<<blocks_for(n, kThreads), kThreads, 0, h->stream>>>(h->s32[h->cur].pt, h->s32[h->cur].vel, n, h->m, seed);
    else
        init_uniform_f64<<<blocks_for(n, kThreads), kThreads, 0, h->stream>>>(h->s64[h->cur], n, h->m, seed);
    return cudaGetLastError();
}

cudaError_t launch_init_uniform_owned(plife_handle *h, int64_t n_global, uint64_t seed, const Grid &g, int *d_counter)
{
    if (n_global == 0) return cudaSuccess;
    init_uniform_owned_f32<<<blocks_for(n_global, kThreads), kThreads, 0, h->stream>>>(
        h->s32[h->cur].pt, h->s32[h->cur].vel, (long long)n_global, h->m, seed, g, (int)h->cap, d_counter);
    return cudaGetLastError();
}

cudaError_t launch_snapshot_f32(plife_handle *h, float2 *pos, float2 *vel, int32_t *type, uint8_t *type8)
{
    int n = (int)h->n;
    if (n == 0) return cudaSuccess;
    if (h->slab.on && h->n_phys != h->n) { // dead slots present: compact while copying (fp32 handles only)
        const int np = (int)h->n_phys, nb = blocks_for(np, kThreads);
        int *d_blocks = h->d_perm; // scratch: not live between steps
        live_counts<<<nb, kThreads, 0, h->stream>>>(h->d_cell, np, d_blocks);
        scan_block_counts<<<1, 1024, 0, h->stream>>>(d_blocks, nb);
        snapshot_live_f32<<<nb, kThreads, 0, h->stream>>>(h->s32[h->cur].pt, h->s32[h->cur].vel, h->d_cell, np, d_blocks, pos, vel, type, type8);
        return cudaGetLastError();
    }
    if (h->precision == PLIFE_F32)
        snapshot_from_f32<<<blocks_for(n, kThreads), kThreads, 0, h->stream>>>(h->s32[h->cur].pt, h->s32[h->cur].vel, n, pos, vel, type, type8);
    else
        snapshot_from_f64<<<blocks_for(n, kThreads), kThreads, 0, h->stream>>>(h->s64[h->cur], n, pos, vel, type, type8);
    return cudaGetLastError();
}

} // namespace plife
