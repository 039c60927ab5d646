// Internal declarations shared by the translation units of libplife.so.
// Not part of the ABI (see include/plife.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/plife.h"

namespace plife {

// fp32 compute-order records carry their type shifted left by kTypeShift: type << 9 is the byte offset of row `type` of
// the force kernel's per-lane matrix table (128 threads x 4 bytes per row), so staging candidates is a plain copy.
constexpr int kTypeShift = 9;

// A count the host may not know when it queues a kernel: the host value, or (slab mode) a device-resident one.
struct DevInt {
    int v;
    const int *dev;
    __device__ __forceinline__ int get() const { return dev ? *dev : v; }
};

// Slab mode: particle counts of this rank, kept in device memory so that a step needs no host round trip
// (slab.cu: finish kernel).  The pre-sort array is [residents 0..n_old) | arrivals from below | arrivals from above].
struct SlabCounts {
    int n;        // live particles
    int n_phys;   // physical length of the pre-sort array (dead slots included)
    int n_old, k_below, k_above;
    int err;      // error bits seen so far (copy of the sticky device word, for the host's ring)
    int sent_dn, sent_up;
    unsigned long long seq; // last finished step
};

// Uniform grid of the reference: cell edge = rmax, nx = ny = floor(1/rmax)
// (B/Physics.java:82-85, :312-313).  cs stays double: cell assignment is done
// with IEEE double division in every precision mode.
//
// Fine bins (fp32 handles at high density): every cell is cut into K = 1 << ks bins along x.  The counting sort is
// keyed by the fine bin, so the working ("compute") order of the force pass is (row, cell, bin) and a target only
// has to scan the bins within rmax of its own - (2 + 1/K) cell widths per row instead of 3 (-29 % candidates at
// K = 8).  Fine bins nest in cells (bin >> ks == cell), so a cell is still ONE contiguous index range, cell END
// offsets are every K-th entry of the bin END offsets, and the reference's particle order (stable by cell,
// B/Physics.java:343-348) is restored when the force pass writes its results (cells.cu: ref slot).  ks = 0 is the
// plain cell list.
struct Grid {
    int nx, ny;  // global grid
    double cs;   // cell edge = rmax
    double inv_cs; // fl(1 / cs): only for cell_coords_fast, which falls back to the division whenever the two could disagree
    // Slab decomposition over grid rows.  This rank owns global rows [row_lo, row_hi);
    // its cell arrays are indexed by LOCAL cells lx + ly*nx with ly = cy + ly_shift, where local row 0
    // and nly-1 are the ghost rows below / above.  Single-GPU: row_lo = 0, row_hi = nly = ny, ly_shift = 0.
    int row_lo, row_hi;
    int ly_shift, nly;
    int rows_up, rows_dn; // rows owned by the ring neighbours (migration reach check)
    int ks;               // log2 of the fine bins per cell along x
    int staged;           // fp32: the staged force kernel serves this grid (host decision, make_grid); else the v1 kernel, ks == 0
    float rho;            // particles per cell the decisions above were made for (slab mode: the same estimate on every rank)
    __host__ __device__ int nxk() const { return nx << ks; } // fine bins per row
};

// Bin word of a position: fx | cy << 16 with fx = (int)(x / containerSize * K) the un-clamped fine x index (so
// fx >> ks = (int)(x / containerSize) is both the un-clamped cx0 of the force pass, B/Physics.java:404 - floor ==
// truncation for x >= 0 - and, after the `== nx -> nx-1` clamp, the container of the sort, :362-375) and cy the
// un-clamped row.  The scaling by K = 2^ks is exact, so fine bins nest in cells bit for bit.  (nx + 1) << ks
// <= 65535 and ny <= 16384 (checked by make_grid), so both fit.  A negative word marks a dead slot (a particle
// that migrated away).
__host__ __device__ inline int cell_coords(double x, double y, const Grid &g)
{
    int fx = (int)((x / g.cs) * (double)(1 << g.ks));
    int cy = (int)(y / g.cs);
    return fx | (cy << 16);
}
// The same word without the two IEEE double divisions (~30 instructions each; at one particle per cell they were a fifth of
// the force kernel's instructions).  fl(x / cs) and x * fl(1 / cs) differ by a few ulps at most, so their scaled floors can
// only differ when the product lies within a few ulps of an integer: exactly then (practically never) the division is done.
__device__ __forceinline__ int fine_index_fast(double x, const Grid &g)
{
    const double a = x * g.inv_cs * (double)(1 << g.ks);
    const double d = a * 1.8e-15; // 8 ulps, relative
    const int lo = (int)(a - d), hi = (int)(a + d);
    if (lo == hi) return lo;
    return (int)((x / g.cs) * (double)(1 << g.ks));
}
__device__ __forceinline__ int cell_coords_fast(double x, double y, const Grid &g)
{
    const int fx = fine_index_fast(x, g);
    // rows are not scaled: the same test with K = 1
    const double a = y * g.inv_cs, d = a * 1.8e-15;
    const int lo = (int)(a - d), hi = (int)(a + d);
    const int cy = lo == hi ? lo : (int)(y / g.cs);
    return fx | (cy << 16);
}
// local fine-bin index of a bin word, or -1 if the particle is dead / its row is not owned by this rank
__host__ __device__ inline int container_of(int cxy, const Grid &g)
{
    if (cxy < 0) return -1;
    int fx = cxy & 0xffff, cy = cxy >> 16;
    if (fx >= g.nxk()) fx = g.nxk() - 1; // cx == nx -> nx-1 (for solid borders, :367-372): the last bin of the last cell
    if (cy == g.ny) cy = g.ny - 1;
    if (cy < g.row_lo || cy >= g.row_hi) return -1;
    return fx + (cy + g.ly_shift) * g.nxk();
}
// local row of a (wrapped) global row reached from an owned row's 3x3 neighbourhood
__host__ __device__ inline int local_row(int cy, const Grid &g)
{
    int ly = cy + g.ly_shift;
    if (ly < 0) ly += g.ny;
    else if (ly >= g.ny) ly -= g.ny;
    return ly;
}
// Slab mode: a particle sitting exactly on y == 1.0 (Range.wrap can return it), or in the strip above
// ny*rmax, has the un-clamped row ny; the reference scans the rows ny-1, 0, 1 around it (B/Physics.java:404-417), but global
// row 1 lives two slabs away from the last slab.  Rows ny-2, ny-1, 0 hold every particle within rmax of it (row 1 starts a
// full rmax above y == 0 == 1), so the last slab scans those instead: the same forces bit for bit (out-of-range candidates
// add exact zeros), only the candidate COUNT of such a particle differs from the reference's.
__host__ __device__ inline int scan_row(int cy0, const Grid &g) { return (g.nly != g.ny && cy0 >= g.ny) ? g.ny - 1 : cy0; }

// [start, end) of the local CELLS c_lo .. c_hi (inclusive, same row) in the sorted array: bin END offsets, cell_end[-1] valid
__device__ __forceinline__ void cell_span(const int32_t *__restrict__ cell_end, int c_lo, int c_hi, int ks, int &s, int &e)
{
    s = __ldg(cell_end + (c_lo << ks) - 1);
    e = __ldg(cell_end + ((c_hi + 1) << ks) - 1);
}

// Logical position of pre-sort slot `src`.  Single GPU (cnt == nullptr): the identity.  Slab mode: the pre-sort array is
// physically [residents 0..n_old) | arrivals from below (k_below) | arrivals from above]; the key puts the
// three groups in the order of their previous GLOBAL array index, which is what the reference's stable
// sort preserves: below < residents < above, except across the periodic seam, where rank 0's arrivals
// from "below" come from the END of the global array and the last rank's arrivals from "above" from its
// beginning.  The counts live in device memory (SlabCounts): the host does not know them when it queues the step.
struct StableKey {
    const SlabCounts *cnt = nullptr;
    int order = 0; // 0: below, residents, above   1 (rank 0, periodic): residents, above, below   2 (last rank, periodic): above, below, residents
    int n_old = 0x7fffffff, k_below = 0, base_res = 0, base_below = 0, base_above = 0;
    __device__ __forceinline__ void load()
    {
        if (!cnt) return;
        n_old = cnt->n_old;
        k_below = cnt->k_below;
        const int ka = cnt->k_above;
        if (order == 1) { base_res = 0; base_below = n_old + ka; base_above = n_old; }
        else if (order == 2) { base_res = ka + k_below; base_below = ka; base_above = 0; }
        else { base_res = k_below; base_below = 0; base_above = k_below + n_old; }
    }
    __device__ __forceinline__ int operator()(int src) const
    {
        if (src < n_old) return src + base_res;
        src -= n_old;
        return src < k_below ? src + base_below : src - k_below + base_above;
    }
    __device__ __forceinline__ bool mixed(int max_src) const { return max_src >= n_old; } // the range holds an arrival
};

// number of entries of pp[s, e) (pre-sort slots of one cell or bin) that precede `src` in the previous array order;
// pp + k is 16-byte aligned where (k - align) % 4 == 0
__device__ __forceinline__ int stable_rank(const int32_t *__restrict__ pp, int s, int e, int src, int align, const StableKey &key)
{
    // Raw pre-sort indices order the residents exactly like their keys do, and almost every cell holds residents
    // only, so rank on the raw indices (four per load: cells of an evolved state hold thousands of particles and
    // this loop is O(count^2) per cell) and track the largest one; only cells that received migrants (slab mode)
    // are ranked again through the key.
    int rank = 0;
    int mx = src;
    int k = s;
    for (; k < e && ((k - align) & 3); ++k) {
        const int q = __ldg(pp + k);
        rank += (q < src) ? 1 : 0;
        mx = max(mx, q);
    }
    for (; k + 4 <= e; k += 4) {
        const int4 q = __ldg(reinterpret_cast<const int4 *>(pp + k));
        rank += ((q.x < src) ? 1 : 0) + ((q.y < src) ? 1 : 0) + ((q.z < src) ? 1 : 0) + ((q.w < src) ? 1 : 0);
        mx = max(max(mx, q.x), max(q.y, max(q.z, q.w)));
    }
    for (; k < e; ++k) {
        const int q = __ldg(pp + k);
        rank += (q < src) ? 1 : 0;
        mx = max(mx, q);
    }
    if (key.mixed(mx)) {
        rank = 0;
        const int ksrc = key(src);
        for (k = s; k < e; ++k) rank += (key(__ldg(pp + k)) < ksrc) ? 1 : 0;
    }
    return rank;
}

// Slot of sorted (compute-order) particle i in the REFERENCE order: its cell's start + its rank among the pre-sort slots
// of the cell (B/Physics.java:343-348 is a stable scatter: inside a cell the previous array order is kept).  Per-target
// numbering (no ghost-row offset); `key` must be loaded.  With ks == 0 the compute order is the reference order.
__device__ __forceinline__ int reference_slot(int i, int cxy, const int32_t *__restrict__ src_sorted, const int32_t *__restrict__ cell_end,
                                              const Grid &g, int first, const StableKey &key)
{
    if (g.ks == 0) return i;
    const int c0 = (container_of(cxy, g) >> g.ks) << g.ks; // first bin of the cell (bins per row is a multiple of K)
    const int cs = __ldg(cell_end + c0 - 1), ce = __ldg(cell_end + c0 + (1 << g.ks) - 1);
    if (!key.cnt) { // one GPU: no arrivals, the raw pre-sort indices are the keys and there is no maximum to track (stable_rank
                    // does, to notice arrivals; a slab cannot skip it even for its interior rows: a fast particle may land there)
        const int32_t *__restrict__ pp = src_sorted - first; // pp + k is 16-byte aligned where (k - first) % 4 == 0
        const int src = __ldg(src_sorted + i);
        int rank = 0, k = cs;
        for (; k < ce && ((k - first) & 3); ++k) rank += (__ldg(pp + k) < src) ? 1 : 0;
        for (; k + 4 <= ce; k += 4) {
            const int4 q = __ldg(reinterpret_cast<const int4 *>(pp + k));
            rank += ((q.x < src) ? 1 : 0) + ((q.y < src) ? 1 : 0) + ((q.z < src) ? 1 : 0) + ((q.w < src) ? 1 : 0);
        }
        for (; k < ce; ++k) rank += (__ldg(pp + k) < src) ? 1 : 0;
        return cs - first + rank;
    }
    return cs - first + stable_rank(src_sorted - first, cs, ce, __ldg(src_sorted + i), first, key);
}

// Kernel parameters of the force/integrate pass for one step, in the
// arithmetic type R of the handle.
template <typename R>
struct ForceParams {
    int n, m;
    const int *n_dev; // slab mode: device-resident particle count (overrides n), or nullptr
    const int *tr;    // slab mode: device-resident target ranges {s0, e0, s1, e1} of this launch, or nullptr: [0, n)
    int bin_lo, bin_hi; // bins a CTA may stage (the interior launch of a slab step must keep off the ghost rows)
    int first; // sorted-array index of target 0 (ghost-below capacity in slab mode, else 0)
    Grid g;
    int wrap;
    int use_smem_matrix;
    R rmax, r2, invr; // rmax, rmax*rmax, 1/rmax
    R mu;             // pow(friction, 60*dt), B/Physics.java:401
    R k2;             // (rmax*force)*dt,      B/Physics.java:437
    R dt;
    R accp[4];        // accelerator parameters
    // kind-0 constants in absolute distance units (fp32 fast path, force_impl.cuh)
    R fast_b, fast_d0, fast_h; // beta*rmax, (1+beta)*rmax/2, (1-beta)*rmax/2
    R fast_a_scale;            // 2*beta/(1-beta), folded into the matrix
    R fast_k;                  // k2/(beta*rmax) = force*dt/beta
};

// Device state of one precision.  F32 packs {x, y, type bits, id bits} into one
// float4 so that a neighbour candidate is a single 16-byte load.
struct StateF32 {
    float4 *pt;
    float2 *vel;
};
struct StateF64 {
    double2 *pos;
    double2 *vel;
    int32_t *type;
    uint32_t *id;
};

struct Timer {
    cudaEvent_t ev[PLIFE_K_COUNT + 1];
};

} // namespace plife

namespace plife {
// Slab-decomposition state of a handle; see slab.cu.
struct SlabState {
    bool on = false;
    int rank = 0, world = 1;
    int64_t halo_cap = 0, mig_cap = 0; // particles per halo row message / per migration message
    // exchange buffers (device pointers handed in by the host, 16-byte records); [0] = down, [1] = up
    float4 *halo_send[2]{}, *halo_recv[2]{}, *mig_send[2]{}, *mig_recv[2]{};
    int64_t n_old = 0, k_below = 0, k_above = 0; // host copy (lagging): layout of the pre-sort array: residents | from below | from above
    int phase = 0;                  // next phase expected by plife_slab_phase
    // peer exchange (library-owned buffers, CUDA IPC)
    bool peer_mode = false;
    float4 *xbuf = nullptr;         // flags + receive slots (exported)
    int64_t xrecords = 0, hrec = 0, mrec = 0;
    int nx_cfg = 0;
    unsigned long long seq = 1;     // step sequence number written into the neighbours' flags
    unsigned long long spin_ns = 30000000000ull; // device-side wait limit for a neighbour's message (PLIFE_SLAB_TIMEOUT_MS)
    float4 *peer_base[2]{};         // the neighbours' xbuf mapped into this process / device
    bool peer_ipc[2]{};
    // device-resident counts: a step needs no host synchronisation
    SlabCounts *counts = nullptr;        // device: TWO structs - a step reads counts[seq & 1], its finish kernel (which runs next to the
                                         // interior force launch) writes counts[(seq + 1) & 1]
    int *d_err = nullptr;                // device: sticky error bits (kErr*, slab.cu)
    SlabCounts *cnt() const { return counts ? counts + (seq & 1) : nullptr; }
    SlabCounts *nxt() const { return counts ? counts + ((seq + 1) & 1) : nullptr; }
    int *d_tr = nullptr;                 // device: target ranges of the force launches (the gather kernel writes them), 12 ints
    volatile unsigned long long *d_go = nullptr; // device: the finish CTA releases the append CTAs of the same launch
    volatile SlabCounts *h_ring = nullptr; // mapped pinned: the counts after each of the last 8 steps (written by slab_finish)
    cudaEvent_t step_done[4]{};          // recorded after each step's FINISH: bounds how far the host runs ahead
    cudaStream_t side = nullptr;         // peer mode: high-priority stream for the halo traffic and the edge rows
    cudaEvent_t ev_sorted = nullptr, ev_edge = nullptr;
    unsigned long long seq_known = 0;    // newest step whose counts the host has read
    int64_t n_bound = 0;                 // upper bound of n_phys the host sizes grids with
    int64_t max_arrivals = 0;            // largest number of arrivals seen in one step
    int err_pending = 0, err_seen = 0;   // device error bits read from the ring / already reported
};
} // namespace plife

struct plife_handle;
namespace plife {
void slab_destroy(plife_handle *h);
int slab_refresh(plife_handle *h, bool block); // host view of the device-resident counts (block: drain the stream first)
int slab_set_counts(plife_handle *h);          // host-known counts -> device (after upload / init)
}

struct plife_handle {
    int device = 0;
    int precision = PLIFE_F32;
    int flags = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool poisoned = false;

    plife_settings settings{0.02, 0.85, 1.0, 1, 0};
    int acc_kind = PLIFE_ACC_PARTICLE_LIFE;
    double acc_params[4] = {0.3, 0, 0, 0};

    int m = 1;
    std::vector<double> matrix{0.0}; // row-major [own][other]
    void *d_matrix_t = nullptr;      // transposed [other][own], in R
    int d_matrix_cap = 0;            // entries
    bool matrix_dirty = true;

    int64_t n = 0;      // live particles
    int64_t n_phys = 0; // physical length of the pre-sort array (== n except in slab mode: dead slots)
    int64_t n_sorted = 0; // particles in the sorted scratch (the n of the last cell-list build)
    int64_t cap = 0;
    int64_t capacity_hint = 0;
    plife::SlabState slab;
    int max_type = -1;
    int bins_override = -1; // PLIFE_BINS: log2 of the fine bins per cell, or -1: chosen from the density
    uint32_t next_id = 0; // id given to the next appended particle
    int cur = 0; // index of the buffer holding the current state
    plife::StateF32 s32[2]{};
    plife::StateF64 s64[2]{};
    int32_t *d_cell = nullptr;        // packed cell coords of particle i (pre-sort order)
    int32_t *d_cell_sorted = nullptr; // the same, permuted into sorted order
    int32_t *d_src_sorted = nullptr;  // fp32: pre-sort slot of every sorted particle (velocities are read through it)
    int32_t *d_perm = nullptr; // source index of sorted slot d
    void *d_snap = nullptr;    // snapshot staging (download_f32)
    int64_t snap_cap = 0;
    // asynchronous, double-buffered snapshot (display-time handoff that overlaps the next steps)
    void *d_snap_async[2]{};
    int64_t snap_async_cap = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t snap_ready[2]{}, snap_done[2]{};
    int64_t snap_issued = 0, snap_awaited = 0; // asynchronous snapshots requested / handed over by plife_snapshot_wait
    bool snap_init = false;

    int32_t *d_count = nullptr;   // per-cell histogram, zero between steps
    int32_t *d_cell_end = nullptr; // `containers`: END offset per cell
    void *d_tile_sums = nullptr; // u64 per scan tile
    int64_t cell_cap = 0;
    int64_t tile_cap = 0;
    unsigned long long *d_scalar = nullptr; // small device scratch (counters)
    unsigned long long *d_hist = nullptr;   // 256 type counters (inside the d_scalar allocation)

    plife::Grid last_grid{0, 0, 0.0};
    bool prebinned = false;   // d_cell / d_count already hold the binning of the current state (fused into the last force pass)
    bool small_step = false;       // the current / last step ran in small mode
    volatile int *h_maxocc = nullptr; // mapped pinned: particles in the fullest bin, written by small_sort every step (read lagging)
    bool prebinned_counts = false; // ... including the histogram d_count (not in small mode: small_sort recounts in shared memory)
    // launch-bound regime: the step replays a captured CUDA graph (one per velocity-buffer parity)
    struct GraphKey {
        double dt, rmax, friction, force, accp[4];
        int wrap, acc_kind, m, flags, ks, staged;
        long long n, matrix_version;
        const void *pt0, *pt1, *vel0, *vel1, *cell_end;
        int cur;
    };
    struct GraphSlot {
        GraphKey key;
        cudaGraphExec_t exec;
        bool valid;
    } graphs[2]{};
    long long matrix_version = 0;
    int stable_steps = 0; // consecutive steps with an unchanged key
    GraphKey last_key{};
    long long graph_launches = 0, graph_captures = 0;
    plife::Grid prebinned_grid{0, 0, 0.0};
    bool count_dirty = false; // d_count is not all-zero
    bool has_sorted = false; // buffer cur^1 holds the sorted pre-step state of the last step
    int64_t steps = 0;

    bool profiling = false;
    cudaEvent_t ev[PLIFE_K_COUNT + 1]{};
    bool ev_created = false;
    double k_ms[PLIFE_K_COUNT]{};
    int64_t k_launches[PLIFE_K_COUNT]{};
    struct PendingTiming {
        cudaEvent_t ev[PLIFE_K_COUNT + 1];
    };
    std::vector<PendingTiming> pending;
    PendingTiming slab_timing{}; // slab mode: one step's events span the SORT and FORCE phases
    bool slab_timing_on = false;

    std::atomic<int> stop_requested{0};
    std::string last_error;
};

namespace plife {

// cells.cu
cudaError_t launch_bin(plife_handle *h, const Grid &g);
cudaError_t launch_scan(plife_handle *h, const Grid &g);
cudaError_t launch_small_sort(plife_handle *h, const Grid &g); // histogram + scan + scatter in one CTA (small grids)
constexpr int kSmallBins = 8192;  // small_sort: the histogram must fit the CTA's shared memory
constexpr int kSmallN = 65536;
cudaError_t launch_scatter(plife_handle *h, const Grid &g);
cudaError_t launch_gather(plife_handle *h, const Grid &g);
cudaError_t launch_apply_sort_f32(plife_handle *h, const Grid &g);
StableKey stable_key_of(const plife_handle *h);
cudaError_t launch_containers(plife_handle *h, const Grid &g, int32_t *d_out);
cudaError_t launch_type_histogram(plife_handle *h, unsigned long long *d_hist);
cudaError_t launch_init_uniform(plife_handle *h, int64_t n, uint64_t seed);
cudaError_t launch_init_uniform_owned(plife_handle *h, int64_t n_global, uint64_t seed, const Grid &g, int *d_counter);
cudaError_t launch_snapshot_f32(plife_handle *h, float2 *pos, float2 *vel, int32_t *type, uint8_t *type8 = nullptr);

// force_f32.cu / force_f64.cu
cudaError_t launch_force_f32(plife_handle *h, const ForceParams<float> &p);
cudaError_t launch_force_f32_part(plife_handle *h, const ForceParams<float> &p, int nblocks, cudaStream_t stream, bool no_leavers = false); // slab mode: one of the two launches
void launch_force_f32_done(plife_handle *h);                                                  // ... then swap the velocity buffers
cudaError_t launch_force_f64(plife_handle *h, const ForceParams<double> &p);
cudaError_t launch_neighbors_f32(plife_handle *h, const ForceParams<float> &p, int32_t *cnt, unsigned long long *hash);
cudaError_t launch_neighbors_f64(plife_handle *h, const ForceParams<double> &p, int32_t *cnt, unsigned long long *hash);
cudaError_t launch_pair_count_f32(plife_handle *h, const ForceParams<float> &p, unsigned long long *d_total);
cudaError_t launch_pair_count_f64(plife_handle *h, const ForceParams<double> &p, unsigned long long *d_total);

} // namespace plife
