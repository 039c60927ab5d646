// C ABI of libplife.so (include/plife.h): handle management, settings, host<->device
// transfer and the per-step kernel sequence.  There is no CPU compute path in
// this library: every entry point that touches particles runs CUDA kernels and
// fails with PLIFE_ERR_CUDA when no device is usable.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "plife_internal.h"

using namespace plife;

namespace {

int fail(plife_handle *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->last_error = buf;
    return code;
}

int cuda_fail(plife_handle *h, cudaError_t e, const char *what)
{
    // sticky errors (illegal address, launch failure, ...) poison the handle
    if (e != cudaErrorMemoryAllocation && e != cudaErrorInvalidValue && e != cudaErrorNotReady) h->poisoned = true;
    cudaGetLastError();
    if (e == cudaErrorMemoryAllocation) return fail(h, PLIFE_ERR_OOM, "%s: %s", what, cudaGetErrorString(e));
    return fail(h, PLIFE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CU(h, expr)                                      \
    do {                                                 \
        cudaError_t e_ = (expr);                         \
        if (e_ != cudaSuccess) return cuda_fail(h, e_, #expr); \
    } while (0)

#define CHECK_HANDLE(h)                                                              \
    do {                                                                             \
        if (!(h)) return PLIFE_ERR_INVALID;                                          \
        if ((h)->poisoned) return fail(h, PLIFE_ERR_CUDA, "handle poisoned by an earlier CUDA error"); \
        cudaError_t e_ = cudaSetDevice((h)->device);                                 \
        if (e_ != cudaSuccess) return cuda_fail(h, e_, "cudaSetDevice");              \
    } while (0)

constexpr int64_t kMaxCells = (int64_t)1 << 28;
constexpr int kScanTile = 4096; // must match cells.cu

template <typename T>
cudaError_t dev_alloc(T **p, size_t count)
{
    return cudaMalloc((void **)p, sizeof(T) * (count ? count : 1));
}

void free_state(plife_handle *h)
{
    for (int b = 0; b < 2; b++) {
        cudaFree(h->s32[b].pt);
        cudaFree(h->s32[b].vel);
        cudaFree(h->s64[b].pos);
        cudaFree(h->s64[b].vel);
        cudaFree(h->s64[b].type);
        cudaFree(h->s64[b].id);
        h->s32[b] = StateF32{};
        h->s64[b] = StateF64{};
    }
    cudaFree(h->d_cell);
    cudaFree(h->d_cell_sorted);
    cudaFree(h->d_src_sorted);
    cudaFree(h->d_perm);
    h->d_cell = h->d_cell_sorted = h->d_src_sorted = h->d_perm = nullptr;
    h->cap = 0;
    h->prebinned = false;
}

// grow particle buffers; contents are NOT preserved
int ensure_capacity(plife_handle *h, int64_t n)
{
    if (n <= h->cap) return PLIFE_OK;
    if (n > 0x7fffffffLL - 1024) return fail(h, PLIFE_ERR_INVALID, "particle count %lld exceeds int32 indexing", (long long)n);
    free_state(h);
    size_t c = (size_t)n;
    for (int b = 0; b < 2; b++) {
        if (h->precision == PLIFE_F32) {
            // slab mode: the sorted array is [ghost row below | owned | ghost row above]
            CU(h, dev_alloc(&h->s32[b].pt, c + (h->slab.on ? 2 * (size_t)h->slab.halo_cap : 0)));
            CU(h, dev_alloc(&h->s32[b].vel, c));
        } else {
            CU(h, dev_alloc(&h->s64[b].pos, c));
            CU(h, dev_alloc(&h->s64[b].vel, c));
            CU(h, dev_alloc(&h->s64[b].type, c));
            CU(h, dev_alloc(&h->s64[b].id, c));
        }
    }
    CU(h, dev_alloc(&h->d_cell, c));
    CU(h, dev_alloc(&h->d_cell_sorted, c));
    CU(h, dev_alloc(&h->d_src_sorted, c));
    CU(h, dev_alloc(&h->d_perm, c));
    h->cap = n;
    return PLIFE_OK;
}

// grow particle buffers keeping the current state (plife_append)
int grow_preserve(plife_handle *h, int64_t cap)
{
    if (cap <= h->cap) return PLIFE_OK;
    if (cap > 0x7fffffffLL - 1024) return fail(h, PLIFE_ERR_INVALID, "particle count %lld exceeds int32 indexing", (long long)cap);
    CU(h, cudaStreamSynchronize(h->stream));
    const size_t c = (size_t)cap, n = (size_t)h->n_phys;
    const int cur = h->cur;
    if (h->precision == PLIFE_F32) {
        StateF32 nw[2]{};
        for (int b = 0; b < 2; b++) {
            CU(h, dev_alloc(&nw[b].pt, c));
            CU(h, dev_alloc(&nw[b].vel, c));
        }
        if (n) {
            CU(h, cudaMemcpy(nw[cur].pt, h->s32[cur].pt, sizeof(float4) * n, cudaMemcpyDeviceToDevice));
            CU(h, cudaMemcpy(nw[cur].vel, h->s32[cur].vel, sizeof(float2) * n, cudaMemcpyDeviceToDevice));
        }
        for (int b = 0; b < 2; b++) {
            cudaFree(h->s32[b].pt);
            cudaFree(h->s32[b].vel);
            h->s32[b] = nw[b];
        }
    } else {
        StateF64 nw[2]{};
        for (int b = 0; b < 2; b++) {
            CU(h, dev_alloc(&nw[b].pos, c));
            CU(h, dev_alloc(&nw[b].vel, c));
            CU(h, dev_alloc(&nw[b].type, c));
            CU(h, dev_alloc(&nw[b].id, c));
        }
        if (n) {
            CU(h, cudaMemcpy(nw[cur].pos, h->s64[cur].pos, sizeof(double2) * n, cudaMemcpyDeviceToDevice));
            CU(h, cudaMemcpy(nw[cur].vel, h->s64[cur].vel, sizeof(double2) * n, cudaMemcpyDeviceToDevice));
            CU(h, cudaMemcpy(nw[cur].type, h->s64[cur].type, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice));
            CU(h, cudaMemcpy(nw[cur].id, h->s64[cur].id, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice));
        }
        for (int b = 0; b < 2; b++) {
            cudaFree(h->s64[b].pos);
            cudaFree(h->s64[b].vel);
            cudaFree(h->s64[b].type);
            cudaFree(h->s64[b].id);
            h->s64[b] = nw[b];
        }
    }
    cudaFree(h->d_cell);
    cudaFree(h->d_cell_sorted);
    cudaFree(h->d_src_sorted);
    cudaFree(h->d_perm);
    h->d_cell = h->d_cell_sorted = h->d_src_sorted = h->d_perm = nullptr;
    CU(h, dev_alloc(&h->d_cell, c));
    CU(h, dev_alloc(&h->d_cell_sorted, c));
    CU(h, dev_alloc(&h->d_src_sorted, c));
    CU(h, dev_alloc(&h->d_perm, c));
    h->cap = cap;
    h->prebinned = false;
    h->has_sorted = false;
    return PLIFE_OK;
}


int ensure_cells(plife_handle *h, int64_t ncell)
{
    if (ncell <= h->cell_cap) return PLIFE_OK;
    cudaFree(h->d_count);
    if (h->d_cell_end) cudaFree(h->d_cell_end - 4);
    cudaFree(h->d_tile_sums);
    h->d_count = h->d_cell_end = nullptr;
    h->d_tile_sums = nullptr;
    h->cell_cap = 0;
    int64_t padded = (ncell + kScanTile - 1) / kScanTile * kScanTile;
    CU(h, dev_alloc(&h->d_count, (size_t)padded));
    // 4 leading ints: cell_end[-1] (the start of local cell 0) is a valid entry and int4 stores stay aligned
    int32_t *raw = nullptr;
    CU(h, dev_alloc(&raw, (size_t)padded + 4));
    CU(h, cudaMemsetAsync(raw, 0, 16, h->stream));
    h->d_cell_end = raw + 4;
    // scan scratch: ScanState (16 bytes) + one 64-bit status word per tile (the 3-launch variant uses it as int32 tile sums)
    CU(h, cudaMalloc(&h->d_tile_sums, 16 + sizeof(unsigned long long) * (size_t)(padded / kScanTile)));
    CU(h, cudaMemsetAsync(h->d_tile_sums, 0, 16 + sizeof(unsigned long long) * (size_t)(padded / kScanTile), h->stream));
    CU(h, cudaMemsetAsync(h->d_count, 0, sizeof(int32_t) * (size_t)padded, h->stream));
    h->cell_cap = padded;
    h->count_dirty = false;
    h->prebinned = false;
    return PLIFE_OK;
}

// Fine bins per cell along x (Grid::ks), fp32 handles only.  The gain is fewer candidates per target ((2 + 1/K) x 3
// cell areas instead of 9); the cost is K times more bins to count and scan.  Worth it from a few particles per bin
// on; the staged force kernel must be the one in use (m <= 32, >= 4 particles per cell).  PLIFE_BINS=K overrides.
// Slab mode: every rank must pick the same K (the halo messages carry per-bin offsets), so the density estimate
// comes from the halo capacity every rank was configured with, not from the rank's own particle count.
void choose_kernel(const plife_handle *h, Grid *g)
{
    g->ks = 0;
    g->staged = 0;
    double rho;
    if (h->slab.on) rho = (double)h->slab.halo_cap / (1.5 * g->nx);
    else rho = (double)h->n / ((double)g->nx * g->ny);
    g->rho = (float)rho;
    // staged kernel whenever the per-lane matrix table fits (m <= 32); below ~4 particles per cell the per-CTA staging and
    // table fill cost more than they save and the v1 kernel (global-memory walk, plain cell list) wins
    if (h->precision != PLIFE_F32 || (h->flags & PLIFE_FLAG_FORCE_V1) || h->m > 32 || rho < 4.0) return;
    // small particle counts (latency-bound: the machine is mostly empty): one warp per cell on the plain cell list
    // - while no cell is much fuller than the average (small_sort reports the fullest cell of every step; the value is read
    // without synchronising, a step or two late): in a clustered state a warp per cell is badly balanced and the staged
    // kernel, which streams dense ranges through shared memory in chunks, is the faster one again (crossover measured at
    // about 6 x the mean on an evolving 10 000-particle state; hysteresis so that the choice does not flip every step).
    if (!h->slab.on && h->n <= kSmallN && (int64_t)g->nx * g->ny <= kSmallBins && h->m <= 64 && rho <= 64.0 && !(h->flags & PLIFE_FLAG_NO_CELLS)) {
        // (the staged path counts per fine bin: K bins per cell, so scale to an upper bound of the cell's count)
        const int occ = h->h_maxocc ? (*h->h_maxocc << (h->last_grid.staged == 1 ? h->last_grid.ks : 0)) : 0;
        const double mean = rho > 4.0 ? rho : 4.0;
        const bool was_cells = h->last_grid.staged == 2;
        if (occ <= (was_cells ? 10.0 : 7.0) * mean) {
            g->staged = 2;
            return;
        }
    }
    g->staged = 1;
    int ks = rho >= 12.0 ? 3 : (rho >= 6.0 ? 2 : 1);
    if (h->bins_override >= 0) ks = h->bins_override;
    while (ks > 0 && (((int64_t)(g->nx + 1) << ks) > 65535 || (((int64_t)g->nx * g->nly) << ks) > kMaxCells)) ks--;
    g->ks = ks;
}

// B/Physics.java:82-85 with containerSize = rmax (:312)
int make_grid(plife_handle *h, Grid *g)
{
    double rmax = h->settings.rmax;
    int nx = (int)floor(1 / rmax);
    if (!(rmax > 0) || nx < 1) return fail(h, PLIFE_ERR_INVALID, "rmax=%g gives nx=%d", rmax, nx);
    if ((int64_t)nx * nx > kMaxCells) return fail(h, PLIFE_ERR_INVALID, "rmax=%g gives %lld cells (max %lld)", rmax, (long long)nx * nx, (long long)kMaxCells);
    g->nx = nx;
    g->ny = nx;
    g->cs = rmax;
    g->inv_cs = 1.0 / rmax;
    g->row_lo = 0;
    g->row_hi = nx;
    g->ly_shift = 0;
    g->nly = nx;
    g->rows_up = g->rows_dn = 0;
    g->ks = 0;
    g->staged = 0;
    g->rho = 0.f;
    if (nx > 16384) return fail(h, PLIFE_ERR_INVALID, "rmax=%g gives nx=%d (max 16384)", rmax, nx);
    if (h->slab.on) {
        const int G = h->slab.world, r = h->slab.rank, ny = g->ny;
        if (ny < 4 * G) return fail(h, PLIFE_ERR_INVALID, "slab mode needs ny >= 4*world (ny=%d, world=%d)", ny, G);
        auto lo = [&](int k) { return (int)((int64_t)k * ny / G); };
        g->row_lo = lo(r);
        g->row_hi = lo(r + 1);
        g->ly_shift = 1 - g->row_lo;
        g->nly = g->row_hi - g->row_lo + 2;
        const int up = (r + 1) % G, dn = (r + G - 1) % G;
        g->rows_up = lo(up + 1) - lo(up);
        g->rows_dn = lo(dn + 1) - lo(dn);
    }
    choose_kernel(h, g);
    return PLIFE_OK;
}

template <typename R>
int upload_matrix_t(plife_handle *h)
{
    int m = h->m;
    if (h->d_matrix_cap < m * m) {
        cudaFree(h->d_matrix_t);
        h->d_matrix_t = nullptr;
        h->d_matrix_cap = 0;
        CU(h, cudaMalloc(&h->d_matrix_t, 2 * sizeof(double) * (size_t)m * m));
        h->d_matrix_cap = m * m;
    }
    // device layout: transposed copy Mt[other][own], then the row-major copy M[own][other]
    std::vector<R> t((size_t)2 * m * m);
    for (int own = 0; own < m; own++)
        for (int other = 0; other < m; other++) {
            t[(size_t)other * m + own] = (R)h->matrix[(size_t)own * m + other];
            t[(size_t)m * m + (size_t)own * m + other] = (R)h->matrix[(size_t)own * m + other];
        }
    // pageable source: the copy is staged before the call returns
    CU(h, cudaMemcpyAsync(h->d_matrix_t, t.data(), sizeof(R) * t.size(), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->matrix_dirty = false;
    return PLIFE_OK;
}

int sync_matrix(plife_handle *h)
{
    if (!h->matrix_dirty) return PLIFE_OK;
    return h->precision == PLIFE_F32 ? upload_matrix_t<float>(h) : upload_matrix_t<double>(h);
}

template <typename R>
ForceParams<R> make_params(const plife_handle *h, const Grid &g, double dt)
{
    const plife_settings &s = h->settings;
    ForceParams<R> p{};
    p.n = (int)h->n;
    p.m = h->m;
    p.n_dev = h->slab.on && h->slab.counts ? &h->slab.cnt()->n : nullptr;
    p.tr = nullptr;
    p.bin_lo = 0;
    p.bin_hi = g.nxk() * g.nly - 1;
    p.first = h->slab.on ? (int)h->slab.halo_cap : 0;
    p.g = g;
    p.wrap = s.wrap ? 1 : 0;
    p.use_smem_matrix = h->m <= 64 ? 1 : 0;
    p.rmax = (R)s.rmax;
    p.r2 = (R)(s.rmax * s.rmax);              // B/Physics.java:432
    p.invr = (R)(1.0 / s.rmax);               // JOML div, :434
    p.mu = (R)pow(s.friction, 60 * dt);       // :401
    p.k2 = (R)(s.rmax * s.force * dt);        // :437
    p.dt = (R)dt;
    for (int k = 0; k < 4; k++) p.accp[k] = (R)h->acc_params[k];
    const double beta = h->acc_params[0];
    p.fast_b = (R)(beta * s.rmax);
    p.fast_d0 = (R)((1.0 + beta) * s.rmax * 0.5);
    p.fast_h = (R)((1.0 - beta) * s.rmax * 0.5);
    p.fast_a_scale = (R)(2.0 * beta / (1.0 - beta));
    p.fast_k = (R)(s.force * dt / beta);
    return p;
}

struct StepTimer {
    plife_handle *h;
    plife_handle::PendingTiming t{};
    bool on;
    explicit StepTimer(plife_handle *h_) : h(h_), on(h_->profiling) {}
    cudaError_t mark(int k)
    {
        if (!on) return cudaSuccess;
        cudaError_t e = cudaEventCreate(&t.ev[k]);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(t.ev[k], h->stream);
    }
};

int resolve_timings(plife_handle *h)
{
    if (h->pending.empty()) return PLIFE_OK;
    CU(h, cudaStreamSynchronize(h->stream));
    for (auto &p : h->pending) {
        for (int k = 0; k < PLIFE_K_COUNT; k++) {
            float ms = 0.f;
            CU(h, cudaEventElapsedTime(&ms, p.ev[k], p.ev[k + 1]));
            h->k_ms[k] += ms;
        }
        h->k_launches[PLIFE_K_BIN] += 1;
        h->k_launches[PLIFE_K_SCAN] += (h->flags & PLIFE_FLAG_SCAN3) ? 3 : 1;
        h->k_launches[PLIFE_K_SCATTER] += 1;
        h->k_launches[PLIFE_K_GATHER] += 1;
        h->k_launches[PLIFE_K_FORCE] += 1;
        for (int k = 0; k <= PLIFE_K_COUNT; k++) cudaEventDestroy(p.ev[k]);
    }
    h->pending.clear();
    return PLIFE_OK;
}

// makeContainers (B/Physics.java:309-354): state buffer cur -> sorted into cur^1
// small grids: one CTA does histogram + scan + scatter (cells.cu: small_sort), and the force pass skips the histogram
bool small_mode(const plife_handle *h, const Grid &g)
{
    return !h->slab.on && (int64_t)g.nxk() * g.nly <= kSmallBins && h->n_phys <= kSmallN && !(h->flags & PLIFE_FLAG_NO_FUSED_BIN);
}

int sort_current(plife_handle *h, const Grid &g, StepTimer *tm, bool mark_gather_end = true)
{
    int rc = ensure_cells(h, (int64_t)g.nxk() * g.nly);
    if (rc) return rc;
    if (tm) CU(h, tm->mark(0));
    const bool small = small_mode(h, g);
    h->small_step = small;
    // the previous force pass already binned its output for this grid: skip K_BIN (the bin words; the big path also needs
    // the histogram, which a small-mode step does not leave behind)
    const bool reuse = h->prebinned && (small || h->prebinned_counts) && h->prebinned_grid.nx == g.nx && h->prebinned_grid.cs == g.cs &&
                       h->prebinned_grid.ks == g.ks && h->prebinned_grid.row_lo == g.row_lo && h->prebinned_grid.row_hi == g.row_hi;
    if (!reuse) {
        if (h->slab.on) { // binning from scratch is sized by the host: it needs the exact counts (rare: first step, rmax change)
            rc = slab_refresh(h, true);
            if (rc) return rc;
        }
        if (h->count_dirty) CU(h, cudaMemsetAsync(h->d_count, 0, sizeof(int32_t) * (size_t)h->cell_cap, h->stream));
        CU(h, launch_bin(h, g));
        h->count_dirty = true;
    }
    h->prebinned = false;
    h->prebinned_counts = false;
    if (tm) CU(h, tm->mark(1));
    if (small) {
        CU(h, launch_small_sort(h, g));
        if (tm) CU(h, tm->mark(2));
    } else {
        CU(h, launch_scan(h, g));
        h->count_dirty = false; // K_SCAN zeroes the histogram after reading it
        if (tm) CU(h, tm->mark(2));
        CU(h, launch_scatter(h, g));
    }
    if (tm) CU(h, tm->mark(3));
    CU(h, launch_gather(h, g));
    h->n_sorted = h->n;
    if (tm && mark_gather_end) CU(h, tm->mark(4));
    return PLIFE_OK;
}

// host-side bookkeeping of a finished (queued) step; the graph replay path runs only this
void step_done_host(plife_handle *h, const Grid &g)
{
    // the force pass binned the new positions into d_cell (and, unless in small mode, d_count) for the same grid
    h->prebinned = !(h->flags & PLIFE_FLAG_NO_FUSED_BIN);
    h->prebinned_counts = h->prebinned && !h->small_step;
    h->prebinned_grid = g;
    if (h->prebinned_counts) h->count_dirty = true;
    h->n_phys = h->n;
    h->n_sorted = h->n;
    h->last_grid = g;
    h->has_sorted = true;
    h->steps++;
}

int run_step(plife_handle *h, double dt)
{
    Grid g;
    int rc = make_grid(h, &g);
    if (rc) return rc;
    rc = sync_matrix(h);
    if (rc) return rc;
    StepTimer tm(h);
    rc = sort_current(h, g, &tm);
    if (rc) return rc;
    if (h->precision == PLIFE_F32) CU(h, launch_force_f32(h, make_params<float>(h, g, dt)));
    else CU(h, launch_force_f64(h, make_params<double>(h, g, dt)));
    if (tm.on) {
        CU(h, tm.mark(PLIFE_K_COUNT));
        h->pending.push_back(tm.t);
        if (h->pending.size() >= 256) {
            rc = resolve_timings(h);
            if (rc) return rc;
        }
    }
    step_done_host(h, g);
    return PLIFE_OK;
}

// ---- launch-bound regime: replay the step as a CUDA graph ----
// Below kGraphMaxN particles a step is a handful of microsecond kernels and the host cannot launch them as fast as the
// device runs them.  Once the settings have been stable for two steps the step is captured (once per velocity-buffer
// parity: the fp32 force pass ping-pongs the two velocity arrays) and replayed with one cudaGraphLaunch.  Everything a
// captured launch depends on is in the key; any change re-captures.  PLIFE_FLAG_NO_GRAPH turns this off.
constexpr int64_t kGraphMaxN = 262144;

plife_handle::GraphKey graph_key(const plife_handle *h, double dt, const Grid &g)
{
    plife_handle::GraphKey k;
    memset(&k, 0, sizeof k);
    k.dt = dt;
    k.rmax = h->settings.rmax;
    k.friction = h->settings.friction;
    k.force = h->settings.force;
    for (int i = 0; i < 4; i++) k.accp[i] = h->acc_params[i];
    k.wrap = h->settings.wrap;
    k.acc_kind = h->acc_kind;
    k.m = h->m;
    k.flags = h->flags;
    k.ks = g.ks;
    k.staged = g.staged;
    k.n = h->n;
    k.matrix_version = h->matrix_version;
    if (h->precision == PLIFE_F32) {
        k.pt0 = h->s32[0].pt; k.pt1 = h->s32[1].pt; k.vel0 = h->s32[0].vel; k.vel1 = h->s32[1].vel;
    } else {
        k.pt0 = h->s64[0].pos; k.pt1 = h->s64[1].pos; k.vel0 = h->s64[0].vel; k.vel1 = h->s64[1].vel;
    }
    k.cell_end = h->d_cell_end;
    k.cur = h->cur;
    return k;
}

bool graph_eligible(const plife_handle *h)
{
    return !(h->flags & PLIFE_FLAG_NO_GRAPH) && !h->slab.on && !h->profiling && h->n > 0 && h->n <= kGraphMaxN && h->has_sorted &&
           h->prebinned && !h->matrix_dirty && h->n_phys == h->n;
}

int step_maybe_graph(plife_handle *h, double dt)
{
    if (!graph_eligible(h)) {
        h->stable_steps = 0;
        return run_step(h, dt);
    }
    Grid g;
    int rc = make_grid(h, &g);
    if (rc) return rc;
    const bool reuse = h->prebinned_grid.nx == g.nx && h->prebinned_grid.cs == g.cs && h->prebinned_grid.ks == g.ks &&
                       (small_mode(h, g) || h->prebinned_counts);
    if (!reuse) { // this step re-bins first: not the steady-state launch sequence
        h->stable_steps = 0;
        return run_step(h, dt);
    }
    const plife_handle::GraphKey key = graph_key(h, dt, g);
    for (auto &slot : h->graphs) {
        if (slot.valid && memcmp(&slot.key, &key, sizeof key) == 0) {
            CU(h, cudaGraphLaunch(slot.exec, h->stream));
            if (h->precision == PLIFE_F32) launch_force_f32_done(h);
            h->small_step = small_mode(h, g);
            step_done_host(h, g);
            h->graph_launches++;
            return PLIFE_OK;
        }
    }
    // stability: the key without the alternating velocity pointers
    plife_handle::GraphKey stable = key;
    stable.vel0 = stable.vel1 = nullptr;
    if (memcmp(&stable, &h->last_key, sizeof stable) == 0) h->stable_steps++;
    else h->stable_steps = 0;
    h->last_key = stable;
    if (h->stable_steps < 2) return run_step(h, dt);
    // capture this step
    plife_handle::GraphSlot &slot = h->graphs[key.vel0 < key.vel1 ? 0 : 1];
    if (slot.valid) {
        cudaGraphExecDestroy(slot.exec);
        slot.valid = false;
    }
    CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
    rc = run_step(h, dt);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (rc) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc;
    }
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(&slot.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return cuda_fail(h, e, "cudaGraphInstantiate");
    slot.key = key;
    slot.valid = true;
    h->graph_captures++;
    CU(h, cudaGraphLaunch(slot.exec, h->stream)); // the capture queued nothing: run the step now
    h->graph_launches++;
    return PLIFE_OK;
}

int valid_settings(plife_handle *h, const plife_settings *s)
{
    if (!s) return fail(h, PLIFE_ERR_INVALID, "settings is NULL");
    if (!(s->rmax > 0) || s->rmax > 1) return fail(h, PLIFE_ERR_INVALID, "rmax=%g outside (0,1]", s->rmax);
    if (!isfinite(s->friction) || !isfinite(s->force)) return fail(h, PLIFE_ERR_INVALID, "friction/force not finite");
    int nx = (int)floor(1 / s->rmax);
    if ((int64_t)nx * nx > kMaxCells) return fail(h, PLIFE_ERR_INVALID, "rmax=%g gives too many cells", s->rmax);
    return PLIFE_OK;
}

} // namespace

namespace plife {
int edit_fail(plife_handle *h, int code, const char *msg) { return fail(h, code, "%s", msg); }
int edit_grow(plife_handle *h, int64_t cap) { return grow_preserve(h, cap); }
int slab_make_grid(plife_handle *h, Grid *g) { return make_grid(h, g); }
int slab_fail(plife_handle *h, int code, const char *msg) { return fail(h, code, "%s", msg); }
// plife_slab_configure: buffers allocated before slab mode was switched on (plife_create with a capacity) have no room
// for the two ghost rows around the owned block of the sorted array; drop them and allocate again with that room.
int slab_reset_capacity(plife_handle *h)
{
    free_state(h);
    return h->capacity_hint > 0 ? ensure_capacity(h, h->capacity_hint) : PLIFE_OK;
}
// Slab-mode profiling: event 4 is recorded right before the first force launch, so the K_GATHER bucket also holds the
// halo pack; K_FORCE spans both force launches and the halo wait + unpack between them.
int slab_sort(plife_handle *h, const Grid &g)
{
    int rc = sync_matrix(h);
    if (rc) return rc;
    StepTimer tm(h);
    rc = sort_current(h, g, tm.on ? &tm : nullptr, false);
    h->slab_timing = tm.t;
    h->slab_timing_on = tm.on && rc == PLIFE_OK;
    return rc;
}
// One of the two force launches of a slab step: targets = the device-resident ranges d_tr[0..3], staging clamped to
// the bins [bin_lo, bin_hi] (the interior launch must not touch the ghost rows: they arrive later).
cudaError_t slab_force(plife_handle *h, const Grid &g, double dt, const int *d_tr, int nblocks, int bin_lo, int bin_hi, cudaStream_t stream,
                       bool first_part, bool last_part)
{
    // the interior launch (first_part) runs next to the migration exchange: its particles must stay in the slab
    const bool no_leavers = first_part && !last_part;
    cudaError_t e = cudaSuccess;
    if (first_part && h->slab_timing_on) { // profiling: event 4 = start of the force pass (main stream)
        StepTimer tm(h);
        tm.t = h->slab_timing;
        tm.on = true;
        e = tm.mark(4);
        h->slab_timing = tm.t;
    }
    ForceParams<float> p = make_params<float>(h, g, dt);
    p.tr = d_tr;
    p.bin_lo = bin_lo;
    p.bin_hi = bin_hi;
    if (e == cudaSuccess) e = launch_force_f32_part(h, p, nblocks, stream, no_leavers);
    return e;
}
// both launches are queued (and the main stream waits for the edge launch): the new state is the current one
cudaError_t slab_force_done(plife_handle *h, const Grid &g)
{
    cudaError_t e = cudaSuccess;
    launch_force_f32_done(h);
    if (h->slab_timing_on) {
        StepTimer tm(h);
        tm.t = h->slab_timing;
        tm.on = true;
        e = tm.mark(PLIFE_K_COUNT);
        if (e == cudaSuccess) {
            h->pending.push_back(tm.t);
            if (h->pending.size() >= 256 && resolve_timings(h) != PLIFE_OK) e = cudaErrorUnknown;
        }
    }
    h->slab_timing_on = false;
    h->prebinned = true; // the epilogue binned the stayers; arrivals are binned by phase FINISH
    h->prebinned_counts = true;
    h->prebinned_grid = g;
    h->count_dirty = true;
    h->last_grid = g;
    h->has_sorted = true;
    return e;
}
} // namespace plife

extern "C" {

int plife_version(void) { return PLIFE_VERSION; }

const char *plife_status_string(int status)
{
    switch (status) {
    case PLIFE_OK: return "ok";
    case PLIFE_ERR_INVALID: return "invalid argument";
    case PLIFE_ERR_OOM: return "out of memory";
    case PLIFE_ERR_CUDA: return "CUDA error";
    case PLIFE_ERR_NCCL: return "NCCL error";
    case PLIFE_ERR_STATE: return "invalid state";
    case PLIFE_ERR_STOPPED: return "stopped";
    default: return "unknown status";
    }
}

int plife_create(const plife_config *cfg, plife_handle **out)
{
    if (!cfg || !out) return PLIFE_ERR_INVALID;
    *out = nullptr;
    if (cfg->precision != PLIFE_F32 && cfg->precision != PLIFE_F64) return PLIFE_ERR_INVALID;
    if (cfg->bins != 0 && cfg->bins != 1 && cfg->bins != 2 && cfg->bins != 4 && cfg->bins != 8) return PLIFE_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return PLIFE_ERR_CUDA; // no CPU fallback
    }
    if (cfg->device < 0 || cfg->device >= ndev) return PLIFE_ERR_INVALID;
    if (cudaSetDevice(cfg->device) != cudaSuccess) return PLIFE_ERR_CUDA;
    plife_handle *h = new (std::nothrow) plife_handle();
    if (!h) return PLIFE_ERR_OOM;
    h->device = cfg->device;
    h->precision = cfg->precision;
    h->flags = cfg->flags;
    if (cfg->stream) {
        h->stream = (cudaStream_t)cfg->stream;
    } else {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete h;
            return PLIFE_ERR_CUDA;
        }
        h->own_stream = true;
    }
    if (cudaMalloc((void **)&h->d_scalar, (8 + 256) * sizeof(unsigned long long)) != cudaSuccess) {
        if (h->own_stream) cudaStreamDestroy(h->stream);
        delete h;
        return PLIFE_ERR_OOM;
    }
    cudaMemset(h->d_scalar, 0, (8 + 256) * sizeof(unsigned long long));
    h->d_hist = h->d_scalar + 8;
    if (cudaHostAlloc((void **)&h->h_maxocc, sizeof(int), cudaHostAllocMapped) == cudaSuccess) *h->h_maxocc = 0;
    else {
        cudaGetLastError();
        h->h_maxocc = nullptr;
    }
    h->capacity_hint = cfg->capacity;
    {   // fine bins per cell along x: plife_config.bins, or the environment (for experiments), else from the density
        int k = cfg->bins;
        if (k == 0)
            if (const char *ev = getenv("PLIFE_BINS")) k = atoi(ev);
        if (k == 1 || k == 2 || k == 4 || k == 8) h->bins_override = k == 1 ? 0 : (k == 2 ? 1 : (k == 4 ? 2 : 3));
    }
    if (cfg->capacity > 0) {
        int rc = ensure_capacity(h, cfg->capacity);
        if (rc) {
            plife_destroy(h);
            return rc;
        }
    }
    *out = h;
    return PLIFE_OK;
}

int plife_destroy(plife_handle *h)
{
    if (!h) return PLIFE_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto &p : h->pending)
        for (int k = 0; k <= PLIFE_K_COUNT; k++) cudaEventDestroy(p.ev[k]);
    for (auto &slot : h->graphs)
        if (slot.valid) cudaGraphExecDestroy(slot.exec);
    slab_destroy(h);
    free_state(h);
    cudaFree(h->d_count);
    if (h->d_cell_end) cudaFree(h->d_cell_end - 4);
    cudaFree(h->d_tile_sums);
    cudaFree(h->d_matrix_t);
    cudaFree(h->d_snap);
    if (h->snap_init) {
        cudaStreamSynchronize(h->copy_stream);
        for (int k = 0; k < 2; k++) {
            cudaFree(h->d_snap_async[k]);
            cudaEventDestroy(h->snap_ready[k]);
            cudaEventDestroy(h->snap_done[k]);
        }
        cudaStreamDestroy(h->copy_stream);
    }
    cudaFree(h->d_scalar);
    if (h->h_maxocc) cudaFreeHost((void *)h->h_maxocc);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
    return PLIFE_OK;
}

int plife_set_settings(plife_handle *h, const plife_settings *s)
{
    if (!h) return PLIFE_ERR_INVALID;
    int rc = valid_settings(h, s);
    if (rc) return rc;
    h->settings = *s;
    h->settings.wrap = s->wrap ? 1 : 0;
    return PLIFE_OK;
}

int plife_get_settings(const plife_handle *h, plife_settings *out)
{
    if (!h || !out) return PLIFE_ERR_INVALID;
    *out = h->settings;
    return PLIFE_OK;
}

int plife_set_matrix(plife_handle *h, int32_t m, const double *row_major)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (m < 1 || m > 256 || !row_major) return fail(h, PLIFE_ERR_INVALID, "matrix size %d outside [1,256] or NULL data", m);
    for (int k = 0; k < m * m; k++)
        if (!isfinite(row_major[k])) return fail(h, PLIFE_ERR_INVALID, "matrix entry %d not finite", k);
    if (h->n > 0 && h->max_type >= m)
        return fail(h, PLIFE_ERR_STATE, "resident particles have type %d >= new matrix size %d (retype on the host first)", h->max_type, m);
    h->m = m;
    h->matrix.assign(row_major, row_major + (size_t)m * m);
    h->matrix_dirty = true;
    h->matrix_version++;
    return PLIFE_OK;
}

int plife_set_matrix_entry(plife_handle *h, int32_t i, int32_t j, double v)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (i < 0 || j < 0 || i >= h->m || j >= h->m || !isfinite(v)) return fail(h, PLIFE_ERR_INVALID, "matrix entry (%d,%d) out of range for size %d", i, j, h->m);
    h->matrix[(size_t)i * h->m + j] = v;
    h->matrix_dirty = true;
    h->matrix_version++;
    return PLIFE_OK;
}

int plife_get_matrix(const plife_handle *h, int32_t *m_out, double *row_major_out, int32_t capacity_m)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (m_out) *m_out = h->m;
    if (row_major_out) {
        if (capacity_m < h->m) return PLIFE_ERR_INVALID;
        memcpy(row_major_out, h->matrix.data(), sizeof(double) * (size_t)h->m * h->m);
    }
    return PLIFE_OK;
}

int plife_set_accelerator(plife_handle *h, int32_t kind, const double *params, int32_t nparams)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (kind < 0 || kind >= PLIFE_ACC_KIND_COUNT) return fail(h, PLIFE_ERR_INVALID, "unknown accelerator kind %d", kind);
    if (nparams < 0 || nparams > 4 || (nparams > 0 && !params)) return fail(h, PLIFE_ERR_INVALID, "bad accelerator params");
    double p[4] = {0.3, 0, 0, 0};
    for (int k = 0; k < nparams; k++) p[k] = params[k];
    if (kind <= PLIFE_ACC_PARTICLE_LIFE_R2 && !(p[0] > 0 && p[0] < 1)) return fail(h, PLIFE_ERR_INVALID, "beta=%g outside (0,1)", p[0]);
    h->acc_kind = kind;
    memcpy(h->acc_params, p, sizeof p);
    return PLIFE_OK;
}

int plife_upload(plife_handle *h, int64_t n, const double *pos_xy, const double *vel_xy, const int32_t *type, const uint32_t *id)
{
    CHECK_HANDLE(h);
    if (n < 0 || (n > 0 && (!pos_xy || !type))) return fail(h, PLIFE_ERR_INVALID, "upload: n=%lld with NULL pos/type", (long long)n);
    int max_type = -1;
    uint32_t max_id = 0;
    for (int64_t i = 0; i < n; i++) {
        if (id && id[i] > max_id) max_id = id[i];
        double x = pos_xy[2 * i], y = pos_xy[2 * i + 1];
        if (!(x >= 0 && x <= 1 && y >= 0 && y <= 1)) return fail(h, PLIFE_ERR_INVALID, "upload: particle %lld position (%g,%g) outside [0,1]^2", (long long)i, x, y);
        int t = type[i];
        if (t < 0 || t >= h->m) return fail(h, PLIFE_ERR_INVALID, "upload: particle %lld type %d outside [0,%d)", (long long)i, t, h->m);
        if (t > max_type) max_type = t;
    }
    if (h->slab.on) { // a slab holds the particles of its own rows only (as the fp32 storage sees them)
        Grid g;
        int rcg = make_grid(h, &g);
        if (rcg) return rcg;
        for (int64_t i = 0; i < n; i++)
            if (container_of(cell_coords((double)(float)pos_xy[2 * i], (double)(float)pos_xy[2 * i + 1], g), g) < 0)
                return fail(h, PLIFE_ERR_INVALID, "upload: particle %lld (y=%g) does not belong to the rows [%d,%d) of slab rank %d", (long long)i,
                            pos_xy[2 * i + 1], g.row_lo, g.row_hi, h->slab.rank);
    }
    int rc = ensure_capacity(h, h->slab.on && h->capacity_hint > n ? h->capacity_hint : n);
    if (rc) return rc;
    h->cur = 0;
    h->has_sorted = false;
    h->prebinned = false;
    const int64_t chunk = 1 << 20;
    if (h->precision == PLIFE_F32) {
        std::vector<float4> pt((size_t)(n < chunk ? n : chunk));
        std::vector<float2> vl(pt.size());
        for (int64_t s = 0; s < n; s += chunk) {
            int64_t c = n - s < chunk ? n - s : chunk;
            for (int64_t k = 0; k < c; k++) {
                int64_t i = s + k;
                uint32_t pid = id ? id[i] : (uint32_t)i;
                float4 q;
                q.x = (float)pos_xy[2 * i];
                q.y = (float)pos_xy[2 * i + 1];
                memcpy(&q.z, &type[i], 4);
                memcpy(&q.w, &pid, 4);
                pt[k] = q;
                vl[k] = vel_xy ? make_float2((float)vel_xy[2 * i], (float)vel_xy[2 * i + 1]) : make_float2(0.f, 0.f);
            }
            CU(h, cudaMemcpyAsync(h->s32[0].pt + s, pt.data(), sizeof(float4) * c, cudaMemcpyHostToDevice, h->stream));
            CU(h, cudaMemcpyAsync(h->s32[0].vel + s, vl.data(), sizeof(float2) * c, cudaMemcpyHostToDevice, h->stream));
            CU(h, cudaStreamSynchronize(h->stream));
        }
    } else {
        if (n > 0) {
            CU(h, cudaMemcpyAsync(h->s64[0].pos, pos_xy, sizeof(double2) * n, cudaMemcpyHostToDevice, h->stream));
            if (vel_xy) CU(h, cudaMemcpyAsync(h->s64[0].vel, vel_xy, sizeof(double2) * n, cudaMemcpyHostToDevice, h->stream));
            else CU(h, cudaMemsetAsync(h->s64[0].vel, 0, sizeof(double2) * n, h->stream));
            CU(h, cudaMemcpyAsync(h->s64[0].type, type, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
            if (id) {
                CU(h, cudaMemcpyAsync(h->s64[0].id, id, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, h->stream));
            } else {
                std::vector<uint32_t> ids((size_t)n);
                for (int64_t i = 0; i < n; i++) ids[i] = (uint32_t)i;
                CU(h, cudaMemcpyAsync(h->s64[0].id, ids.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, h->stream));
                CU(h, cudaStreamSynchronize(h->stream));
            }
            CU(h, cudaStreamSynchronize(h->stream));
        }
    }
    h->n = n;
    h->n_phys = n;
    h->slab.n_old = n;
    h->slab.k_below = h->slab.k_above = 0;
    h->max_type = max_type;
    h->next_id = id ? (n ? max_id + 1 : 0) : (uint32_t)n;
    return slab_set_counts(h);
}

int plife_download(plife_handle *h, double *pos_xy, double *vel_xy, int32_t *type, uint32_t *id)
{
    CHECK_HANDLE(h);
    CU(h, cudaStreamSynchronize(h->stream));
    if (int rcs = slab_refresh(h, true)) return rcs;
    int64_t n = h->n_phys;
    if (n == 0) return PLIFE_OK;
    if (h->slab.on && h->slab.phase != PLIFE_SLAB_SORT) return fail(h, PLIFE_ERR_STATE, "download between slab phases");
    if (h->precision == PLIFE_F32) {
        const int64_t chunk = 1 << 20;
        std::vector<float4> pt((size_t)(n < chunk ? n : chunk));
        std::vector<float2> vl(pt.size());
        std::vector<int32_t> live;
        const bool filter = h->slab.on && h->n_phys != h->n; // dead slots of particles that migrated away
        if (filter) live.resize(pt.size());
        int64_t o = 0;
        for (int64_t s = 0; s < n; s += chunk) {
            int64_t c = n - s < chunk ? n - s : chunk;
            CU(h, cudaMemcpy(pt.data(), h->s32[h->cur].pt + s, sizeof(float4) * c, cudaMemcpyDeviceToHost));
            if (vel_xy) CU(h, cudaMemcpy(vl.data(), h->s32[h->cur].vel + s, sizeof(float2) * c, cudaMemcpyDeviceToHost));
            if (filter) CU(h, cudaMemcpy(live.data(), h->d_cell + s, sizeof(int32_t) * c, cudaMemcpyDeviceToHost));
            for (int64_t k = 0; k < c; k++) {
                if (filter && live[k] < 0) continue;
                int64_t i = o++;
                if (pos_xy) {
                    pos_xy[2 * i] = pt[k].x;
                    pos_xy[2 * i + 1] = pt[k].y;
                }
                if (vel_xy) {
                    vel_xy[2 * i] = vl[k].x;
                    vel_xy[2 * i + 1] = vl[k].y;
                }
                if (type) memcpy(&type[i], &pt[k].z, 4);
                if (id) memcpy(&id[i], &pt[k].w, 4);
            }
        }
    } else {
        const StateF64 &s = h->s64[h->cur];
        if (pos_xy) CU(h, cudaMemcpy(pos_xy, s.pos, sizeof(double2) * n, cudaMemcpyDeviceToHost));
        if (vel_xy) CU(h, cudaMemcpy(vel_xy, s.vel, sizeof(double2) * n, cudaMemcpyDeviceToHost));
        if (type) CU(h, cudaMemcpy(type, s.type, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
        if (id) CU(h, cudaMemcpy(id, s.id, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    }
    return PLIFE_OK;
}

int plife_download_f32(plife_handle *h, float *pos_xy, float *vel_xy, int32_t *type)
{
    CHECK_HANDLE(h);
    if (int rcs = slab_refresh(h, true)) return rcs;
    int64_t n = h->n;
    if (n == 0) return PLIFE_OK;
    if (h->snap_cap < n) { // sized by the handle's capacity: in slab mode n changes every step
        const int64_t want = h->cap > n ? h->cap : n;
        cudaFree(h->d_snap);
        h->d_snap = nullptr;
        h->snap_cap = 0;
        CU(h, cudaMalloc(&h->d_snap, (size_t)want * 20));
        h->snap_cap = want;
    }
    float2 *dp = (float2 *)h->d_snap;
    float2 *dv = dp + h->snap_cap;
    int32_t *dt = (int32_t *)(dv + h->snap_cap);
    CU(h, launch_snapshot_f32(h, pos_xy ? dp : nullptr, vel_xy ? dv : nullptr, type ? dt : nullptr));
    if (pos_xy) CU(h, cudaMemcpyAsync(pos_xy, dp, sizeof(float2) * n, cudaMemcpyDeviceToHost, h->stream));
    if (vel_xy) CU(h, cudaMemcpyAsync(vel_xy, dv, sizeof(float2) * n, cudaMemcpyDeviceToHost, h->stream));
    if (type) CU(h, cudaMemcpyAsync(type, dt, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return PLIFE_OK;
}

// Display-time handoff without stalling the physics: the snapshot kernel runs on the compute stream into one
// of two staging buffers, the device->host copies run on a separate copy stream, and the physics may step on
// while they are in flight.  The caller's buffers (ideally pinned) must stay valid until plife_snapshot_wait().
static int snapshot_async_impl(plife_handle *h, float *pos_xy, float *vel_xy, int32_t *type, uint8_t *type8)
{
    CHECK_HANDLE(h);
    if (int rcs = slab_refresh(h, true)) return rcs; // slab mode: the copy sizes need the exact count
    const int64_t n = h->n;
    if (n == 0) return PLIFE_OK;
    if (!h->snap_init) {
        CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CU(h, cudaEventCreateWithFlags(&h->snap_ready[k], cudaEventDisableTiming));
            CU(h, cudaEventCreateWithFlags(&h->snap_done[k], cudaEventDisableTiming));
        }
        h->snap_init = true;
    }
    if (h->snap_async_cap < n) { // sized by the handle's capacity: in slab mode n changes every step
        const int64_t want = h->cap > n ? h->cap : n;
        CU(h, cudaStreamSynchronize(h->copy_stream));
        CU(h, cudaStreamSynchronize(h->stream));
        for (int k = 0; k < 2; k++) {
            cudaFree(h->d_snap_async[k]);
            h->d_snap_async[k] = nullptr;
        }
        h->snap_async_cap = 0;
        h->snap_awaited = h->snap_issued; // (both streams are idle: nothing is in flight any more)
        for (int k = 0; k < 2; k++) CU(h, cudaMalloc(&h->d_snap_async[k], (size_t)want * 20));
        h->snap_async_cap = want;
    }
    // two snapshots may be in flight (one being copied while the next one is taken); a third waits for the oldest
    while (h->snap_issued - h->snap_awaited >= 2) {
        CU(h, cudaEventSynchronize(h->snap_done[h->snap_awaited & 1]));
        h->snap_awaited++;
    }
    const int k = (int)(h->snap_issued & 1);
    h->snap_issued++;
    float2 *dp = (float2 *)h->d_snap_async[k];
    float2 *dv = dp + h->snap_async_cap;
    int32_t *dt = (int32_t *)(dv + h->snap_async_cap);
    CU(h, cudaStreamWaitEvent(h->stream, h->snap_done[k], 0)); // the copy that last used this buffer has finished
    uint8_t *dt8 = reinterpret_cast<uint8_t *>(dt); // the compact form reuses the type region
    if (type && type8) return fail(h, PLIFE_ERR_INVALID, "snapshot: int32 and u8 types requested together");
    if (type8 && h->m > 256) return fail(h, PLIFE_ERR_STATE, "snapshot: more than 256 types do not fit u8");
    CU(h, launch_snapshot_f32(h, pos_xy ? dp : nullptr, vel_xy ? dv : nullptr, type ? dt : nullptr, type8 ? dt8 : nullptr));
    CU(h, cudaEventRecord(h->snap_ready[k], h->stream));
    CU(h, cudaStreamWaitEvent(h->copy_stream, h->snap_ready[k], 0));
    if (pos_xy) CU(h, cudaMemcpyAsync(pos_xy, dp, sizeof(float2) * n, cudaMemcpyDeviceToHost, h->copy_stream));
    if (vel_xy) CU(h, cudaMemcpyAsync(vel_xy, dv, sizeof(float2) * n, cudaMemcpyDeviceToHost, h->copy_stream));
    if (type) CU(h, cudaMemcpyAsync(type, dt, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->copy_stream));
    if (type8) CU(h, cudaMemcpyAsync(type8, dt8, (size_t)n, cudaMemcpyDeviceToHost, h->copy_stream));
    CU(h, cudaEventRecord(h->snap_done[k], h->copy_stream));
    return PLIFE_OK;
}

int plife_snapshot_async(plife_handle *h, float *pos_xy, float *vel_xy, int32_t *type)
{
    return snapshot_async_impl(h, pos_xy, vel_xy, type, nullptr);
}

int plife_snapshot_async_u8(plife_handle *h, float *pos_xy, float *vel_xy, uint8_t *type_u8)
{
    return snapshot_async_impl(h, pos_xy, vel_xy, nullptr, type_u8);
}

int plife_snapshot_wait(plife_handle *h)
{
    CHECK_HANDLE(h);
    // the OLDEST snapshot not yet handed over: with two requests in flight the caller gets the earlier one while the
    // later one is still crossing PCIe (true double buffering; one request in flight: the old behaviour)
    if (h->snap_init && h->snap_awaited < h->snap_issued) {
        CU(h, cudaEventSynchronize(h->snap_done[h->snap_awaited & 1]));
        h->snap_awaited++;
    }
    return PLIFE_OK;
}

int plife_init_uniform(plife_handle *h, int64_t n, uint64_t seed)
{
    CHECK_HANDLE(h);
    if (n < 0) return fail(h, PLIFE_ERR_INVALID, "n < 0");
    if (h->slab.on) {
        // n is the GLOBAL particle count; this rank keeps the particles of its own rows
        Grid g;
        int rc = make_grid(h, &g);
        if (rc) return rc;
        const int64_t share = n * (g.row_hi - g.row_lo) / g.ny;
        int64_t want = h->capacity_hint > 0 ? h->capacity_hint : share + share / 8 + 65536;
        rc = ensure_capacity(h, want);
        if (rc) return rc;
        h->cur = 0;
        h->has_sorted = false;
        h->prebinned = false;
        int *d_counter = reinterpret_cast<int *>(h->d_scalar + 6);
        CU(h, cudaMemsetAsync(d_counter, 0, sizeof(int), h->stream));
        CU(h, launch_init_uniform_owned(h, n, seed, g, d_counter));
        int kept = 0;
        CU(h, cudaMemcpyAsync(&kept, d_counter, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        if (kept > h->cap) return fail(h, PLIFE_ERR_OOM, "init_uniform: %d owned particles exceed capacity %lld", kept, (long long)h->cap);
        h->n = h->n_phys = kept;
        h->slab.n_old = kept;
        h->slab.k_below = h->slab.k_above = 0;
        h->slab.phase = PLIFE_SLAB_SORT;
        h->max_type = kept > 0 ? h->m - 1 : -1;
        return slab_set_counts(h);
    }
    int rc = ensure_capacity(h, n);
    if (rc) return rc;
    h->cur = 0;
    h->has_sorted = false;
    h->prebinned = false;
    h->n = n;
    h->n_phys = n;
    h->slab.n_old = n;
    h->slab.k_below = h->slab.k_above = 0;
    CU(h, launch_init_uniform(h, n, seed));
    h->max_type = n > 0 ? h->m - 1 : -1;
    h->next_id = (uint32_t)n;
    return PLIFE_OK;
}

int plife_random_matrix(plife_handle *h, int32_t m, uint64_t seed)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (m < 1 || m > 256) return fail(h, PLIFE_ERR_INVALID, "matrix size %d outside [1,256]", m);
    std::vector<double> v((size_t)m * m);
    const uint64_t s = seed ^ 0x4D41545249583634ull;
    for (int k = 0; k < m * m; k++) {
        uint64_t z = s + ((uint64_t)k + 1ull) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z = z ^ (z >> 31);
        v[k] = 2.0 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)) - 1.0; // B/DefaultMatrix.java:28
    }
    return plife_set_matrix(h, m, v.data());
}

int plife_step(plife_handle *h, double dt, int32_t nsteps)
{
    CHECK_HANDLE(h);
    if (!isfinite(dt) || nsteps < 0) return fail(h, PLIFE_ERR_INVALID, "step: dt=%g nsteps=%d", dt, nsteps);
    if (h->slab.on) return fail(h, PLIFE_ERR_STATE, "slab mode: drive the step with plife_slab_phase");
    for (int s = 0; s < nsteps; s++) {
        if (h->stop_requested.exchange(0)) return fail(h, PLIFE_ERR_STOPPED, "stopped after %d of %d steps", s, nsteps);
        int rc = step_maybe_graph(h, dt);
        if (rc) return rc;
    }
    return PLIFE_OK;
}

int plife_sync(plife_handle *h)
{
    CHECK_HANDLE(h);
    CU(h, cudaStreamSynchronize(h->stream));
    return slab_refresh(h, true); // slab mode: errors the device found in the queued steps surface here
}

int64_t plife_count(const plife_handle *h)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (h->slab.on && h->slab.counts && !h->poisoned) { // the count lives on the device: drain the queued steps and read it
        plife_handle *m = const_cast<plife_handle *>(h);
        if (cudaSetDevice(m->device) == cudaSuccess) slab_refresh(m, true);
    }
    return h->n;
}

int plife_type_histogram(plife_handle *h, int64_t *out_m)
{
    CHECK_HANDLE(h);
    if (!out_m) return fail(h, PLIFE_ERR_INVALID, "out is NULL");
    if (int rcs = slab_refresh(h, true)) return rcs;
    unsigned long long *d_hist = h->d_hist; // 256 counters, allocated with the handle
    cudaError_t e = cudaMemsetAsync(d_hist, 0, sizeof(unsigned long long) * 256, h->stream);
    if (e == cudaSuccess) e = launch_type_histogram(h, d_hist);
    unsigned long long host[256];
    if (e == cudaSuccess) e = cudaMemcpyAsync(host, d_hist, sizeof host, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "type histogram");
    for (int k = 0; k < h->m; k++) out_m[k] = (int64_t)host[k];
    return PLIFE_OK;
}

int plife_request_stop(plife_handle *h)
{
    if (!h) return PLIFE_ERR_INVALID;
    h->stop_requested.store(1);
    return PLIFE_OK;
}

const char *plife_last_error(const plife_handle *h) { return h ? h->last_error.c_str() : "null handle"; }

int plife_get_containers(plife_handle *h, int32_t *out, int64_t capacity)
{
    CHECK_HANDLE(h);
    if (!h->has_sorted) return fail(h, PLIFE_ERR_STATE, "no step has run since the last upload");
    int64_t ncell = (int64_t)h->last_grid.nx * h->last_grid.nly;
    if (!out || capacity < ncell) return fail(h, PLIFE_ERR_INVALID, "containers: capacity %lld < %lld", (long long)capacity, (long long)ncell);
    // the device array holds one END offset per fine bin; a cell's END offset is that of its last bin.  d_perm is
    // scratch between steps and has at least one int per particle, which may be fewer than the cells
    int32_t *d_tmp = nullptr;
    CU(h, dev_alloc(&d_tmp, (size_t)ncell));
    cudaError_t e = launch_containers(h, h->last_grid, d_tmp);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_tmp, sizeof(int32_t) * ncell, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_tmp);
    if (e != cudaSuccess) return cuda_fail(h, e, "containers");
    return PLIFE_OK;
}

int plife_get_step_stats(plife_handle *h, plife_step_stats *out)
{
    CHECK_HANDLE(h);
    if (!out) return fail(h, PLIFE_ERR_INVALID, "out is NULL");
    if (!h->has_sorted) return fail(h, PLIFE_ERR_STATE, "no step has run since the last upload");
    // buffer cur^1 still holds the sorted pre-step state of the last step
    const Grid g = h->last_grid;
    CU(h, cudaMemsetAsync(h->d_scalar, 0, sizeof(unsigned long long), h->stream));
    // the sorted scratch holds the particles of the last SORT: in slab mode h->n has moved on since (migration),
    // and walking h->n targets would read cell words past the sorted block (the round-1 8-GPU abort, profiles/r2_n8_repro.md)
    auto pf = make_params<float>(h, g, 0.0);
    auto pd = make_params<double>(h, g, 0.0);
    pf.n = pd.n = (int)h->n_sorted;
    pf.n_dev = nullptr;
    if (h->slab.on && h->slab.d_tr) { // ... and the host may not know it: pack_halo left it on the device (d_tr[8..11] = {0, n, 0, 0})
        pf.tr = h->slab.d_tr + 8;
        pf.n = (int)h->cap;
    }
    if (h->precision == PLIFE_F32) CU(h, launch_pair_count_f32(h, pf, h->d_scalar));
    else CU(h, launch_pair_count_f64(h, pd, h->d_scalar));
    unsigned long long total = 0;
    CU(h, cudaMemcpyAsync(&total, h->d_scalar, sizeof total, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (int rcs = slab_refresh(h, true)) return rcs;
    out->n = h->slab.on ? h->n : h->n_sorted;
    out->nx = g.nx;
    out->ny = g.ny;
    out->pair_evals = (int64_t)total;
    out->steps = h->steps;
    out->graph_steps = h->graph_launches;
    return PLIFE_OK;
}

int plife_debug_neighbors(plife_handle *h, int32_t *count, uint64_t *hash)
{
    CHECK_HANDLE(h);
    if (!count || !hash) return fail(h, PLIFE_ERR_INVALID, "NULL output");
    int64_t n = h->n;
    if (n == 0) return PLIFE_OK;
    Grid g;
    int rc = make_grid(h, &g);
    if (rc) return rc;
    rc = sort_current(h, g, nullptr); // cur -> cur^1 (sorted)
    if (rc) return rc;
    int32_t *d_cnt = nullptr;
    unsigned long long *d_hash = nullptr;
    CU(h, dev_alloc(&d_cnt, (size_t)n));
    cudaError_t e = dev_alloc(&d_hash, (size_t)n);
    if (e == cudaSuccess)
        e = h->precision == PLIFE_F32 ? launch_neighbors_f32(h, make_params<float>(h, g, 0.0), d_cnt, d_hash)
                                      : launch_neighbors_f64(h, make_params<double>(h, g, 0.0), d_cnt, d_hash);
    if (e == cudaSuccess) e = cudaMemcpyAsync(count, d_cnt, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hash, d_hash, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_cnt);
    cudaFree(d_hash);
    if (e != cudaSuccess) return cuda_fail(h, e, "debug neighbors");
    // like makeContainers, the sort is visible: the sorted copy becomes the current state.  fp32: the scratch is in
    // compute order with shifted types and the velocities never moved: write the reference order back into the
    // current buffers (records cur^1 -> cur, velocities cur -> cur^1 and swap those two)
    if (h->precision == PLIFE_F32) {
        CU(h, launch_apply_sort_f32(h, g));
        float2 *t = h->s32[0].vel;
        h->s32[0].vel = h->s32[1].vel;
        h->s32[1].vel = t;
    } else {
        h->cur ^= 1;
    }
    h->has_sorted = false;
    h->prebinned = false;
    h->last_grid = g;
    return PLIFE_OK;
}

int plife_set_profiling(plife_handle *h, int32_t enabled)
{
    CHECK_HANDLE(h);
    int rc = resolve_timings(h);
    if (rc) return rc;
    h->profiling = enabled != 0;
    if (enabled) {
        for (int k = 0; k < PLIFE_K_COUNT; k++) {
            h->k_ms[k] = 0;
            h->k_launches[k] = 0;
        }
    }
    return PLIFE_OK;
}

int plife_kernel_times(plife_handle *h, double *ms_out, int64_t *launches_out)
{
    CHECK_HANDLE(h);
    int rc = resolve_timings(h);
    if (rc) return rc;
    for (int k = 0; k < PLIFE_K_COUNT; k++) {
        if (ms_out) ms_out[k] = h->k_ms[k];
        if (launches_out) launches_out[k] = h->k_launches[k];
    }
    return PLIFE_OK;
}

int plife_device_ptrs(plife_handle *h, void **pos, void **vel, void **type, void **id)
{
    if (!h) return PLIFE_ERR_INVALID;
    if (h->precision == PLIFE_F32) {
        if (pos) *pos = h->s32[h->cur].pt;
        if (vel) *vel = h->s32[h->cur].vel;
        if (type) *type = nullptr;
        if (id) *id = nullptr;
    } else {
        if (pos) *pos = h->s64[h->cur].pos;
        if (vel) *vel = h->s64[h->cur].vel;
        if (type) *type = h->s64[h->cur].type;
        if (id) *id = h->s64[h->cur].id;
    }
    return PLIFE_OK;
}

} // extern "C"
