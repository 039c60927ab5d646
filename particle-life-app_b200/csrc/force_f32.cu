// fp32 instantiations of the force/integrate pass (fused multiply-add allowed).
#include <stdlib.h>

#define PLIFE_FORCE_F32_TU 1 // the staged fp32 kernels are instantiated here and nowhere else
#include "force_impl.cuh"

namespace plife {

static IOF32 make_io(plife_handle *h)
{
    const int src = h->cur ^ 1, dst = h->cur; // sorted scratch -> current
    // velocities: read from the current buffer (pre-sort order, through d_src_sorted), written to the scratch one;
    // launch_force_f32 swaps the two pointers afterwards so that s32[cur] is again the complete new state
    return IOF32{h->s32[src].pt, h->s32[dst].vel, h->s32[dst].pt, h->s32[src].vel, h->d_src_sorted, h->slab.on ? (int)h->slab.halo_cap : 0,
                 (h->flags & PLIFE_FLAG_UNSTABLE_SORT) ? 0 : 1, stable_key_of(h)};
}

static NextBin next_bin(plife_handle *h, bool no_leavers = false)
{
    if (h->flags & PLIFE_FLAG_NO_FUSED_BIN) return NextBin{nullptr, nullptr, {nullptr, nullptr}, 0, nullptr};
    NextBin nb{h->d_cell, h->small_step ? nullptr : h->d_count, {h->slab.on ? h->slab.mig_send[0] : nullptr, h->slab.on ? h->slab.mig_send[1] : nullptr},
               (int)h->slab.mig_cap, no_leavers ? h->slab.d_err : nullptr};
    return nb;
}

// The kernels read the old velocities from s32[cur].vel and write the new ones into s32[cur ^ 1].vel; swapping the two
// pointers (launch_force_f32_done) makes s32[cur] = {new positions, new velocities} again for every other entry point.
// `nblocks` CTAs of 128 targets; p.tr / p.n_dev select device-resident target ranges (slab mode).
cudaError_t launch_force_f32_part(plife_handle *h, const ForceParams<float> &p, int nblocks, cudaStream_t stream, bool no_leavers)
{
    const float *mt = (const float *)h->d_matrix_t;
    const float *mrow = mt + (size_t)p.m * p.m; // row-major copy follows the transposed one
    const double rho = p.g.rho; // particles per cell (make_grid's estimate: the same on every rank of a slab run)
    if (p.g.staged == 2) // small particle counts: one warp per cell
        return dispatch_force_cells(make_io(h), h->d_cell_end, h->d_cell_sorted, p, mt, h->acc_kind, next_bin(h, no_leavers), stream);
    if (p.g.staged) {
        // capacity of one staged row range: the CTA's 128 targets + K bins on either side (2 rho (1 + 1/K) particles on
        // average) + 4.5 sigma of that count (uniform state; anything denser streams in chunks).
        // PLIFE_STAGE_CAP overrides (experiments).
        const double mean = kForceThreads + 2.0 * rho * (1.0 + 1.0 / (1 << p.g.ks));
        int cap = (int)(mean + 4.5 * sqrt(mean) + 8.0);
        static const int cap_env = getenv("PLIFE_STAGE_CAP") ? atoi(getenv("PLIFE_STAGE_CAP")) : 0;
        if (cap_env > 0) cap = cap_env;
        cap = (cap + 15) / 16 * 16;
        if (cap > 1536) cap = 1536;
        return dispatch_force_staged(make_io(h), h->d_cell_end, h->d_cell_sorted, p, nblocks, mrow, h->acc_kind, cap, next_bin(h, no_leavers), stream);
    }
    if (p.g.ks != 0) return cudaErrorInvalidValue; // the v1 kernel writes results at the compute slot (make_grid never pairs it with fine bins)
    return dispatch_force<IOF32, true>(make_io(h), h->d_cell_end, h->d_cell_sorted, p, nblocks, mt, h->acc_kind, next_bin(h, no_leavers), stream);
}

void launch_force_f32_done(plife_handle *h)
{
    float2 *t = h->s32[0].vel;
    h->s32[0].vel = h->s32[1].vel;
    h->s32[1].vel = t;
}

cudaError_t launch_force_f32(plife_handle *h, const ForceParams<float> &p)
{
    if (p.n == 0) return cudaSuccess;
    const cudaError_t e = launch_force_f32_part(h, p, (p.n + kForceThreads - 1) / kForceThreads, h->stream);
    if (e == cudaSuccess) launch_force_f32_done(h);
    return e;
}

cudaError_t launch_neighbors_f32(plife_handle *h, const ForceParams<float> &p, int32_t *cnt, unsigned long long *hash)
{
    if (p.n == 0) return cudaSuccess;
    const int nb = (p.n + kForceThreads - 1) / kForceThreads;
    neighbors_kernel<IOF32><<<nb, kForceThreads, 0, h->stream>>>(make_io(h), h->d_cell_end, h->d_cell_sorted, p, cnt, hash);
    return cudaGetLastError();
}

cudaError_t launch_pair_count_f32(plife_handle *h, const ForceParams<float> &p, unsigned long long *d_total)
{
    if (p.n == 0) return cudaSuccess;
    const int nb = (p.n + kForceThreads - 1) / kForceThreads;
    pair_count_kernel<IOF32><<<nb, kForceThreads, 0, h->stream>>>(make_io(h), h->d_cell_end, h->d_cell_sorted, p, d_total);
    return cudaGetLastError();
}

} // namespace plife
