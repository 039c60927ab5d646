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
    return IOF32{h->s32[src].pt, h->s32[dst].vel, h->s32[dst].pt, h->s32[src].vel, h->d_src_sorted};
}

static NextBin next_bin(plife_handle *h)
{
    if (h->flags & PLIFE_FLAG_NO_FUSED_BIN) return NextBin{nullptr, nullptr, {nullptr, nullptr}, 0};
    NextBin nb{h->d_cell, h->d_count, {h->slab.on ? h->slab.mig_send[0] : nullptr, h->slab.on ? h->slab.mig_send[1] : nullptr}, (int)h->slab.mig_cap};
    return nb;
}

static cudaError_t launch_force_f32_impl(plife_handle *h, const ForceParams<float> &p);

// The kernels read the old velocities from s32[cur].vel and write the new ones into s32[cur ^ 1].vel; swapping the two
// pointers makes s32[cur] = {new positions, new velocities} again for every other entry point.
cudaError_t launch_force_f32(plife_handle *h, const ForceParams<float> &p)
{
    const cudaError_t e = launch_force_f32_impl(h, p);
    if (e == cudaSuccess && p.n > 0) {
        float2 *t = h->s32[0].vel;
        h->s32[0].vel = h->s32[1].vel;
        h->s32[1].vel = t;
    }
    return e;
}

static cudaError_t launch_force_f32_impl(plife_handle *h, const ForceParams<float> &p)
{
    const float *mt = (const float *)h->d_matrix_t;
    const float *mrow = mt + (size_t)p.m * p.m; // row-major copy follows the transposed one
    const double rho = (double)p.n / ((double)p.g.nx * (p.g.row_hi - p.g.row_lo)); // particles per owned cell
    // v3: two targets per lane.  Opt-in: on B200 it executes 6 % fewer instructions than v2 but its 36 KB of
    // shared memory per CTA caps occupancy at 24 warps/SM and it ends up 3 % slower (profiles/r1_force_kernel.md).
    if ((h->flags & PLIFE_FLAG_PAIRS) && !h->slab.on && h->acc_kind == PLIFE_ACC_PARTICLE_LIFE && p.m <= 16 && rho >= 2.0) {
        int cap = (int)(2 * kForceThreads + 4.0 * rho + 10.0 * sqrt(rho + 1.0) + 32.0);
        cap = (cap + 31) / 32 * 32;
        if (cap <= 1280) {
            const int64_t ncell = (int64_t)p.g.nx * (p.g.row_hi - p.g.row_lo);
            const int max_pairs = (int)((p.n + (p.n < ncell ? p.n : ncell)) / 2 + 1);
            return launch_force_pairs(make_io(h), h->d_cell_end, h->d_cell_sorted, h->d_pair_first,
                                      reinterpret_cast<const int32_t *>(h->d_scalar + 7), max_pairs, p, mrow, cap, next_bin(h), h->stream);
        }
    }
    // v2 staged kernel whenever the per-lane matrix table fits; v1 (global-memory walk) otherwise
    // (below ~4 particles per cell the per-CTA staging and table fill cost more than they save: v1 wins)
    if (p.m <= kTabMaxM && rho >= 4.0 && !(h->flags & PLIFE_FLAG_FORCE_V1)) {
        // capacity of one staged row range: 128 targets + the cells hanging over both ends + margin
        int cap = (int)(kForceThreads + 4.0 * rho + 8.0 * sqrt(rho + 1.0) + 32.0);
        cap = (cap + 31) / 32 * 32;
        if (cap > 1536) cap = 1536;
        return dispatch_force_staged(make_io(h), h->d_cell_end, h->d_cell_sorted, p, mrow, h->acc_kind, cap, next_bin(h), h->stream);
    }
    return dispatch_force<IOF32, true>(make_io(h), h->d_cell_end, h->d_cell_sorted, p, mt, h->acc_kind, next_bin(h), h->stream);
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
