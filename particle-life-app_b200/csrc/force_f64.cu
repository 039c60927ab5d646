// fp64 instantiations of the force/integrate pass.  This file is compiled with
// -fmad=false: the reference's JOML arithmetic is non-fused, and with every
// operation rounded separately this path is bit-identical to the CPU oracle.
#include "force_impl.cuh"

namespace plife {

static IOF64 make_io(plife_handle *h)
{
    const int src = h->cur ^ 1, dst = h->cur;
    return IOF64{h->s64[src], h->s64[dst]};
}

cudaError_t launch_force_f64(plife_handle *h, const ForceParams<double> &p)
{
    NextBin nb{nullptr, nullptr, {nullptr, nullptr}, 0, nullptr};
    if (!(h->flags & PLIFE_FLAG_NO_FUSED_BIN)) nb = NextBin{h->d_cell, h->small_step ? nullptr : h->d_count, {nullptr, nullptr}, 0, nullptr};
    return dispatch_force<IOF64, false>(make_io(h), h->d_cell_end, h->d_cell_sorted, p, (p.n + kForceThreads - 1) / kForceThreads,
                                        (const double *)h->d_matrix_t, h->acc_kind, nb, h->stream);
}

cudaError_t launch_neighbors_f64(plife_handle *h, const ForceParams<double> &p, int32_t *cnt, unsigned long long *hash)
{
    if (p.n == 0) return cudaSuccess;
    const int nb = (p.n + kForceThreads - 1) / kForceThreads;
    neighbors_kernel<IOF64><<<nb, kForceThreads, 0, h->stream>>>(make_io(h), h->d_cell_end, h->d_cell_sorted, p, cnt, hash);
    return cudaGetLastError();
}

cudaError_t launch_pair_count_f64(plife_handle *h, const ForceParams<double> &p, unsigned long long *d_total)
{
    if (p.n == 0) return cudaSuccess;
    const int nb = (p.n + kForceThreads - 1) / kForceThreads;
    pair_count_kernel<IOF64><<<nb, kForceThreads, 0, h->stream>>>(make_io(h), h->d_cell_end, h->d_cell_sorted, p, d_total);
    return cudaGetLastError();
}

} // namespace plife
