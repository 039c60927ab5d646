// fp32 instantiations of the force/integrate pass (fused multiply-add allowed).
#include "force_impl.cuh"

namespace plife {

static IOF32 make_io(plife_handle *h)
{
    const int src = h->cur ^ 1, dst = h->cur; // sorted scratch -> current
    return IOF32{h->s32[src].pt, h->s32[src].vel, h->s32[dst].pt, h->s32[dst].vel};
}

cudaError_t launch_force_f32(plife_handle *h, const ForceParams<float> &p)
{
    return dispatch_force<IOF32, true>(make_io(h), h->d_cell_end, p, (const float *)h->d_matrix_t, h->acc_kind, h->stream);
}

cudaError_t launch_neighbors_f32(plife_handle *h, const ForceParams<float> &p, int32_t *cnt, unsigned long long *hash)
{
    if (p.n == 0) return cudaSuccess;
    const int nb = (p.n + kForceThreads - 1) / kForceThreads;
    neighbors_kernel<IOF32><<<nb, kForceThreads, 0, h->stream>>>(make_io(h), h->d_cell_end, p, cnt, hash);
    return cudaGetLastError();
}

cudaError_t launch_pair_count_f32(plife_handle *h, const ForceParams<float> &p, unsigned long long *d_total)
{
    if (p.n == 0) return cudaSuccess;
    const int nb = (p.n + kForceThreads - 1) / kForceThreads;
    pair_count_kernel<IOF32><<<nb, kForceThreads, 0, h->stream>>>(make_io(h), h->d_cell_end, p, d_total);
    return cudaGetLastError();
}

} // namespace plife
