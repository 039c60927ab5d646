// Microbenchmark: cycles per warp-level load for LDS/LDG of 4/8/16 bytes when the 32 lanes
// read K distinct addresses (K = 1, 2, 4, 32).  Decides how the force kernel should fetch
// neighbour candidates.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lsu lsu_wavefronts.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int UNROLL = 8;

template <typename T> __device__ __forceinline__ float sum(T v);
template <> __device__ __forceinline__ float sum<float>(float v) { return v; }
template <> __device__ __forceinline__ float sum<float2>(float2 v) { return v.x + v.y; }
template <> __device__ __forceinline__ float sum<float4>(float4 v) { return v.x + v.y + v.z + v.w; }

// lanes are split into `groups` groups; every group reads its own address; addresses advance each iteration
template <typename T, bool SHARED>
__global__ void __launch_bounds__(1024) probe(const T *__restrict__ g, int groups, int group_stride, float *out, long long *cycles)
{
    extern __shared__ __align__(16) unsigned char raw[];
    T *s = reinterpret_cast<T *>(raw);
    const int NE = 2048; // elements in the table
    for (int k = threadIdx.x; k < NE; k += blockDim.x) s[k] = g[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int grp = groups >= 32 ? lane : lane / (32 / groups);
    int idx = (grp * group_stride + (threadIdx.x >> 5)) & (NE - 1);
    float acc = 0.f;
    const T *base = SHARED ? s : g;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            T v = SHARED ? base[(idx + u * 5) & (NE - 1)] : __ldg(base + ((idx + u * 5) & (NE - 1)));
            acc += sum<T>(v);
        }
        idx = (idx + 41) & (NE - 1);
    }
    long long t1 = clock64();
    if (acc == 123.456f) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <typename T, bool SHARED>
void run(const char *name, const void *g, float *out, long long *dcyc)
{
    const int nb = 148;
    auto k = probe<T, SHARED>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * (int)sizeof(T));
    int groupss[] = {1, 2, 4, 8, 32};
    for (int gi = 0; gi < 5; gi++) {
        for (int stride : {1, 3, 8}) {
            int groups = groupss[gi];
            if (groups == 1 && stride != 1) continue;
            k<<<nb, 1024, 2048 * sizeof(T)>>>((const T *)g, groups, stride, out, dcyc);
            cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, dcyc, sizeof h, cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int b = 0; b < nb; b++) avg += (double)h[b];
            avg /= nb;
            double loads = (double)ITERS * UNROLL * 32; // warp-level loads per SM (32 warps)
            printf("%-10s groups=%2d stride=%d : %.2f cycles per warp-load per SM\n", name, groups, stride, avg / loads);
        }
    }
}

int main()
{
    void *g;
    cudaMalloc(&g, 2048 * 16);
    cudaMemset(g, 0, 2048 * 16);
    float *out;
    cudaMalloc(&out, 4);
    long long *dcyc;
    cudaMalloc(&dcyc, 148 * 8);
    run<float, true>("LDS.32", g, out, dcyc);
    run<float2, true>("LDS.64", g, out, dcyc);
    run<float4, true>("LDS.128", g, out, dcyc);
    run<float, false>("LDG.32", g, out, dcyc);
    run<float2, false>("LDG.64", g, out, dcyc);
    run<float4, false>("LDG.128", g, out, dcyc);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
