// Does sm_100a's packed FFMA2 (fma.rn.f32x2) buy issue slots?  Compares, per SM-clock, loops of
//   A: 16 FFMA            B: 8 FFMA2 (same FLOPs)         C: 16 FFMA + 8 IADD3/LOP (ALU)      D: 8 FFMA2 + 8 ALU
//   E: 16 FFMA + 8 FMNMX  F: 8 FFMA2 + 8 FMNMX
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void fma2(float &x, float &y, float a, float b)
{
    asm volatile("{ .reg .b64 rx, ra, rb; mov.b64 rx, {%0,%1}; mov.b64 ra, {%2,%2}; mov.b64 rb, {%3,%3}; fma.rn.f32x2 rx, rx, ra, rb; mov.b64 {%0,%1}, rx; }"
                 : "+f"(x), "+f"(y) : "f"(a), "f"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b, long long *cyc)
{
    float x[16];
    int q[8];
    float m[8];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; i++) { q[i] = threadIdx.x + i; m[i] = threadIdx.x * 0.5f + i; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) fma2(x[2 * i], x[2 * i + 1], a, b);
        }
        if (MODE == 2 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; i++) q[i] = (q[i] ^ (q[i] >> 3)) + it;
        }
        if (MODE == 4 || MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; i++) m[i] = fminf(m[i] + 0.f, x[i]);
        }
    }
    long long t1 = clock64();
    float s = 0; int qs = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
#pragma unroll
    for (int i = 0; i < 8; i++) { qs += q[i]; s += m[i]; }
    out[blockIdx.x * 256 + threadIdx.x] = s + (float)qs;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char *name, float *d, long long *dc)
{
    const int iters = 4096;
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0000001f, 1e-7f, dc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    // 8 CTAs x 8 warps per SM = 64 warps = 16 per SMSP
    printf("%-28s %.2f cycles per loop iteration per warp-slot (16 warps/SMSP): %.3f cycles/iter/warp\n", name, (double)c / iters, (double)c / iters / 16.0);
}
int main()
{
    float *d; long long *dc; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&dc, 8);
    run<0>("A 16 FFMA", d, dc); run<1>("B 8 FFMA2", d, dc);
    run<2>("C 16 FFMA + 8x2 ALU", d, dc); run<3>("D 8 FFMA2 + 8x2 ALU", d, dc);
    run<4>("E 16 FFMA + 8 FADD/FMNMX", d, dc); run<5>("F 8 FFMA2 + 8 FADD/FMNMX", d, dc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
