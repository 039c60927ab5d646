// What is the real FP32 issue rate per SM sub-partition (SMSP) on B200 for the operand patterns of the force
// kernel?  One CTA per SM with W warps per SMSP; every warp runs ITER x 32 independent instructions of one kind.
// Reports warp-instructions per cycle per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 16; // independent chains per thread
template <int MODE>
__global__ void k(float *out, int iters, float a, float b, long long *cyc)
{
    float x[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 1e-3f + i;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; r++)
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(a), "f"(b));          // x*a+b : 3 regs
                if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %1;" : "+f"(x[i]) : "f"(a));                   // x*x+a : 2 regs
                if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f33D6BF95;" : "+f"(x[i]));             // immediates
                if (MODE == 3) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));                        // FADD 2 regs
                if (MODE == 4) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a));                           // FMNMX
                if (MODE == 5) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(x[i]) : "f"(x[(i + 1) % CH]), "f"(x[(i + 5) % CH])); // 3 varying regs
            }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char *name, float *d, long long *dc)
{
    const int iters = 2048;
    for (int wps : {1, 2, 4, 8}) {
        k<MODE><<<148, wps * 128>>>(d, iters, 1.0000001f, 1e-7f, dc);
        cudaDeviceSynchronize();
        long long c[148]; cudaMemcpy(c, dc, sizeof c, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; i++) avg += c[i]; avg /= 148;
        printf("%-22s warps/SMSP=%d : %.3f warp-instr/cycle/SMSP\n", name, wps, (double)wps * iters * 2 * CH / avg);
    }
}
int main()
{
    float *d; long long *dc; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&dc, 148 * 8);
    run<0>("FFMA x*a+b (3 regs)", d, dc); run<1>("FFMA x*x+a (2 regs)", d, dc); run<2>("FFMA imm", d, dc);
    run<3>("FADD", d, dc); run<4>("FMNMX", d, dc); run<5>("FFMA 3 varying regs", d, dc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
