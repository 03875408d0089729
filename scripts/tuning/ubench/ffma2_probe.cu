// ffma2_probe.cu -- does the packed FP32 FMA of sm_100 (PTX fma.rn.f32x2, SASS FFMA2) free issue slots?
// Every kernel is a loop over an unrolled body of independent chains; each warp reports clock64 deltas, the host
// prints warp-instructions per cycle per SM sub-partition (issue rate) and FP32 FMA lanes per cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// MODE 0: 16 scalar FFMA      1: 8 FFMA2      2: 16 FFMA + 8 FMNMX + 2 MUFU      3: 8 FFMA2 + 8 FMNMX + 2 MUFU
// MODE 4: 8 FMUL2 (mul.f32x2) 5: 8 FADD2      6: 16 FFMA + 8 FMNMX + 4 MUFU + 4 FSETP/FSEL   7: same with 8 FFMA2
// MODE 8: 8 FFMA2 whose inputs are re-packed from scalars every trip (pack cost)
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, long long *cyc, int iters, float s0, float s1) {
    float a[16];
    u64 p[8];
    float m[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = s0 + threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { p[i] = pk(a[2 * i], a[2 * i + 1]); m[i] = a[i] * 0.5f; }
    const u64 c2 = pk(s1, s1 * 0.999f), d2 = pk(s0 * 1e-3f, s0 * 2e-3f);
    float x0 = s0, x1 = s1, x2 = s0 + s1, x3 = s0 - s1;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2 || MODE == 6) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s1, s0);
        }
        if (MODE == 1 || MODE == 3 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], c2, d2);
        }
        if (MODE == 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = mul2(p[i], c2);
        }
        if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = add2(p[i], d2);
        }
        if (MODE == 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const u64 r = fma2(pk(a[2 * i], a[(2 * i + 3) & 15]), c2, d2);
                a[2 * i] = lo(r);
                a[(2 * i + 3) & 15] = hi(r);
            }
        }
        if (MODE == 2 || MODE == 3 || MODE == 6 || MODE == 7) {
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], MODE & 1 ? hi(p[i]) : a[i]) ;
            x0 = ex2(x0); x1 = ex2(x1);
        }
        if (MODE == 6 || MODE == 7) {
            x2 = ex2(x2); x3 = ex2(x3);
#pragma unroll
            for (int i = 0; i < 4; ++i) m[i] = (m[i + 4] > s0) ? m[i] : s1;
        }
    }
    const long long t1 = clock64();
    float acc = x0 + x1 + x2 + x3;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += lo(p[i]) + hi(p[i]) + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

template <int MODE>
void run(const char *name, int instr_per_trip, int fma_lanes_per_trip, int warps_per_sm) {
    const int sms = 148, threads = 256, blocks_per_sm = warps_per_sm * 32 / threads;
    const int blocks = sms * blocks_per_sm, iters = 20000;
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks * threads / 32);
    k<MODE><<<blocks, threads>>>(out, cyc, 100, 1.0001f, 0.9999f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, cyc, iters, 1.0001f, 0.9999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const int nw = blocks * threads / 32;
    long long *h = (long long *)malloc(sizeof(long long) * nw);
    cudaMemcpy(h, cyc, sizeof(long long) * nw, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < nw; ++i) mean += (double)h[i]; mean /= nw;
    const double warps_per_smsp = warps_per_sm / 4.0;
    printf("%-44s warps/SM %2d  ms %.3f  cycles/trip/warp %.2f  issue/clk/SMSP %.3f  fma lanes/clk/SM %.1f (of 128)%s\n", name,
           warps_per_sm, ms, mean / iters, warps_per_smsp * instr_per_trip / (mean / iters),
           4.0 * warps_per_smsp * fma_lanes_per_trip * 32 / (mean / iters),
           cudaGetLastError() == cudaSuccess ? "" : "  CUDA ERROR");
    cudaFree(out); cudaFree(cyc); free(h);
}

int main() {
    for (int w : {16, 32}) {
        run<0>("16 FFMA", 16 + 2, 16, w);
        run<1>("8 FFMA2", 8 + 2, 16, w);
        run<4>("8 FMUL2 (mul.f32x2)", 8 + 2, 16, w);
        run<5>("8 FADD2 (add.f32x2)", 8 + 2, 16, w);
        run<2>("16 FFMA + 8 FMNMX + 2 MUFU", 26 + 2, 16, w);
        run<3>("8 FFMA2 + 8 FMNMX + 2 MUFU", 18 + 2, 16, w);
        run<6>("16 FFMA + 8 FMNMX + 4 MUFU + 4 FSETP/FSEL", 36 + 2, 16, w);
        run<7>("8 FFMA2 + 8 FMNMX + 4 MUFU + 4 FSETP/FSEL", 28 + 2, 16, w);
        run<8>("8 FFMA2, operands re-packed from scalars", 8 + 2, 16, w);
    }
    return 0;
}
