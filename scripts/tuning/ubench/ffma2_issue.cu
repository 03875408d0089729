// ffma2_issue.cu -- does FFMA2 (fma.rn.f32x2) cost one issue slot or two?
// Loop body: NF independent packed FMAs (or 2*NF scalar FMAs: same FP32 lane work) + NA independent integer ALU ops.
//   two-slot model : cycles per warp-trip = 2*NF + NA  (same as scalar)
//   one-slot model : cycles per warp-trip = max(2*NF, NF + NA)  -- flat in NA until NA = NF
// Every SM gets the same number of warps (grid = 148 * k blocks of 1024 threads cannot be relied on to spread evenly, so
// the kernel is timed as a whole and reported as cycles per warp-trip per SM sub-partition assuming an even spread AND as
// plain milliseconds).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_issue ffma2_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int NF, int NA, bool PACKED, int NM>
__global__ void __launch_bounds__(512) k(float *out, int iters, float s0, float s1, int q) {
    u64 p[NF];
    float a[2 * NF];
    unsigned m[NA > 0 ? NA : 1];
    float x[NM > 0 ? NM : 1];
#pragma unroll
    for (int i = 0; i < 2 * NF; ++i) a[i] = s0 + threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < NF; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
#pragma unroll
    for (int i = 0; i < NA; ++i) m[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < NM; ++i) x[i] = s0 * (i + 1);
    const u64 c2 = pk(s1, s1 * 0.999f), d2 = pk(s0 * 1e-3f, s0 * 2e-3f);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (NF > NA ? NF : NA); ++i) {      // interleave the two kinds
            if (i < NF) {
                if (PACKED) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(d2));
                else {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[2 * i]) : "f"(s1), "f"(s0));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[2 * i + 1]) : "f"(s1), "f"(s0));
                }
            }
            if (i < NA) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[i]) : "r"(q), "r"(it));
            if (i < NM) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
    }
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < NF; ++i) { float u, v; up(p[i], u, v); acc += u + v + a[2 * i] + a[2 * i + 1]; }
#pragma unroll
    for (int i = 0; i < NA; ++i) acc += (float)m[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) acc += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int NF, int NA, bool PACKED, int NM>
void run() {
    const int blocks = 148 * 2, threads = 512, iters = 20000;      // 2 blocks of 16 warps per SM = 8 warps per sub-partition
    float *out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    k<NF, NA, PACKED, NM><<<blocks, threads>>>(out, 100, 1.0001f, 0.9999f, 12345);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<NF, NA, PACKED, NM><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f, 12345);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / iters / 8.0;      // cycles per warp-trip per sub-partition (8 warps each)
    printf("%s NF=%2d (%2d FP32 lane-FMAs) NA=%2d NM=%d : %.3f ms  %.1f cycles per warp-trip  (two-slot %d, one-slot %d)%s\n",
           PACKED ? "FFMA2" : "FFMA ", NF, 2 * NF, NA, NM, ms, cyc, 2 * NF + NA + NM + 3, (2 * NF > NF + NA + NM + 3 ? 2 * NF : NF + NA + NM + 3),
           cudaGetLastError() == cudaSuccess ? "" : " CUDA ERROR");
    cudaFree(out);
}

int main() {
    run<12, 0, false, 0>(); run<12, 0, true, 0>();
    run<12, 4, false, 0>(); run<12, 4, true, 0>();
    run<12, 8, false, 0>(); run<12, 8, true, 0>();
    run<12, 12, false, 0>(); run<12, 12, true, 0>();
    run<12, 8, false, 2>(); run<12, 8, true, 2>();
    run<8, 8, false, 2>(); run<8, 8, true, 2>();
    run<6, 12, false, 2>(); run<6, 12, true, 2>();
    return 0;
}
