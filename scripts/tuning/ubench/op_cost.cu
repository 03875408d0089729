// op_cost.cu -- effective issue cost of one instruction of each kind inside an FFMA-heavy stream (sm_100a):
// cycles per warp-trip of (24 independent FFMA + n ops of kind X) minus the same with n = 0, divided by n.
// 8 warps per SM sub-partition, every op on its own register chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o op_cost op_cost.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { FMNMX, FSETP_FSEL, FADD, FMUL, LOP3R, IADD3, MUFU_EX2, MUFU_RCP, MUFU_SIN, FSEL_ONLY, LDS32, LDS128, STS128, MOVR };

template <int KIND, int N>
__global__ void __launch_bounds__(512) k(float *out, int iters, float s0, float s1, int q) {
    __shared__ float4 sm[512];
    float a[24], m[N > 0 ? N : 1];
    unsigned u[N > 0 ? N : 1];
#pragma unroll
    for (int i = 0; i < 24; ++i) a[i] = s0 + threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < N; ++i) { m[i] = s1 * (i + 1) + threadIdx.x * 1e-4f; u[i] = threadIdx.x * 7 + i; }
    sm[threadIdx.x] = make_float4(s0, s1, s0, s1);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(&sm[threadIdx.x]);
    float4 v4 = make_float4(0, 0, 0, 0);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 24; ++i) {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(s1), "f"(s0));
            if (i < N) {
                if (KIND == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(a[(i + 12) % 24]));
                if (KIND == FSETP_FSEL) asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %2, p; }" : "+f"(m[i]) : "f"(a[(i + 12) % 24]), "f"(s1));
                if (KIND == FSEL_ONLY) asm volatile("{ .reg .pred p; setp.gt.s32 p, %3, 5; selp.f32 %0, %0, %2, p; }" : "+f"(m[i]) : "f"(a[(i + 12) % 24]), "f"(s1), "r"(it));
                if (KIND == FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(a[(i + 12) % 24]));
                if (KIND == FMUL) asm volatile("mul.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(a[(i + 12) % 24]));
                if (KIND == LOP3R) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % N]), "r"(__float_as_uint(a[(i + 12) % 24])));
                if (KIND == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(__float_as_uint(a[(i + 12) % 24])));
                if (KIND == MUFU_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
                if (KIND == MUFU_RCP) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
                if (KIND == MUFU_SIN) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(m[i]));
                if (KIND == LDS32) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m[i]) : "r"(sbase + 0 * i) : "memory");
                if (KIND == LDS128) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v4.x), "=f"(v4.y), "=f"(v4.z), "=f"(m[i]) : "r"(sbase) : "memory");
                if (KIND == STS128) asm volatile("st.shared.v4.f32 [%4], {%0, %1, %2, %3};" :: "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(m[i]), "r"(sbase) : "memory");
                if (KIND == MOVR) asm volatile("mov.b32 %0, %1;" : "=f"(m[i]) : "f"(a[(i + 12) % 24]));
            }
        }
    }
    float acc = v4.x + v4.y + v4.z;
#pragma unroll
    for (int i = 0; i < 24; ++i) acc += a[i];
#pragma unroll
    for (int i = 0; i < N; ++i) acc += m[i] + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int KIND, int N>
double run() {
    const int blocks = 148 * 2, threads = 512, iters = 20000;
    float *out; cudaMalloc(&out, sizeof(float) * blocks * threads);
    k<KIND, N><<<blocks, threads>>>(out, 100, 1.0001f, 0.9999f, 12345);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND, N><<<blocks, threads>>>(out, iters, 1.0001f, 0.9999f, 12345);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(out);
    return ms * 1e-3 * 1.965e9 / iters / 8.0;
}

template <int KIND>
void kind(const char *name, double base) {
    const double c4 = run<KIND, 4>(), c8 = run<KIND, 8>(), c12 = run<KIND, 12>();
    printf("%-12s +4: %5.1f (%.2f/op)  +8: %5.1f (%.2f/op)  +12: %5.1f (%.2f/op)   [24 FFMA alone: %.1f cycles]\n", name, c4,
           (c4 - base) / 4, c8, (c8 - base) / 8, c12, (c12 - base) / 12, base);
}

int main() {
    const double base = run<FADD, 0>();
    kind<FMNMX>("FMNMX", base); kind<FSETP_FSEL>("FSETP+FSEL", base); kind<FSEL_ONLY>("FSEL", base); kind<FADD>("FADD", base);
    kind<FMUL>("FMUL", base); kind<LOP3R>("LOP3", base); kind<IADD3>("IADD3", base); kind<MOVR>("MOV", base);
    kind<MUFU_EX2>("MUFU.EX2", base); kind<MUFU_RCP>("MUFU.RCP", base); kind<MUFU_SIN>("MUFU.SIN(+FMUL)", base);
    kind<LDS32>("LDS.32", base); kind<LDS128>("LDS.128", base); kind<STS128>("STS.128", base);
    return 0;
}
