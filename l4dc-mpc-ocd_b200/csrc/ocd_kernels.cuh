// ocd_kernels.cuh -- the planner kernels of the batched MPC engine and their launchers.
//
//   k_solve    NaivePlanner.generate_plan for B problems: one thread per (problem, start); the
//              block holds P problems x S starts, start-major (thread = s*P + p), so a warp runs
//              one start of 32 consecutive problems and every global / shared access is
//              coalesced / conflict-free in p.
//   k_episode  MPC_ORD's receding-horizon loop for B worlds in ONE launch: per control step the
//              block rebuilds the other cars' predicted tracks, runs the same solve, picks the
//              first-minimum start, steps every car with the simulator dynamics and accumulates
//              the true-weight reward of the past state.
// Each exists in three forms picked at launch by the batch size: the throughput form above, its latency
// variant (template flag LAT: straight-line forward sweep, registers spent on overlapping the horizon steps)
// and the time-parallel form k_solve_tp / k_episode_tp (eight lanes per (problem, start), one horizon step
// per lane).  Runtime horizons (HT == 0) use the segmented adjoint with controls in shared memory.
// Each (H, other cars, math mode) specialisation is instantiated in its own translation unit
// (ocd_inst.cu compiled with -DOCD_HT/-DOCD_NO/-DOCD_PRECISE) so the library builds in parallel.
// Reference lines for each piece are cited in ocd_device.cuh and include/ocd_b200.h.
#pragma once

#include <cuda_runtime.h>

#include <cstdlib>

#include "ocd_b200.h"
#include "ocd_device.cuh"

namespace ocd {

// Register budgets (LAT is the kernel form: 0 throughput, 1 latency, 2 wide).  The register-resident throughput
// kernels (HT > 0) are capped at 72 registers: seven warps per SM sub-partition (nine 96-thread blocks per SM)
// instead of six at the 76-80 the code would like; the 24 bytes of spill that costs are paid back by the extra
// warp (5.09 -> 4.97 ms at the bench shape; 64 registers / eight warps measured no better).  The segmented
// throughput kernels get 96 registers (five warps per sub-partition).  The latency form takes what it needs
// (140-250).  The wide form is the same straight-line code held to 128 registers (four warps per sub-partition)
// with one other car, and with three at a compile-time horizon (which would otherwise settle just above 128), and
// to 168 (three warps) elsewhere: measured against 112 / 128 / 168 on every shape of the sweep
// (scripts/tuning/wide_regs.sh); it is the fastest form for large batches of most shapes (see pick_form).
// The medium-horizon kernels (HT = 9..24, "Q" kernels: controls and saved states in registers, the rest of the
// reverse sweep's inputs through shared memory) run straight-line code in every form: OCD_Q_REGS registers
// (wide and throughput form), 255 for the latency form.
#ifdef OCD_Q_REGS_ALL
#define OCD_Q_REGS(NOT_) OCD_Q_REGS_ALL
#else
#define OCD_Q_REGS(NOT_) ((NOT_) <= 2 ? 104 : 128)
#endif
// World-tile staging of the solve kernels: 1-D TMA bulk copies (stock) or per-thread loads (-DOCD_LDG_STAGE), see
// stage_world_tile below
#if !defined(OCD_TMA_STAGE) && !defined(OCD_LDG_STAGE)
#define OCD_TMA_STAGE 1
#endif
// Q kernels unroll the whole horizon: 15 steps with one other car are 1 470 hot instructions (24 KB) and already pay
// for it in instruction-fetch stalls; with five other cars the sweep is 64 KB and the kernel ran at 40 % issue
// utilisation (ncu: no_instruction 2.0 per issue).  So medium horizons with more than OCD_Q_MAX_NO other cars take
// the segmented adjoint with a constant segment count instead, whose loop bodies are five steps long.
#ifndef OCD_Q_MAX_NO
#define OCD_Q_MAX_NO 1
#endif
#define OCD_IS_Q(HT, NOT_) ((HT) >= 9 && (HT) <= 24 && (NOT_) <= OCD_Q_MAX_NO)
// Long compile-time horizons (HT >= 25, FAST), and medium ones with many cars: the segmented adjoint with a constant
// segment count
#define OCD_IS_SEGC(HT, NOT_) ((HT) >= 25 || ((HT) >= 9 && (NOT_) > OCD_Q_MAX_NO))
// registers of the constant-segment-count kernels: measured at H = 50, 2^20 problems, one other car -- 128: 59.9 ms,
// 152: 58.3 ms, 168: 59.4 ms (profiles/tuning/r02_long_h_b.log); 168 with more cars
#ifdef OCD_SEGC_REGS
#define OCD_SEGC_REGS_(NOT_) OCD_SEGC_REGS
#else
#define OCD_SEGC_REGS_(NOT_) ((NOT_) == 1 ? 152 : 168)
#endif
#ifndef OCD_SEGC_FR
#define OCD_SEGC_FR 1
#endif
#ifndef OCD_SEGC_SEG
#define OCD_SEGC_SEG 5      // steps per segment of the long compile-time horizons (checkpoint rows are sized for 5)
#endif
#ifndef OCD_WIDE_REGS_MANY
#define OCD_WIDE_REGS_MANY 168  // wide form, other car counts (tuning knob)
#endif
#ifndef OCD_WIDE_REGS1
#define OCD_WIDE_REGS1 128      // wide form, one other car (tuning knob)
#endif
// Q kernels: problems per block (tuning knob; larger blocks with a barrier per iteration keep an SM's warps on the
// same stretch of the unrolled sweep) -- OCD_Q_P problems x 3 starts; other start counts keep 32
#ifndef OCD_Q_P
#define OCD_Q_P 32
#endif
#ifndef OCD_Q_SYNC
#define OCD_Q_SYNC 0
#endif
#define OCD_KERNEL_BOUNDS(HT, NOT_, LAT)                                  \
    __launch_bounds__((OCD_IS_Q(HT, NOT_) && 3 * OCD_Q_P > kMaxThreads) ? 3 * OCD_Q_P : kMaxThreads, ((LAT) != 0 || (HT) > 0) ? 1 : 3)      \
    __maxnreg__((LAT) == 1 ? 255 : (OCD_IS_Q(HT, NOT_) ? OCD_Q_REGS(NOT_) : OCD_IS_SEGC(HT, NOT_) ? OCD_SEGC_REGS_(NOT_) : ((LAT) == 2 ? (((NOT_) == 1 || ((HT) > 0 && (NOT_) == 3)) ? OCD_WIDE_REGS1 : OCD_WIDE_REGS_MANY) : ((HT) > 0 ? 72 : 96))))
static constexpr int kP = 32;             // problems per block: one warp per start
static constexpr int kMaxThreads = 6 * kP; // S=6 starts

// ---------------------------------------------------------------------------------------------
// argument blocks (passed by value as kernel parameters)
// ---------------------------------------------------------------------------------------------
struct SolveArgs {
    const float *world;            // [C][4][B]
    const float *other_controls;   // [NO][H][2][Bo] or null
    long long    Bo;
    const float *weights;          // [K][Bw]
    long long    Bw;
    const int32_t *weight_idx;     // [B] or null
    const float *cur_speed;        // [B] or null
    float       *plan;             // [H][2][B]
    float       *losses;           // [S][B]
    int32_t     *best;             // [B]
    float       *all_plans;        // [S][H][2][B] or null
    long long    B;
    int          P;                // problems per block
};

struct EpisodeArgs {
    const float *robot_init;       // [4][B]
    const float *other_init;       // [NO][4][B] or null
    const float *plan_weights;     // [K][Bw]
    long long    Bw;
    const int32_t *weight_idx;     // [B] or null
    const float *true_weights;     // [K]
    const int32_t *unlucky_idx;    // [B] or null
    int          t0, T;
    float       *returns;          // [B]
    float       *traj_controls;    // [T][2][B] or null
    int32_t     *traj_best;        // [T][B] or null
    float       *traj_states;      // [T][C][4][B] or null
    float       *final_world;      // [C][4][B] or null
    long long    B;
    int          P;
};

// Column of the weight table problem b plans with.  An index outside [0, Bw) is clamped: the *_host entry points
// and the Python front-end reject such batches before the launch (OCD_EINVAL / ValueError); for device-resident
// indices, which the launcher cannot inspect without a synchronising copy, clamping keeps a bad index from
// becoming an out-of-bounds read.
__device__ __forceinline__ long long weight_column(const int32_t *idx, long long Bw, long long b) {
    if (idx) {
        const long long i = (long long)idx[b];
        return i < 0 ? 0 : (i >= Bw ? Bw - 1 : i);
    }
    return Bw == 1 ? 0 : b;
}

// shared-memory carve-up common to k_solve and k_episode (all float, P columns each)
struct Smem {
    float *oth;     // [H][NO][2][P]   other cars' predicted positions at steps 1..H
    float *wraw;    // [K][P]          planning weights, raw feature order
    float *loss;    // [S][P]
    float *u0;      // [S][2][P]       first control of every start (episode only)
    float *world;   // [C][4][P]       live world state (episode only)
    float *wtrue;   // [K]             true weights (episode only)
    float *useg;    // [S*P][2H | 1]     controls of every thread (segmented kernels only)
    float *ckpt;    // [S*P][4 nseg | 1] segment-start states (segmented kernels only)
    float4 *q;      // [S*P][H | 1]      (d_t, gx, hy, ke) of every step (medium-horizon Q kernels only; first in the carve-up)
#ifdef OCD_TMA_STAGE
    float *tile;    // [C*4][P]          world tile staged by bulk copies (A/B variant)
#endif
};

static constexpr int kSeg = 5;    // steps per segment of the runtime-horizon kernels (register budget of the H=5 kernel)

// ---------------------------------------------------------------------------------------------
// OCD_TMA_STAGE (on in the stock build; -DOCD_LDG_STAGE builds the per-thread-LDG variant -- the A/B measurement is in
// DESIGN.md and profiles/tuning/r02_tma_ab.log): stage the block's world
// tile [C*4][P] with 1-D bulk copies (cp.async.bulk.shared::cluster.global + mbarrier complete_tx; SASS UBLKCP)
// instead of per-thread LDG.  One elected thread arms the barrier with the tile's byte count and issues one copy
// per row (P consecutive problems = P*4 bytes, 16-byte aligned when B % 4 == 0 and the block is full); every
// thread then waits on the barrier's phase 0.  Ragged blocks and odd batch sizes take the plain loads.
// ---------------------------------------------------------------------------------------------
#ifdef OCD_TMA_STAGE
template <int P>
__device__ __forceinline__ void stage_world_tile(float *tile, const float *world, long long B, long long b0, int rows,
                                                 bool live_all, unsigned long long *bar) {
    const bool tma_ok = live_all && (B % 4 == 0) && (P % 4 == 0);
    if (tma_ok) {
        const unsigned bar_s = (unsigned)__cvta_generic_to_shared(bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)(rows * P * sizeof(float));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
            for (int r = 0; r < rows; ++r) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(tile + r * P);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(world + (size_t)r * B + b0), "r"((unsigned)(P * sizeof(float))), "r"(bar_s)
                             : "memory");
            }
        }
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar_s) : "memory");
    } else {
        for (int i = threadIdx.x; i < rows * P; i += blockDim.x) {
            const int r = i / P, q = i % P;
            const long long b = b0 + q < B ? b0 + q : B - 1;
            tile[i] = world[(size_t)r * B + b];
        }
        __syncthreads();
    }
}
#endif

// lin: the slab holds (x0, dx, y0, dy) per other car instead of a position per horizon step
// (worth its two extra FFMA per car and step only once the per-step slab would crowd out resident blocks)
// The Q kernels (qk) can take it too (-DOCD_Q_LIN=1): their per-step slab is what stands between five and six resident
// blocks per SM.  Measured (profiles/tuning/r02_ab_qlin.log, H = 15, 2^20 problems): six blocks are SLOWER, 17.0 vs
// 16.0 ms -- the unrolled 15-step sweep is bound by instruction fetch, and more warps at different addresses make that
// worse -- so it is off.
#ifndef OCD_Q_LIN
#define OCD_Q_LIN 0
#endif
__host__ __device__ inline bool slab_is_linear(bool seg, bool precise, int other_mode, int H, int NO, bool qk = false) {
    if (qk) return OCD_Q_LIN && !precise && other_mode == 0;
    return seg && !precise && other_mode == 0 && (H * NO >= 100 || H >= 32);
}

__host__ __device__ inline size_t smem_floats(int H, int NO, int K, int S, int P, bool episode, bool seg, bool lin,
                                              bool qk = false) {
    size_t n = (size_t)(lin ? 4 * NO : H * NO * 2) * P + (size_t)K * P + (size_t)S * P;
    if (episode) n += (size_t)S * 2 * P + (size_t)(NO + 1) * 4 * P + ((K + 3) / 4) * 4;
    if (seg) n += (size_t)S * P * (seg_u_stride(H) + seg_ck_stride(H, kSeg));
    if (qk) n += (size_t)S * P * q_thread_floats(H);
#ifdef OCD_TMA_STAGE
    if (!episode) n += (size_t)(NO + 1) * 4 * P + 4;      // the staged world tile (16-byte aligned) and its mbarrier
#endif
    return n;
}

__device__ __forceinline__ Smem carve(float *base, const KParams &k, int P, bool episode, bool seg, bool lin,
                                      bool qk = false) {
    Smem m;
    m.q = reinterpret_cast<float4 *>(base);            // 16-byte aligned: the float4 rows come first
    if (qk) base += (size_t)k.S * P * q_thread_floats(k.H);
#ifdef OCD_TMA_STAGE
    m.tile = base;                                     // [C*4][P] world tile, then the mbarrier (solve kernels)
    if (!episode) base += (size_t)(k.NO + 1) * 4 * P + 4;
#endif
    m.oth = base;
    m.wraw = m.oth + (size_t)(lin ? 4 * k.NO : k.H * k.NO * 2) * P;
    m.loss = m.wraw + (size_t)k.K * P;
    float *next = m.loss + (size_t)k.S * P;
    m.u0 = m.world = m.wtrue = m.useg = m.ckpt = nullptr;
    if (episode) {
        m.u0 = next;
        m.world = m.u0 + (size_t)k.S * 2 * P;
        m.wtrue = m.world + (size_t)(k.NO + 1) * 4 * P;
        next = m.wtrue + ((k.K + 3) / 4) * 4;
    }
    if (seg) {
        m.useg = next;
        m.ckpt = m.useg + (size_t)k.S * P * seg_u_stride(k.H);
    }
    return m;
}

// The planner's prediction of other car j over the horizon (naive_planner.py:47-67), written to
// the block slab.  (x, y, v, th) is the car's current state; oc = its known controls [H][2]
// with element stride ocs, or null for the constant-velocity model.
template <bool PRECISE>
__device__ __forceinline__ void predict_other(const KParams &k, float x, float y, float v, float th,
                                              const float *oc, long long ocs, float *oth_col, int j, int P,
                                              bool lin = false) {
    if (lin) {      // constant velocity: position after n steps = start + n * (cos th * v * dt, sin th * v * dt)
        float sn, cs;
        Mth<PRECISE>::sincos_(th, sn, cs);
        oth_col[(size_t)(j * 4 + 0) * P] = slab_x<PRECISE>(x);
        oth_col[(size_t)(j * 4 + 1) * P] = slab_x<PRECISE>(__fmul_rn(__fmul_rn(cs, v), k.dt));
        oth_col[(size_t)(j * 4 + 2) * P] = slab_y<PRECISE>(y);
        oth_col[(size_t)(j * 4 + 3) * P] = slab_y<PRECISE>(__fmul_rn(__fmul_rn(sn, v), k.dt));
        return;
    }
    for (int t = 0; t < k.H; ++t) {
        float a = 0.0f, om = 0.0f;
        if (oc) {
            a = oc[(size_t)(t * 2 + 0) * ocs];
            om = oc[(size_t)(t * 2 + 1) * ocs];
        }
        other_model_step<PRECISE>(x, y, v, th, oc != nullptr, a, om, k.dt, k.dt2);
        oth_col[(size_t)((t * k.NO + j) * 2 + 0) * P] = slab_x<PRECISE>(x);
        oth_col[(size_t)((t * k.NO + j) * 2 + 1) * P] = slab_y<PRECISE>(y);
    }
}

// ---------------------------------------------------------------------------------------------
// k_solve
// ---------------------------------------------------------------------------------------------
// LAT != 0: the straight-line forward sweep (see sgd_iteration): 1 the latency form launched for small batches,
// 2 the wide form (same code, 128 registers) launched for large batches of the one-other-car shapes.
#ifdef OCD_BLOCK_TIMES
__device__ unsigned long long g_block_times[3 * 4096];      // tuning: (start ns, end ns, SM id) of the first 4096 blocks
#endif
template <int HT, int NOT_, int LT, bool PRECISE, int LAT = 0>
__global__ void OCD_KERNEL_BOUNDS(HT, NOT_, LAT)
k_solve(const __grid_constant__ KParams k, const SolveArgs a) {
    extern __shared__ __align__(16) float smem_raw[];
#ifdef OCD_BLOCK_TIMES
    unsigned long long t_start = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
#endif
    constexpr bool SEGK = (HT == 0) || OCD_IS_SEGC(HT, NOT_);   // runtime or long horizon: segmented adjoint, controls in shared memory
    constexpr bool QK = OCD_IS_Q(HT, NOT_) && !PRECISE;   // medium horizon: (d, gx, hy, ke) of every step through shared memory
    constexpr int P = QK ? OCD_Q_P : kP;     // compile-time, so every slab access is base + immediate
    const bool lin = slab_is_linear(SEGK, PRECISE, k.other_mode, k.H, k.NO, QK);
    const Smem m = carve(smem_raw, k, P, false, SEGK, lin, QK);
    const int p = threadIdx.x % P, s = threadIdx.x / P;
    const long long b_raw = (long long)blockIdx.x * P + p;
    const bool live = b_raw < a.B;
    const long long b = live ? b_raw : a.B - 1;
    const long long B = a.B;

#ifdef OCD_TMA_STAGE
    const float *wsrc = m.tile + p;                    // this problem's column of the staged tile
    const long long wstr = P;
    stage_world_tile<P>(m.tile, a.world, B, (long long)blockIdx.x * P, (k.NO + 1) * 4, (long long)blockIdx.x * P + P <= B,
                        reinterpret_cast<unsigned long long *>(m.tile + (size_t)(k.NO + 1) * 4 * P));
#else
    const float *wsrc = a.world + b;
    const long long wstr = B;
#endif
    if (s == 0) {
        const long long wc = weight_column(a.weight_idx, a.Bw, b);
        for (int i = 0; i < k.K; ++i) m.wraw[i * P + p] = a.weights[(size_t)i * a.Bw + wc];
        for (int j = 0; j < k.NO; ++j) {
            const float *st = wsrc + (size_t)(j + 1) * 4 * wstr;
            const float *oc = nullptr;
            long long ocs = 0;
            if (k.other_mode == 1) {
                ocs = a.Bo;
                oc = a.other_controls + (size_t)j * k.H * 2 * a.Bo + (a.Bo == 1 ? 0 : b);
            }
            predict_other<PRECISE>(k, st[0], st[wstr], st[2 * wstr], st[3 * wstr], oc, ocs, m.oth + p, j, P, lin);
        }
    }
    __syncthreads();

    const float x0 = wsrc[0], y0 = wsrc[wstr], v0 = wsrc[2 * wstr], th0 = wsrc[3 * wstr];
    const GradW gw = make_gradw<LT>(k, m.wraw + p, P);
    const float speed = a.cur_speed ? a.cur_speed[b] : v0;
#ifdef OCD_STAGGER_NS
    // tuning: start the warps of an SM out of phase (a kernel of a few waves otherwise runs its warps in lock step:
    // every warp in the MUFU-heavy stretch of the sweep at the same time)
    if (!PRECISE && LAT == 2) __nanosleep(((blockIdx.x * 3u + (unsigned)s) % OCD_STAGGER_MOD) * OCD_STAGGER_NS);
#endif
    const int H = HT > 0 ? HT : k.H;
    Traj<(SEGK ? 1 : HT)> u;                         // register-resident controls (short and medium compile-time horizons)
    const SmemTraj us{SEGK ? m.useg + (size_t)threadIdx.x * seg_u_stride(k.H) : nullptr};   // shared-memory controls
    float loss;
    if constexpr (SEGK) {
        constexpr int HC = HT;                       // 0: runtime horizon
        constexpr bool FR = HC > 0 && OCD_SEGC_FR && !PRECISE;
        constexpr int SEGL = HC > 0 ? OCD_SEGC_SEG : kSeg;
        float *ck = m.ckpt + (size_t)threadIdx.x * seg_ck_stride(k.H, kSeg);
        loss = (!PRECISE && lin)
                   ? solve_start_seg<SEGL, NOT_, LT, PRECISE, !PRECISE, LAT, HC, FR>(k, gw, m.wraw + p, P, x0, y0, v0, th0,
                                                                                      m.oth + p, P, s, speed, us, ck)
                   : solve_start_seg<SEGL, NOT_, LT, PRECISE, false, LAT, HC, FR>(k, gw, m.wraw + p, P, x0, y0, v0, th0,
                                                                                  m.oth + p, P, s, speed, us, ck);
    } else if constexpr (QK) {
        init_start<HT>(k, s, speed, u);
        float4 *q = m.q + (size_t)threadIdx.x * q_stride(HT);
        // (cos, sin) rows behind the float4 rows; entry t of a row sits at index t - 1 (steps 1 .. HT-1 are stored)
        float2 *q2 = reinterpret_cast<float2 *>(m.q + (size_t)k.S * P * q_stride(HT)) + (size_t)threadIdx.x * q2_stride(HT) - 1;
        loss = lin ? solve_start_q<HT, NOT_, LT, LAT, true>(k, gw, m.wraw + p, P, x0, y0, v0, th0, m.oth + p, P, u, q, q2)
                   : solve_start_q<HT, NOT_, LT, LAT, false>(k, gw, m.wraw + p, P, x0, y0, v0, th0, m.oth + p, P, u, q, q2);
    } else {
        init_start<(SEGK ? 1 : HT)>(k, s, speed, u);
        loss = solve_start<(SEGK ? 1 : HT), NOT_, LT, PRECISE, LAT>(k, gw, m.wraw + p, P, x0, y0, v0, th0, m.oth + p, P, u);
    }
    m.loss[s * P + p] = loss;
    if (live) {
        a.losses[(size_t)s * B + b] = loss;
        if (a.all_plans) {
#pragma unroll(SEGK ? 1 : HT)
            for (int t = 0; t < H; ++t) {
                a.all_plans[((size_t)(s * H + t) * 2 + 0) * B + b] = SEGK ? us.acc(t) : u.acc(SEGK ? 0 : t);
                a.all_plans[((size_t)(s * H + t) * 2 + 1) * B + b] = SEGK ? us.ang(t) : u.ang(SEGK ? 0 : t);
            }
        }
    }
    __syncthreads();
    // losses.index(min(losses)): first strict minimum (naive_planner.py:161-164)
    int bi = 0;
    float bl = m.loss[p];
    for (int q = 1; q < k.S; ++q) {
        const float l = m.loss[q * P + p];
        if (l < bl) { bl = l; bi = q; }
    }
    if (live && s == bi) {
        a.best[b] = bi;
#pragma unroll(SEGK ? 1 : HT)
        for (int t = 0; t < H; ++t) {
            a.plan[(size_t)(t * 2 + 0) * B + b] = SEGK ? us.acc(t) : u.acc(SEGK ? 0 : t);
            a.plan[(size_t)(t * 2 + 1) * B + b] = SEGK ? us.ang(t) : u.ang(SEGK ? 0 : t);
        }
    }
#ifdef OCD_BLOCK_TIMES
    if (threadIdx.x == 0 && blockIdx.x < 4096) {
        unsigned long long t_end;
        unsigned smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_block_times[3 * blockIdx.x] = t_start;
        g_block_times[3 * blockIdx.x + 1] = t_end;
        g_block_times[3 * blockIdx.x + 2] = smid;
    }
#endif
}

// ---------------------------------------------------------------------------------------------
// k_episode
// ---------------------------------------------------------------------------------------------
template <int HT, int NOT_, int LT, bool PRECISE, int LAT = 0>
__global__ void OCD_KERNEL_BOUNDS(HT, NOT_, LAT)
k_episode(const __grid_constant__ KParams k, const __grid_constant__ ocd_scenario sc, const EpisodeArgs a) {
    extern __shared__ __align__(16) float smem_raw[];
    constexpr int P = kP;
    constexpr bool SEGK = (HT == 0);
    const bool lin = slab_is_linear(SEGK, PRECISE, k.other_mode, k.H, k.NO);
    const Smem m = carve(smem_raw, k, P, true, SEGK, lin);
    const int p = threadIdx.x % P, s = threadIdx.x / P;
    const long long b_raw = (long long)blockIdx.x * P + p;
    const bool live = b_raw < a.B;
    const long long b = live ? b_raw : a.B - 1;
    const long long B = a.B;
    const int C = k.NO + 1;

    if (threadIdx.x < k.K) m.wtrue[threadIdx.x] = a.true_weights[threadIdx.x];
    int unlucky = 0;
    if (s == 0) {
        const long long wc = weight_column(a.weight_idx, a.Bw, b);
        for (int i = 0; i < k.K; ++i) m.wraw[i * P + p] = a.plan_weights[(size_t)i * a.Bw + wc];
        for (int c = 0; c < 4; ++c) m.world[c * P + p] = a.robot_init[(size_t)c * B + b];
        for (int j = 0; j < k.NO; ++j)
            for (int c = 0; c < 4; ++c)
                m.world[((j + 1) * 4 + c) * P + p] =
                    a.other_init ? a.other_init[(size_t)(j * 4 + c) * B + b] : sc.init_state[j][c];
        if (a.unlucky_idx) unlucky = a.unlucky_idx[b];
    }
    __syncthreads();
    const GradW gw = make_gradw<LT>(k, m.wraw + p, P);
    float ret = 0.0f;

    for (int i = 0; i < a.T; ++i) {
        const int ti = a.t0 + i;
        if (s == 0) {
            // ReplanningCarWorld.step: teleport before anything else (replanning_world.py:31-34)
            if (sc.critical_t > 0 && ti + 1 == sc.critical_t && unlucky >= 1 && unlucky < C)
                for (int c = 0; c < 4; ++c) m.world[(unlucky * 4 + c) * P + p] = sc.teleport_state[c];
            if (a.traj_states && live)
                for (int c = 0; c < C * 4; ++c)
                    a.traj_states[((size_t)i * C * 4 + c) * B + b] = m.world[c * P + p];
            // true-weight reward of the past state (mpc_ord.py:96-99)
            {
                const float x = m.world[p], y = m.world[P + p], v = m.world[2 * P + p], th = m.world[3 * P + p];
                float sn, cs;
                Mth<PRECISE>::sincos_(th, sn, cs);
                ret = __fadd_rn(ret, reward_value<LT, PRECISE, false>(k, m.wtrue, 1, x, y, v, sn, m.world + 4 * P + p,
                                                           4 * P, P));
            }
            // what the planner assumes about the other cars (planner_car.py:58-80, naive_planner.py:47-67)
            for (int j = 0; j < k.NO; ++j) {
                const float *w = m.world + (size_t)(j + 1) * 4 * P + p;
                float x = w[0], y = w[P], v = w[2 * P], th = w[3 * P];
                if (lin) {
                    predict_other<PRECISE>(k, x, y, v, th, nullptr, 0, m.oth + p, j, P, true);
                    continue;
                }
                for (int t = 0; t < k.H; ++t) {
                    float oa = 0.0f, oo = 0.0f;
                    if (k.other_mode == 1 && sc.kind[j] == 1) {   // plan replayed from index 0 (quirk Q4)
                        const bool in_plan = t < sc.plan_len[j];
                        oa = in_plan ? sc.plan[j][t][0] : sc.control[j][0];
                        oo = in_plan ? sc.plan[j][t][1] : sc.control[j][1];
                    }
                    other_model_step<PRECISE>(x, y, v, th, k.other_mode == 1, oa, oo, k.dt, k.dt2);
                    m.oth[(size_t)((t * k.NO + j) * 2 + 0) * P + p] = slab_x<PRECISE>(x);
                    m.oth[(size_t)((t * k.NO + j) * 2 + 1) * P + p] = slab_y<PRECISE>(y);
                }
            }
        }
        __syncthreads();

        const float x0 = m.world[p], y0 = m.world[P + p], v0 = m.world[2 * P + p], th0 = m.world[3 * P + p];
        Traj<(HT > 0 ? HT : 1)> u;
        const SmemTraj us{SEGK ? m.useg + (size_t)threadIdx.x * seg_u_stride(k.H) : nullptr};
        float loss;
        if (SEGK) {
            float *ck = m.ckpt + (size_t)threadIdx.x * seg_ck_stride(k.H, kSeg);
            loss = (!PRECISE && lin)
                       ? solve_start_seg<kSeg, NOT_, LT, PRECISE, !PRECISE>(k, gw, m.wraw + p, P, x0, y0, v0, th0,
                                                                             m.oth + p, P, s, v0, us, ck)
                       : solve_start_seg<kSeg, NOT_, LT, PRECISE, false>(k, gw, m.wraw + p, P, x0, y0, v0, th0,
                                                                         m.oth + p, P, s, v0, us, ck);
        } else {
            init_start<(HT > 0 ? HT : 1)>(k, s, v0, u);
            loss = solve_start<(HT > 0 ? HT : 1), NOT_, LT, PRECISE, LAT>(k, gw, m.wraw + p, P, x0, y0, v0, th0,
                                                                           m.oth + p, P, u);
        }
        m.loss[s * P + p] = loss;
        m.u0[(s * 2 + 0) * P + p] = SEGK ? us.acc(0) : u.acc(0);
        m.u0[(s * 2 + 1) * P + p] = SEGK ? us.ang(0) : u.ang(0);
        __syncthreads();

        if (s == 0) {
            int bi = 0;
            float bl = m.loss[p];
            for (int q = 1; q < k.S; ++q) {
                const float l = m.loss[q * P + p];
                if (l < bl) { bl = l; bi = q; }
            }
            const float ua = m.u0[(bi * 2 + 0) * P + p], uw = m.u0[(bi * 2 + 1) * P + p];
            if (live) {
                if (a.traj_controls) {
                    a.traj_controls[((size_t)i * 2 + 0) * B + b] = ua;
                    a.traj_controls[((size_t)i * 2 + 1) * B + b] = uw;
                }
                if (a.traj_best) a.traj_best[(size_t)i * B + b] = bi;
            }
            // CarWorld.step second phase: every car integrates (world.py:106-107, car.py:87)
            {
                float x = x0, y = y0, v = v0, th = th0;
                dynamics_step<PRECISE>(x, y, v, th, ua, uw, k.dt, k.dt2, k.mu);
                m.world[p] = x; m.world[P + p] = y; m.world[2 * P + p] = v; m.world[3 * P + p] = th;
            }
            for (int j = 0; j < k.NO; ++j) {
                float *w = m.world + (size_t)(j + 1) * 4 * P + p;
                float x = w[0], y = w[P], v = w[2 * P], th = w[3 * P];
                const bool in_plan = sc.kind[j] == 1 && ti < sc.plan_len[j];   // fixed_plan_car.py:25-31
                const float oa = in_plan ? sc.plan[j][ti][0] : sc.control[j][0];
                const float oo = in_plan ? sc.plan[j][ti][1] : sc.control[j][1];
                dynamics_step<PRECISE>(x, y, v, th, oa, oo, k.dt, k.dt2, sc.friction[j]);
                w[0] = x; w[P] = y; w[2 * P] = v; w[3 * P] = th;
            }
        }
        __syncthreads();
    }
    if (s == 0 && live) {
        a.returns[b] = ret;
        if (a.final_world)
            for (int c = 0; c < C * 4; ++c) a.final_world[(size_t)c * B + b] = m.world[c * P + p];
    }
}

// ---------------------------------------------------------------------------------------------
// Time-parallel variants (solve_start_tp): TG = 8 lanes per (problem, start) and 4 problems per block for horizons
// up to 8, 16 lanes and 2 problems for horizons 9 .. 16 (k_solve_tp only: the episode kernels are instantiated for
// H = 5 / 6); thread = ((s * P + p) * TG + t): warp s still runs start s, of P problems x TG step-lanes.
// Everything but the solve itself is k_solve / k_episode with the per-problem work on lane t == 0.
// ---------------------------------------------------------------------------------------------
static constexpr int kTP = 4;                                       // problems per block with 8 lanes per (problem, start)
__host__ __device__ constexpr int tp_problems(int HT) { return 32 / tp_lanes(HT); }   // 4, or 2 with 16 lanes

__device__ __forceinline__ void tp_start_controls(const KParams &k, int s, float cur_speed, float &a0, float &w0) {
    a0 = (s >= 3) ? __fmul_rn(k.mu, __fmul_rn(cur_speed, cur_speed)) : 0.0f;      // naive_planner.py:107-118
    const int m = s % 3;
    w0 = (m == 0) ? 0.0f : ((m == 1) ? -k.turn : k.turn);
}

template <int HT, int NOT_, int LT>
__global__ void __launch_bounds__(6 * 32, 1) k_solve_tp(const __grid_constant__ KParams k, const SolveArgs a) {
    extern __shared__ __align__(16) float smem_raw[];
    constexpr int P = tp_problems(HT), TG = tp_lanes(HT);
    const Smem m = carve(smem_raw, k, P, false, false, false);
    const int t = threadIdx.x % TG, g = threadIdx.x / TG, p = g % P, s = g / P;
    const long long b_raw = (long long)blockIdx.x * P + p;
    const bool live = b_raw < a.B;
    const long long b = live ? b_raw : a.B - 1;
    const long long B = a.B;
#ifdef OCD_TMA_STAGE
    const float *wsrc = m.tile + p;
    const long long wstr = P;
    stage_world_tile<P>(m.tile, a.world, B, (long long)blockIdx.x * P, (k.NO + 1) * 4, (long long)blockIdx.x * P + P <= B,
                        reinterpret_cast<unsigned long long *>(m.tile + (size_t)(k.NO + 1) * 4 * P));
#else
    const float *wsrc = a.world + b;
    const long long wstr = B;
#endif
    if (s == 0 && t == 0) {
        const long long wc = weight_column(a.weight_idx, a.Bw, b);
        for (int i = 0; i < k.K; ++i) m.wraw[i * P + p] = a.weights[(size_t)i * a.Bw + wc];
        for (int j = 0; j < k.NO; ++j) {
            const float *st = wsrc + (size_t)(j + 1) * 4 * wstr;
            const float *oc = nullptr;
            long long ocs = 0;
            if (k.other_mode == 1) {
                ocs = a.Bo;
                oc = a.other_controls + (size_t)j * k.H * 2 * a.Bo + (a.Bo == 1 ? 0 : b);
            }
            predict_other<false>(k, st[0], st[wstr], st[2 * wstr], st[3 * wstr], oc, ocs, m.oth + p, j, P, false);
        }
    }
    __syncthreads();
    const float x0 = wsrc[0], y0 = wsrc[wstr], v0 = wsrc[2 * wstr], th0 = wsrc[3 * wstr];
    const GradW gw = make_gradw<LT>(k, m.wraw + p, P);
    float a0, w0;
    tp_start_controls(k, s, a.cur_speed ? a.cur_speed[b] : v0, a0, w0);
    Traj<HT> u;
    const float loss = solve_start_tp<HT, NOT_, LT>(k, gw, m.wraw + p, P, x0, y0, v0, th0, m.oth + p, P, t, a0, w0, u);
    if (t == 0) {
        m.loss[s * P + p] = loss;
        if (live) {
            a.losses[(size_t)s * B + b] = loss;
            if (a.all_plans) {
#pragma unroll
                for (int j = 0; j < HT; ++j) {
                    a.all_plans[((size_t)(s * HT + j) * 2 + 0) * B + b] = u.ua[j];
                    a.all_plans[((size_t)(s * HT + j) * 2 + 1) * B + b] = u.uw[j];
                }
            }
        }
    }
    __syncthreads();
    int bi = 0;
    float bl = m.loss[p];
    for (int q = 1; q < k.S; ++q) {
        const float l = m.loss[q * P + p];
        if (l < bl) { bl = l; bi = q; }
    }
    if (live && s == bi && t == 0) {
        a.best[b] = bi;
#pragma unroll
        for (int j = 0; j < HT; ++j) {
            a.plan[(size_t)(j * 2 + 0) * B + b] = u.ua[j];
            a.plan[(size_t)(j * 2 + 1) * B + b] = u.uw[j];
        }
    }
}

template <int HT, int NOT_, int LT>
__global__ void __launch_bounds__(6 * kTP * kTG, 1)
k_episode_tp(const __grid_constant__ KParams k, const __grid_constant__ ocd_scenario sc, const EpisodeArgs a) {
    extern __shared__ __align__(16) float smem_raw[];
    constexpr int P = kTP;
    const Smem m = carve(smem_raw, k, P, true, false, false);
    const int t = threadIdx.x % kTG, g = threadIdx.x / kTG, p = g % P, s = g / P;
    const bool lead = s == 0 && t == 0;                  // the thread that owns problem p's world
    const long long b_raw = (long long)blockIdx.x * P + p;
    const bool live = b_raw < a.B;
    const long long b = live ? b_raw : a.B - 1;
    const long long B = a.B;
    const int C = k.NO + 1;

    if (threadIdx.x < k.K) m.wtrue[threadIdx.x] = a.true_weights[threadIdx.x];
    int unlucky = 0;
    if (lead) {
        const long long wc = weight_column(a.weight_idx, a.Bw, b);
        for (int i = 0; i < k.K; ++i) m.wraw[i * P + p] = a.plan_weights[(size_t)i * a.Bw + wc];
        for (int c = 0; c < 4; ++c) m.world[c * P + p] = a.robot_init[(size_t)c * B + b];
        for (int j = 0; j < k.NO; ++j)
            for (int c = 0; c < 4; ++c)
                m.world[((j + 1) * 4 + c) * P + p] =
                    a.other_init ? a.other_init[(size_t)(j * 4 + c) * B + b] : sc.init_state[j][c];
        if (a.unlucky_idx) unlucky = a.unlucky_idx[b];
    }
    __syncthreads();
    const GradW gw = make_gradw<LT>(k, m.wraw + p, P);
    float ret = 0.0f;

    for (int i = 0; i < a.T; ++i) {
        const int ti = a.t0 + i;
        if (lead) {
            if (sc.critical_t > 0 && ti + 1 == sc.critical_t && unlucky >= 1 && unlucky < C)
                for (int c = 0; c < 4; ++c) m.world[(unlucky * 4 + c) * P + p] = sc.teleport_state[c];
            if (a.traj_states && live)
                for (int c = 0; c < C * 4; ++c)
                    a.traj_states[((size_t)i * C * 4 + c) * B + b] = m.world[c * P + p];
            {
                const float x = m.world[p], y = m.world[P + p], v = m.world[2 * P + p], th = m.world[3 * P + p];
                float sn, cs;
                Mth<false>::sincos_(th, sn, cs);
                ret = __fadd_rn(ret, reward_value<LT, false, false>(k, m.wtrue, 1, x, y, v, sn, m.world + 4 * P + p,
                                                                    4 * P, P));
            }
            for (int j = 0; j < k.NO; ++j) {
                const float *w = m.world + (size_t)(j + 1) * 4 * P + p;
                float x = w[0], y = w[P], v = w[2 * P], th = w[3 * P];
                for (int tq = 0; tq < k.H; ++tq) {
                    float oa = 0.0f, oo = 0.0f;
                    if (k.other_mode == 1 && sc.kind[j] == 1) {   // plan replayed from index 0 (quirk Q4)
                        const bool in_plan = tq < sc.plan_len[j];
                        oa = in_plan ? sc.plan[j][tq][0] : sc.control[j][0];
                        oo = in_plan ? sc.plan[j][tq][1] : sc.control[j][1];
                    }
                    other_model_step<false>(x, y, v, th, k.other_mode == 1, oa, oo, k.dt, k.dt2);
                    m.oth[(size_t)((tq * k.NO + j) * 2 + 0) * P + p] = slab_x<false>(x);
                    m.oth[(size_t)((tq * k.NO + j) * 2 + 1) * P + p] = slab_y<false>(y);
                }
            }
        }
        __syncthreads();

        const float x0 = m.world[p], y0 = m.world[P + p], v0 = m.world[2 * P + p], th0 = m.world[3 * P + p];
        float a0, w0;
        tp_start_controls(k, s, v0, a0, w0);
        Traj<HT> u;
        const float loss = solve_start_tp<HT, NOT_, LT>(k, gw, m.wraw + p, P, x0, y0, v0, th0, m.oth + p, P, t, a0, w0,
                                                        u);
        if (t == 0) {
            m.loss[s * P + p] = loss;
            m.u0[(s * 2 + 0) * P + p] = u.ua[0];
            m.u0[(s * 2 + 1) * P + p] = u.uw[0];
        }
        __syncthreads();

        if (lead) {
            int bi = 0;
            float bl = m.loss[p];
            for (int q = 1; q < k.S; ++q) {
                const float l = m.loss[q * P + p];
                if (l < bl) { bl = l; bi = q; }
            }
            const float ua = m.u0[(bi * 2 + 0) * P + p], uw = m.u0[(bi * 2 + 1) * P + p];
            if (live) {
                if (a.traj_controls) {
                    a.traj_controls[((size_t)i * 2 + 0) * B + b] = ua;
                    a.traj_controls[((size_t)i * 2 + 1) * B + b] = uw;
                }
                if (a.traj_best) a.traj_best[(size_t)i * B + b] = bi;
            }
            {
                float x = x0, y = y0, v = v0, th = th0;
                dynamics_step<false>(x, y, v, th, ua, uw, k.dt, k.dt2, k.mu);
                m.world[p] = x; m.world[P + p] = y; m.world[2 * P + p] = v; m.world[3 * P + p] = th;
            }
            for (int j = 0; j < k.NO; ++j) {
                float *w = m.world + (size_t)(j + 1) * 4 * P + p;
                float x = w[0], y = w[P], v = w[2 * P], th = w[3 * P];
                const bool in_plan = sc.kind[j] == 1 && ti < sc.plan_len[j];   // fixed_plan_car.py:25-31
                const float oa = in_plan ? sc.plan[j][ti][0] : sc.control[j][0];
                const float oo = in_plan ? sc.plan[j][ti][1] : sc.control[j][1];
                dynamics_step<false>(x, y, v, th, oa, oo, k.dt, k.dt2, sc.friction[j]);
                w[0] = x; w[P] = y; w[2 * P] = v; w[3 * P] = th;
            }
        }
        __syncthreads();
    }
    if (lead && live) {
        a.returns[b] = ret;
        if (a.final_world)
            for (int c = 0; c < C * 4; ++c) a.final_world[(size_t)c * B + b] = m.world[c * P + p];
    }
}

inline int cuda_status() { return cudaGetLastError() == cudaSuccess ? OCD_OK : OCD_ECUDA; }

template <typename KernelT>
inline int prepare_smem(KernelT kern, size_t bytes) {
    if (bytes > 48 * 1024) {
        if (bytes > 227 * 1024) return OCD_EUNSUP;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
            return OCD_ECUDA;
    }
    return OCD_OK;
}

// Which form runs (measured on B200, scripts/tuning/form_sweep*.py, form_ep.py, wide_regs.sh; times in
// DESIGN.md):
//  * time-parallel: up to 888 warps of 4 starts (two of its three-warp blocks per SM: 1 184 problems) -- below that its shorter dependent chain
//    wins, above it its 4x instruction count per solve loses;
//  * latency form: up to 2 048 warps (one other car: one wave of its four blocks per SM) or 4 096 warps;
//  * beyond that the WIDE form -- the same straight-line code under a register cap -- wherever it beats the
//    vote-guarded throughput form: compile-time horizons with one or two other cars (4.77 vs 4.96 ms at the
//    bench shape, 6.29 vs 6.84 ms with three cars) and every segmented kernel (H = 15: 4.74 vs 5.88 ms at 2.6*10^5
//    problems; H = 50: 4.37 vs 5.99 ms at 6.5*10^4).  Without the votes a horizon step is one basic block and the
//    scheduler overlaps the steps; that is worth more than the throughput form's extra warps and skipped blocks;
//  * compile-time horizons with four or more cars: the wide form with a step fence (forward_sweep's SF) above
//    10 240 warps (8.45 vs 9.57 ms with six cars at 2^20 problems, 2.17 vs 2.50 ms at 2.6*10^5; it loses at
//    6.5*10^4), throughput form below;
//  * throughput form: whole episodes beyond 8 192 warps (one other car) / 4 096 warps, PRECISE math, and when
//    forced.
// OCD_KERNEL_FORM=throughput|latency|wide|tp overrides the choice (tests and tuning; read at every launch).
enum { kFormAuto = 0, kFormThroughput, kFormLatency, kFormTp, kFormWide };
inline int forced_form() {
    const char *e = std::getenv("OCD_KERNEL_FORM");
    if (!e) return kFormAuto;
    if (e[0] == 't') return e[1] == 'h' ? kFormThroughput : (e[1] == 'p' ? kFormTp : kFormAuto);
    return e[0] == 'l' ? kFormLatency : (e[0] == 'w' ? kFormWide : kFormAuto);
}
inline long long batch_warps(long long B, int P, int S) { return ((B + P - 1) / P) * S; }
inline bool tiny_batch(long long B, int S, int P = kTP) {      // P: problems per warp of the time-parallel form
    const int f = forced_form();
    // 8 lanes per start: up to two three-warp blocks per SM, 888 warps (1 080 finite_horizon episodes 0.515 ms against 0.599 for
    // the latency form, 1 620 episodes 0.716 against 0.599: profiles/tuning/r02_tp_cross.log); 16 lanes (H = 9..16) up to ~1 800: H = 15, 720 problems 0.146 ms against
    // 0.207 ms for the latency form, 1 440 problems 0.191 against 0.186 (profiles/tuning/r02_tp16.log)
    return f ? f == kFormTp : batch_warps(B, P, S) <= (P == kTP ? 888 : 1800);
}
// -> 0 throughput, 1 latency, 2 wide.
inline int pick_form(long long B, int P, int S, bool has_lat, bool has_wide, bool one_other, bool episode,
                     bool many_cars_fixed_h = false) {
    const int f = forced_form();
    if (f == kFormThroughput || !has_lat) return 0;
    if (f == kFormLatency) return 1;
    if (f == kFormWide) return has_wide ? 2 : 1;
    const long long w = batch_warps(B, P, S);
    if (w <= (one_other ? 2048 : 4096)) return 1;
    if (!has_wide) return 0;
    if (episode) return (one_other && w <= 8192) ? 2 : 0;
    if (many_cars_fixed_h && w <= 10240) return 0;     // step-fenced wide form: three warps per sub-partition need big batches
    return 2;
}

// The form the (HT, NOT_, LT, PRECISE) specialisation runs for B problems: 0 throughput, 1 latency, 2 wide,
// 3 time-parallel.  One function for both launchers and for ocd_kernel_form.
template <int HT, int NOT_, int LT, bool PRECISE>
inline int choose_form(const KParams &k, long long B, bool episode) {
    constexpr bool HAS_LAT = HT > 0 && !PRECISE && NOT_ >= 1;          // compile-time horizon and car count
    constexpr bool SEG_LAT = HT == 0 && !PRECISE;                       // segmented kernels (solve only)
    constexpr bool REGRES = HT > 0 && !OCD_IS_Q(HT, NOT_) && !OCD_IS_SEGC(HT, NOT_);  // register-resident (short) horizons
    // time-parallel: episodes up to 8 steps of horizon (the instantiated episode kernels), solves up to 16
    if (HAS_LAT && HT <= (episode ? kTG : kTGMax) && tiny_batch(B, k.S, tp_problems(HT > 0 ? HT : 1))) return 3;
    if constexpr (OCD_IS_SEGC(HT, NOT_) && !PRECISE) {
        // constant-segment-count kernels: the wide form (152 / 168 registers) is at least as fast as the latency form
        // (255) from ~10^4 problems on, and 30-35 % faster at 16-32 thousand problems with two or more other cars,
        // where the general rule below would still pick the latency form (profiles/tuning/r02_form_sweep3.log)
        const int f = forced_form();
        if (!episode && (f == kFormAuto || f == kFormTp)) return batch_warps(B, kP, k.S) <= 768 ? 1 : 2;
    }
    if (episode) return pick_form(B, kP, k.S, HAS_LAT, HAS_LAT && NOT_ == 1, NOT_ == 1, true);
    return pick_form(B, kP, k.S, HAS_LAT || SEG_LAT, HAS_LAT || SEG_LAT, REGRES && NOT_ == 1, false, REGRES && NOT_ >= 3);
}

template <int HT, int NOT_, int LT, bool PRECISE>
int launch_solve_t(const KParams &k, const SolveArgs &a, cudaStream_t st) {
    constexpr bool SEGK = HT == 0 || OCD_IS_SEGC(HT, NOT_);
    const int P = (OCD_IS_Q(HT, NOT_) && !PRECISE) ? OCD_Q_P : a.P;
    if (P != a.P && k.S != 3) return OCD_EUNSUP;       // tuning builds only
    const size_t bytes = smem_floats(k.H, k.NO, k.K, k.S, P, false, SEGK,
                                     slab_is_linear(SEGK, PRECISE, k.other_mode, k.H, k.NO, OCD_IS_Q(HT, NOT_) && !PRECISE),
                                     OCD_IS_Q(HT, NOT_) && !PRECISE) * sizeof(float);
    constexpr bool HAS_LAT = HT > 0 && !PRECISE && NOT_ >= 1;
    constexpr bool ANY_LAT = HAS_LAT || (HT == 0 && !PRECISE);
    const int form = choose_form<HT, NOT_, LT, PRECISE>(k, a.B, false);
    if constexpr (HAS_LAT && HT <= kTGMax) {
        if (form == 3) {
            constexpr int TPP = tp_problems(HT);
            const size_t tb = smem_floats(k.H, k.NO, k.K, k.S, TPP, false, false, false) * sizeof(float);
            k_solve_tp<HT, NOT_, LT><<<(unsigned)((a.B + TPP - 1) / TPP), k.S * 32, tb, st>>>(k, a);
            return cuda_status();
        }
    }
    auto kern = k_solve<HT, NOT_, LT, PRECISE>;
    if (form == 1) kern = k_solve<HT, NOT_, LT, PRECISE, ANY_LAT ? 1 : 0>;
    if (form == 2) kern = k_solve<HT, NOT_, LT, PRECISE, ANY_LAT ? 2 : 0>;
    int rc = prepare_smem(kern, bytes);
    if (rc) return rc;
    const unsigned grid = (unsigned)((a.B + P - 1) / P);
    kern<<<grid, k.S * P, bytes, st>>>(k, a);
    return cuda_status();
}

template <int HT, int NOT_, int LT, bool PRECISE>
int launch_episode_t(const KParams &k, const ocd_scenario &sc, const EpisodeArgs &a, cudaStream_t st) {
    const size_t bytes = smem_floats(k.H, k.NO, k.K, k.S, a.P, true, HT == 0,
                                     slab_is_linear(HT == 0, PRECISE, k.other_mode, k.H, k.NO)) * sizeof(float);
    constexpr bool HAS_LAT = HT > 0 && !PRECISE && NOT_ >= 1;
    constexpr bool HAS_WIDE = HAS_LAT && NOT_ == 1;
    const int form = choose_form<HT, NOT_, LT, PRECISE>(k, a.B, true);
    if constexpr (HAS_LAT && HT <= kTG) {
        if (form == 3) {
            const size_t tb = smem_floats(k.H, k.NO, k.K, k.S, kTP, true, false, false) * sizeof(float);
            k_episode_tp<HT, NOT_, LT><<<(unsigned)((a.B + kTP - 1) / kTP), k.S * kTP * kTG, tb, st>>>(k, sc, a);
            return cuda_status();
        }
    }
    auto kern = k_episode<HT, NOT_, LT, PRECISE>;
    if (form == 1) kern = k_episode<HT, NOT_, LT, PRECISE, HAS_LAT ? 1 : 0>;
    if (form == 2) kern = k_episode<HT, NOT_, LT, PRECISE, HAS_WIDE ? 2 : 0>;
    int rc = prepare_smem(kern, bytes);
    if (rc) return rc;
    const unsigned grid = (unsigned)((a.B + a.P - 1) / a.P);
    kern<<<grid, k.S * a.P, bytes, st>>>(k, sc, a);
    return cuda_status();
}

// choose_form behind the launchers' signature, so that the same dispatch macro can ask for it
template <int HT, int NOT_, int LT, bool PRECISE>
int form_t(const KParams &k, long long B, bool episode) { return choose_form<HT, NOT_, LT, PRECISE>(k, B, episode); }

}  // namespace ocd
