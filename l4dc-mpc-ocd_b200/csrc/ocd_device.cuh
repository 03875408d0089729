// ocd_device.cuh -- device-side math of the batched MPC engine (sm_100a).
//
// One thread owns one (problem, start) pair and keeps the whole trajectory optimisation in
// registers: controls, saved forward quantities and feature gradients for all H steps.  The
// forward rollout evaluates the feature GRADIENT at each new state (the reward value itself is
// only needed once, for the final loss), the reverse sweep is the closed-form adjoint of the
// car dynamics, and the SGD update is applied in the same sweep.  What each block shares --
// the other cars' predicted positions and the raw weight vectors -- lives in shared memory.
//
// Reference semantics implemented here (paths relative to the reference checkout):
//   dynamics            interact_drive/simulation_utils.py:9-21
//   smooth helpers      interact_drive/math_utils.py:7-31, 59-97, 135-180
//   features            experiments/merging.py:32-83 ; interact_drive/world.py:206-218
//   reward              interact_drive/car/linear_reward_car.py:49-55
//   other-car model     interact_drive/planner/naive_planner.py:47-67
//   rollout + SGD       interact_drive/planner/naive_planner.py:32-79, 107-164
// Gradient conventions are TensorFlow 2.1's: clip masks inclusive, Minimum
// passes on <=, reduce_min/max split evenly among ties, where() blocks the unselected branch.
//
// Two math back-ends share every formula:
//   PRECISE  libdevice sinf/cosf/expf, IEEE division, reference op order where it matters.
//   FAST     one MUFU op per transcendental (sin/cos/ex2/rcp .approx), branch-free collision
//            bump, constants folded into the weights.  The hot loop is issue-bound, so FAST is
//            written to minimise the instruction count per (iteration x horizon step).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ocd_b200.h"

namespace ocd {

// Device-side digest of ocd_params: every Python-float constant already cast to float32 the
// way TensorFlow casts it at op time.
struct KParams {
    int   H, NO, L, K, n_iter, S, other_mode, extra_inits, optimizer;
    float lr, dt, dt2, hdt2;      // dt2 = (float)(dt*dt) squared in double; hdt2 = 0.5f*dt2
    float mu, ts, bound;          // friction, target speed, 4*ts^2
    float ebound;                 // the largest float e with e*e <= bound in float32: |e| <= ebound <=> e*e <= bound
    float thr_lo, thr_w, fshape;  // fence ramp: |x| in [thr_lo, thr_lo + thr_w], shape = 5/width
    float turn;                   // 5*0.13 start angular velocity
    float lane_x[OCD_MAX_LANES];
    float lane_mid[OCD_MAX_LANES];  // midpoints between lanes adjacent in sorted order: where the lane-min ties
    float lane_sorted[OCD_MAX_LANES];   // lane positions in ascending order
    float fs_lo, fs_w;            // fence ramp in shape-scaled units: fshape*thr_lo, fshape*thr_w
    // the same in units of ln 2 (FAST kernels): with q' = ln2 * shape * q the argument of the ramp's exponential comes
    // out in base 2 (1/q' - 1/u' = log2(e) (1/qs - 1/us)): fl_shape = ln2 fshape, fl_lo = ln2 fs_lo, fl_w = ln2 fs_w,
    // and dT/dq = T (1-T) fl_c (r1'^2 + r2'^2) with fl_c = fshape ln2^2
    float fl_shape, fl_lo, fl_w, fl_c;
    float mid_c, mid_h;           // three lanes: centre and half-distance of the two midpoints (|x - mid_c| = mid_h at either)
};

// collision bump half-widths (merging.py:70-73) and their reciprocals
#define OCD_BUMP_HX 0.08f
#define OCD_BUMP_HY 0.15f
#define OCD_BUMP_IX 12.5f
#define OCD_BUMP_IY 6.6666667f
#define OCD_LOG2E   1.4426950408889634f
#define OCD_FULL    0xffffffffu

// ---------------------------------------------------------------------------------------------
// Math back-ends.
// ---------------------------------------------------------------------------------------------
template <bool PRECISE>
struct Mth;

template <>
struct Mth<false> {
    static __device__ __forceinline__ void sincos_(float th, float &s, float &c) {
        s = __sinf(th);
        c = __cosf(th);
    }
    static __device__ __forceinline__ float rcp_(float x) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float ex2_(float x) {   // 2^x
        float r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float exp_(float x) { return ex2_(x * OCD_LOG2E); }
    static __device__ __forceinline__ float div_(float a, float b) { return a * rcp_(b); }
};

template <>
struct Mth<true> {
    static __device__ __forceinline__ void sincos_(float th, float &s, float &c) {
        s = sinf(th);
        c = cosf(th);
    }
    static __device__ __forceinline__ float rcp_(float x) { return __fdiv_rn(1.0f, x); }
    static __device__ __forceinline__ float exp_(float x) { return expf(x); }
    static __device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }
};

// Other cars' positions are kept in a per-block slab.  PRECISE stores them raw; FAST stores them
// already divided by the bump half-width, so that the normalised offset is one FFMA.
template <bool PRECISE>
__device__ __forceinline__ float slab_x(float ox) { return PRECISE ? ox : ox * OCD_BUMP_IX; }
template <bool PRECISE>
__device__ __forceinline__ float slab_y(float oy) { return PRECISE ? oy : oy * OCD_BUMP_IY; }

// ---------------------------------------------------------------------------------------------
// a1  car_dynamics_step (simulation_utils.py:9-21), one car.
// ---------------------------------------------------------------------------------------------
template <bool PRECISE>
__device__ __forceinline__ void dynamics_step(float &x, float &y, float &v, float &th, float a, float om,
                                              float dt, float dt2, float mu) {
    const float ac = fmaxf(fminf(a, 4.0f), -8.0f);
    const float oc = fmaxf(fminf(om, 4.0f), -4.0f);
    float sn, cs;
    Mth<PRECISE>::sincos_(th, sn, cs);
    if (PRECISE) {   // one rounding per reference op
        const float total = __fsub_rn(ac, __fmul_rn(mu, __fmul_rn(v, v)));
        const float dist = __fadd_rn(__fmul_rn(v, dt), __fmul_rn(__fmul_rn(0.5f, total), dt2));
        x = __fadd_rn(x, __fmul_rn(cs, dist));
        y = __fadd_rn(y, __fmul_rn(sn, dist));
        v = __fadd_rn(v, __fmul_rn(total, dt));
        th = __fadd_rn(th, __fmul_rn(oc, dt));
    } else {
        const float total = fmaf(-mu, v * v, ac);
        const float dist = fmaf(total, 0.5f * dt2, v * dt);
        x = fmaf(cs, dist, x);
        y = fmaf(sn, dist, y);
        v = fmaf(total, dt, v);
        th = fmaf(oc, dt, th);
    }
}

// Planner's model of another car (naive_planner.py:53-66): frictionless, unclipped.
template <bool PRECISE>
__device__ __forceinline__ void other_model_step(float &x, float &y, float &v, float &th, bool known,
                                                 float a, float om, float dt, float dt2) {
    float sn, cs;
    Mth<PRECISE>::sincos_(th, sn, cs);
    if (known) {
        const float dist = __fadd_rn(__fmul_rn(v, dt), __fmul_rn(__fmul_rn(0.5f, a), dt2));
        x = __fadd_rn(x, __fmul_rn(cs, dist));
        y = __fadd_rn(y, __fmul_rn(sn, dist));
        v = __fadd_rn(v, __fmul_rn(a, dt));
        th = __fadd_rn(th, __fmul_rn(om, dt));
    } else {
        x = __fadd_rn(x, __fmul_rn(__fmul_rn(cs, v), dt));
        y = __fadd_rn(y, __fmul_rn(__fmul_rn(sn, v), dt));
    }
}

// ---------------------------------------------------------------------------------------------
// Smooth helpers (math_utils.py).  Each returns the value and its derivative.
// ---------------------------------------------------------------------------------------------
// smooth_bump(c - hw, c + hw)(z) as a function of the normalised offset n = (z - c)/hw:
// value b and db/dn.
template <bool PRECISE>
__device__ __forceinline__ void bump_n(float n, float &b, float &dbdn) {
    const float om = fmaf(-n, n, 1.0f);
    b = 0.0f;
    dbdn = 0.0f;
    if (n * n < 1.0f) {
        const float r = Mth<PRECISE>::rcp_(om);
        b = Mth<PRECISE>::exp_(1.0f - r);
        dbdn = b * (-2.0f * n) * (r * r);
    }
}

// Normalised offset of z from a slab coordinate.  PRECISE recomputes width and center from
// start/end in float32 like math_utils.py:169-171; FAST is one FFMA on the pre-scaled centre.
template <bool PRECISE, bool SCALED>
__device__ __forceinline__ float bump_offset(float z, float c, float hw, float inv_hw, float &inv_w) {
    if (PRECISE) {
        const float start = __fsub_rn(c, hw), end = __fadd_rn(c, hw);
        const float width = __fmul_rn(__fsub_rn(end, start), 0.5f);
        const float center = __fmul_rn(__fadd_rn(start, end), 0.5f);
        inv_w = __fdiv_rn(1.0f, width);
        return __fdiv_rn(__fsub_rn(z, center), width);
    }
    inv_w = inv_hw;
    return SCALED ? fmaf(z, inv_hw, -c) : (z - c) * inv_hw;
}

// fence(x) = (T(x) + T(-x)) * |x| with T = smooth_threshold(0.05*num_lanes, 0.05)
// (merging.py:80-81).  T(-|x|) is identically 0 (its _f argument is <= 0), so the feature is
// T(|x|)*|x|: 0 below the ramp, |x| above it, two exponentials only inside the ramp.
// Precondition for the caller: q = |x| - thr_lo > 0.
template <bool PRECISE>
__device__ __forceinline__ void fence_inside(const KParams &k, float x, float ax, float q, float &f, float &df) {
    const float u2 = k.thr_w - q;
    if (u2 > 0.0f) {
        const float r1 = Mth<PRECISE>::rcp_(k.fshape * q);
        const float r2 = Mth<PRECISE>::rcp_(k.fshape * u2);
        const float F1 = Mth<PRECISE>::exp_(-r1);
        const float F2 = Mth<PRECISE>::exp_(-r2);
        const float inv = Mth<PRECISE>::rcp_(F1 + F2);
        const float T = F1 * inv;
        // F'(q) = F(q)/(shape q^2) = F * shape * r^2
        const float dT = (F1 * F2) * (k.fshape * fmaf(r1, r1, r2 * r2)) * (inv * inv);
        f = T * ax;
        df = copysignf(fmaf(dT, ax, T), x);
    } else {
        f = ax;
        df = copysignf(1.0f, x);
    }
}

template <bool PRECISE>
__device__ __forceinline__ void fence_vg(const KParams &k, float x, float &f, float &df) {
    const float ax = fabsf(x);
    const float q = ax - k.thr_lo;
    f = 0.0f;
    df = 0.0f;
    if (q > 0.0f) fence_inside<PRECISE>(k, x, ax, q, f, df);
}

// Per-problem weights, pre-combined for the gradient:
//   sum_i w_i * d/dx[10 (x-l_i)^2] = GA*x + GB ;  w0x2 = 2 w_speed ; wmin20 = 20 w_min ;
//   wcx / wcy = collision weight times the bump's d n / d position and the -2 of bump'.
//   compile-time lane count: the nearest lane l* enters d/dx as wmin20 (x - l*), so with GAm = GA + wmin20
//   and one constant per lane GBl[i] = GB - wmin20 l_i (sorted order) the whole lane term is GAm x + GBl[*].
struct GradW {
    float w0x2, GA, GB, wmin20, wcol, wfence, wcx, wcy;
    float GAm, GBl[OCD_MAX_LANES];
};

template <int LT>
__device__ __forceinline__ GradW make_gradw(const KParams &k, const float *w /*[K] stride ws*/, int ws) {
    const int L = LT > 0 ? LT : k.L;
    GradW g;
    g.w0x2 = 2.0f * w[0];
    float sa = 0.0f, sb = 0.0f;
#pragma unroll
    for (int i = 0; i < (LT > 0 ? LT : OCD_MAX_LANES); ++i)
        if (i < L) {
            const float wi = w[(1 + i) * ws];
            sa += wi;
            sb = fmaf(wi, k.lane_x[i], sb);
        }
    g.GA = 20.0f * sa;
    g.GB = -20.0f * sb;
    g.wmin20 = 20.0f * w[(1 + L) * ws];
    g.wcol = w[(2 + L) * ws];
    g.wfence = w[(3 + L) * ws];
    g.wcx = g.wcol * (-2.0f * OCD_BUMP_IX);
    g.wcy = g.wcol * (-2.0f * OCD_BUMP_IY);
    g.GAm = g.GA + g.wmin20;
#pragma unroll
    for (int i = 0; i < OCD_MAX_LANES; ++i) g.GBl[i] = (i < L) ? fmaf(-g.wmin20, k.lane_sorted[i], g.GB) : 0.0f;
    if (LT > 0) {
        asm volatile("" : "+f"(g.GAm), "+f"(g.GBl[0]), "+f"(g.GBl[1]));
        if (LT > 2) asm volatile("" : "+f"(g.GBl[2]));
        if (LT > 3) asm volatile("" : "+f"(g.GBl[3]));
    }
    // keep the combined constants in registers: without the barrier ptxas rematerialises GA/GB from
    // the raw weights inside the hot loop (4 extra instructions per horizon step)
    asm volatile("" : "+f"(g.GA), "+f"(g.GB), "+f"(g.wmin20), "+f"(g.w0x2));
    asm volatile("" : "+f"(g.wcx), "+f"(g.wcy), "+f"(g.wfence), "+f"(g.wcol));
    return g;
}

// d/dx of min_i 10 (x - l_i)^2, divided by 20: the offset to the nearest lane, averaged over
// exact ties of the feature values (TF reduce_min splits the gradient evenly).
template <int LT>
__device__ __forceinline__ float lane_min_offset_exact(const KParams &k, float x) {
    const int L = LT > 0 ? LT : k.L;
    float dsel = x - k.lane_x[0];
    float fm = (dsel * dsel) * 10.0f;
    bool tie = false;
#pragma unroll
    for (int i = 1; i < (LT > 0 ? LT : OCD_MAX_LANES); ++i)
        if (i < L) {
            const float d = x - k.lane_x[i];
            const float f = (d * d) * 10.0f;
            tie = tie || (f == fm);
            dsel = (f < fm) ? d : dsel;
            fm = fminf(fm, f);
        }
    if (tie) {   // rare: an earlier tie may have been beaten later, so recount against the minimum
        float sum = 0.0f, cnt = 0.0f;
        for (int i = 0; i < L; ++i) {
            const float d = x - k.lane_x[i];
            if ((d * d) * 10.0f == fm) {
                sum += d;
                cnt += 1.0f;
            }
        }
        dsel = (cnt > 1.0f) ? __fdiv_rn(sum, cnt) : dsel;
    }
    return dsel;
}

// FAST: the nearest lane by |x - l_i| (no feature values needed); feature values can only tie when
// x sits within rounding of a midpoint between two neighbouring lanes, and only then -- decided by a
// warp vote, so normally one uniform branch -- the exact rule above runs.
// VM (vote mode) 1 -- the straight-line variant of the latency kernels -- only reports `close` through
// `flag`; the caller re-runs the step with the exact rule when any lane raised it.
template <int LT, bool PRECISE, int VM = 0>
__device__ __forceinline__ float lane_min_offset(const KParams &k, float x, bool &flag) {
    if (PRECISE) return lane_min_offset_exact<LT>(k, x);
    const int L = LT > 0 ? LT : k.L;
    float dsel = x - k.lane_x[0];
    float near = 1.0f;
#pragma unroll
    for (int i = 1; i < (LT > 0 ? LT : OCD_MAX_LANES); ++i)
        if (i < L) {
            const float d = x - k.lane_x[i];
            dsel = (fabsf(d) < fabsf(dsel)) ? d : dsel;
            near = fminf(near, fabsf(x - k.lane_mid[i - 1]));
        }
    const bool close = near < 1e-6f;
    if (VM == 1) {
        flag = flag || close;
        return dsel;
    }
    if (__any_sync(OCD_FULL, close)) {        // whole warp takes the exact rule: no divergence to reconverge
        const float exact = lane_min_offset_exact<LT>(k, x);
        dsel = close ? exact : dsel;
    }
    return dsel;
}

// Gradient of w.phi with respect to the robot state (x, y, v, th) at one world state.
//   oth: other cars' slab coordinates, x of car j at oth[j*jstride], y at oth[j*jstride + cstride].
// LIN (FAST, constant-velocity other cars only): instead of a position per horizon step the slab holds
// (x0, dx, y0, dy) per car -- four rows, car stride jstride = 4*cstride -- and the position after
// tf steps is x0 + tf*dx.  That keeps the slab independent of H, which is what lets long horizons with
// many cars keep several blocks per SM.
// VM = 1 (FAST only): no warp votes and no rare-path branches -- every block below is computed
// unconditionally (the clamped formulations make that safe), so the whole horizon is one basic block the
// scheduler can interleave; whatever needs an exact rule (lane tie, collision tie) raises `flag` instead.
// RAWG (every FAST kernel): return the FACTORS of three of the gradient components instead of the products --
// gv := ke (d/dv = ke sin th), gth := unused, gy := hy (d/dy = wcy hy).  The consumer adds each of them to an
// adjoint, and whether the compiler fuses that product and sum into one FMA depends on where the basic-block
// boundaries of the kernel form at hand fall (and is impossible behind the time-parallel kernels' shuffles): with
// the factors handed over, every form writes the same explicit fmaf and rounds the same way.
// FOLD: use the per-lane folded constants of GradW for the lane term (the register-resident kernels; the
// segmented kernels are short of registers and keep the three-instruction form).
template <int NOT_, int LT, bool PRECISE, bool LIN = false, int VM = 0, bool RAWG = !PRECISE, bool FOLD = !LIN>
__device__ __forceinline__ void feature_grad(const KParams &k, const GradW &w, float x, float y, float v,
                                             float sn, float cs, const float *oth, int jstride, int cstride,
                                             float &gx, float &gy, float &gv, float &gth, float tf, bool &flag) {
    const int NO = NOT_ > 0 ? NOT_ : k.NO;
    // ONE_RCP (register-resident kernels with one or two other cars): the collision bump and the fence each take
    // their two reciprocals from ONE MUFU (1/a, 1/b = b R, a R with R = 1/(a b)).  The XU pipe -- 8 cycles per
    // MUFU -- is the top stall of the straight-line forms, and three more FMULs for one MUFU less gain 2.4 % at the
    // bench shape (4.77 -> 4.66 ms); measured, it costs 1-4 % with more cars and in the segmented kernels.
    constexpr bool ONE_RCP = FOLD && !PRECISE && (NOT_ == 1 || NOT_ == 2);
    // speed: min((v sin th - ts)^2, 4 ts^2)                                  merging.py:58-59
    {
        const float e = fmaf(v, sn, -k.ts);
        const float ke = (fabsf(e) <= k.ebound) ? (w.w0x2 * e) : 0.0f;      // e*e <= bound, without the square
        gv = RAWG ? ke : ke * sn;
        gth = RAWG ? 0.0f : (ke * v) * cs;
    }
    // lanes: sum_i w_i 10 (x - l_i)^2 and the min over lanes                 merging.py:61-65
    if (LT > 0 && !PRECISE && FOLD) {
        // nearest lane by the midpoints of the sorted lanes (ascending: the last one passed wins); exact ties
        // of the feature values are only possible within rounding of a midpoint -- then (warp vote, or the
        // caller's redo in vote mode 1) the exact rule runs
        float gb = w.GBl[0], near = 1.0f;
#pragma unroll
        for (int i = 1; i < LT; ++i) {
            gb = (x > k.lane_mid[i - 1]) ? w.GBl[i] : gb;
            if (LT != 3) near = fminf(near, fabsf(x - k.lane_mid[i - 1]));
        }
        // three lanes: x is near one of the two midpoints  <=>  | |x - centre| - half-distance | is small (one
        // subtraction less than a distance to each)
        if (LT == 3) near = fabsf(fabsf(x - k.mid_c) - k.mid_h);
        gx = fmaf(w.GAm, x, gb);
        const bool close = near < 1e-6f;
        if (VM == 1) {
            flag = flag || close;
        } else if (__any_sync(OCD_FULL, close)) {
            const float exact = fmaf(w.wmin20, lane_min_offset_exact<LT>(k, x), fmaf(w.GA, x, w.GB));
            gx = close ? exact : gx;
        }
    } else {
        gx = fmaf(w.wmin20, lane_min_offset<LT, PRECISE, VM>(k, x, flag), fmaf(w.GA, x, w.GB));
    }
    // collision: max_j bump_x * bump_y                                        merging.py:67-78
    if (PRECISE) {
        float best = 0.0f, sx = 0.0f, sy = 0.0f, cnt = 1.0f;
#pragma unroll(NOT_ > 0 ? NOT_ : 1)
        for (int j = 0; j < NO; ++j) {
            float iwx, iwy, bx, dbx, by, dby;
            const float nx = bump_offset<true, false>(x, oth[j * jstride], OCD_BUMP_HX, OCD_BUMP_IX, iwx);
            const float ny = bump_offset<true, false>(y, oth[j * jstride + cstride], OCD_BUMP_HY, OCD_BUMP_IY, iwy);
            bump_n<true>(nx, bx, dbx);
            bump_n<true>(ny, by, dby);
            const float val = bx * by, vx = (dbx * iwx) * by, vy = bx * (dby * iwy);
            if (j == 0 || val > best) {
                best = val; sx = vx; sy = vy; cnt = 1.0f;
            } else if (val == best) {
                sx += vx; sy += vy; cnt += 1.0f;
            }
        }
        if (cnt != 1.0f) {
            sx = __fdiv_rn(sx, cnt);
            sy = __fdiv_rn(sy, cnt);
        }
        gx = fmaf(w.wcol, sx, gx);
        gy = w.wcol * sy;
    } else {
        // Both bumps share one exponential: bx*by = exp(2 - 1/(1-nx^2) - 1/(1-ny^2)).  Outside
        // the support 1-n^2 is clamped to a tiny positive number: the exponential underflows to
        // exactly 0 and takes every derivative with it, so no branch and no select is needed.
        // hx, hy below are d(val)/d(position) without the constant -2/half-width (folded into
        // wcx, wcy).  A warp with every lane outside every support skips the transcendental part.
        float hx = 0.0f, hy = 0.0f;
        if (NOT_ == 1 || NOT_ == 2) {
            // one or two other cars (the shipped scenarios): value and derivatives per car, TF's even
            // split among exact ties applied on the fly (the replanning scenario's two other cars start
            // on top of each other, so exact ties are the normal case there)
            float best = 0.0f, cnt = 1.0f;
#pragma unroll
            for (int j = 0; j < NOT_; ++j) {
                const float cxj = LIN ? fmaf(tf, oth[j * jstride + cstride], oth[j * jstride]) : oth[j * jstride];
                const float cyj = LIN ? fmaf(tf, oth[j * jstride + 3 * cstride], oth[j * jstride + 2 * cstride])
                                      : oth[j * jstride + cstride];
                const float nx = fmaf(x, OCD_BUMP_IX, -cxj);
                const float ny = fmaf(y, OCD_BUMP_IY, -cyj);
                const float ux = fmaf(-nx, nx, 1.0f), uy = fmaf(-ny, ny, 1.0f);
                float val = 0.0f, vx = 0.0f, vy = 0.0f;
                if (VM == 1 || __any_sync(OCD_FULL, fminf(ux, uy) > 0.0f)) {
                    const float uxc = fmaxf(ux, 1e-6f), uyc = fmaxf(uy, 1e-6f);
                    float rx, ry;
                    if (ONE_RCP) {                   // 1/ux and 1/uy from one MUFU
                        const float R = Mth<false>::rcp_(uxc * uyc);
                        rx = uyc * R;
                        ry = uxc * R;
                    } else {
                        rx = Mth<false>::rcp_(uxc);
                        ry = Mth<false>::rcp_(uyc);
                    }
                    val = Mth<false>::ex2_(fmaf(rx + ry, -OCD_LOG2E, 2.0f * OCD_LOG2E));
                    vx = (val * nx) * (rx * rx);
                    vy = (val * ny) * (ry * ry);
                }
                if (NOT_ == 1) {
                    hx = vx; hy = vy;
                } else if (j == 0 || val > best) {
                    best = val; hx = vx; hy = vy; cnt = 1.0f;
                } else if (val == best) {
                    hx += vx; hy += vy; cnt += 1.0f;
                }
            }
            if (NOT_ != 1) {                     // cnt is 1 or 2 here: 1/cnt without a division or a branch
                const float r = (cnt != 1.0f) ? 0.5f : 1.0f;
                hx *= r;
                hy *= r;
            }
        } else {
            // three or more other cars, or a runtime count (sweeps with many cars).  bx*by =
            // exp(2 - 1/ux - 1/uy) with u = 1 - n^2 is monotone in 1/ux + 1/uy = (ux + uy)/(ux uy), so the
            // maximum over cars is an argmin of that score: ONE reciprocal per car and no exponential; the
            // winner's offsets ride along through two selects and its value and derivatives are formed once,
            // after the loop (3 more MUFU).  Outside the support u is clamped to a tiny positive number -- a
            // different one per car, so that two far cars never produce equal scores -- and the score is
            // >= 1e5: the exponential underflows to exactly 0.  Fully unrolled (with uniform guards when the
            // count is a runtime value), so every slab address and clamp is an immediate.
            // Ties: an exact tie of two scores inside the support (cars at the same spot, or the robot dead
            // centre between two cars -- TF splits the gradient there) sets a sticky flag and sends the warp
            // through the exact rule below.
            float bsum = 3.0e38f, bnx = 0.0f, bny = 0.0f;
            bool tie = false;
#pragma unroll
            for (int j = 0; j < (NOT_ > 0 ? NOT_ : OCD_MAX_OTHER); ++j) {
                if (NOT_ == 0 && j >= NO) break;
                const float cxj = LIN ? fmaf(tf, oth[j * jstride + cstride], oth[j * jstride]) : oth[j * jstride];
                const float cyj = LIN ? fmaf(tf, oth[j * jstride + 3 * cstride], oth[j * jstride + 2 * cstride])
                                      : oth[j * jstride + cstride];
                const float nx = fmaf(x, OCD_BUMP_IX, -cxj);
                const float ny = fmaf(y, OCD_BUMP_IY, -cyj);
                const float eps = 1e-6f * (1.0f + 0.125f * (float)j);
                const float ux = fmaxf(fmaf(-nx, nx, 1.0f), eps), uy = fmaxf(fmaf(-ny, ny, 1.0f), eps);
                const float sum = (ux + uy) * Mth<false>::rcp_(ux * uy);
                tie = tie || (sum == bsum);
                const bool lt = sum < bsum;
                bsum = fminf(bsum, sum);
                bnx = lt ? nx : bnx;
                bny = lt ? ny : bny;
            }
            tie = tie && bsum < 1.0e5f;
            if (VM == 1 || __any_sync(OCD_FULL, bsum < 1.0e5f)) {
                const float rx = Mth<false>::rcp_(fmaxf(fmaf(-bnx, bnx, 1.0f), 1e-6f));
                const float ry = Mth<false>::rcp_(fmaxf(fmaf(-bny, bny, 1.0f), 1e-6f));
                const float best = Mth<false>::ex2_(fmaf(rx + ry, -OCD_LOG2E, 2.0f * OCD_LOG2E));
                hx = (best * bnx) * (rx * rx);
                hy = (best * bny) * (ry * ry);
            }
            if (VM == 1) flag = flag || tie;
            if (VM != 1 && __any_sync(OCD_FULL, tie)) {
                float b2 = 0.0f, sx = 0.0f, sy = 0.0f, cnt = 1.0f;
                for (int j = 0; j < NO; ++j) {
                    const float cxj = LIN ? fmaf(tf, oth[j * jstride + cstride], oth[j * jstride]) : oth[j * jstride];
                    const float cyj = LIN ? fmaf(tf, oth[j * jstride + 3 * cstride], oth[j * jstride + 2 * cstride])
                                          : oth[j * jstride + cstride];
                    const float nx = fmaf(x, OCD_BUMP_IX, -cxj), ny = fmaf(y, OCD_BUMP_IY, -cyj);
                    const float rx = Mth<false>::rcp_(fmaxf(fmaf(-nx, nx, 1.0f), 1e-6f));
                    const float ry = Mth<false>::rcp_(fmaxf(fmaf(-ny, ny, 1.0f), 1e-6f));
                    const float val = Mth<false>::ex2_(fmaf(rx + ry, -OCD_LOG2E, 2.0f * OCD_LOG2E));
                    const float vx = (val * nx) * (rx * rx), vy = (val * ny) * (ry * ry);
                    if (j == 0 || val > b2) {
                        b2 = val; sx = vx; sy = vy; cnt = 1.0f;
                    } else if (val == b2) {
                        sx += vx; sy += vy; cnt += 1.0f;
                    }
                }
                const float r = __fdiv_rn(1.0f, cnt);
                hx = tie ? sx * r : hx;
                hy = tie ? sy * r : hy;
            }
        }
        gx = fmaf(w.wcx, hx, gx);
        gy = RAWG ? hy : w.wcy * hy;
    }
    // fence                                                                    merging.py:80-81
    {
        const float ax = fabsf(x);
        if (PRECISE) {
            const float q = ax - k.thr_lo;
            if (q > 0.0f) {
                float f, df;
                fence_inside<true>(k, x, ax, q, f, df);
                gx = fmaf(w.wfence, df, gx);
            }
        } else if (const float qs = fmaf(ax, k.fl_shape, -k.fl_lo); VM == 1 || __any_sync(OCD_FULL, qs > 0.0f)) {
            // in units of ln2 * shape (qs = ln2 shape q): T = F1/(F1+F2) = 1/(1 + 2^(r1 - r2)), r1 = 1/qs,
            // r2 = 1/(ln2 shape width - qs); dT/dq = T (1-T) fl_c (r1^2 + r2^2).  Clamping qs and its complement to a
            // tiny positive number makes the exponential saturate: T = 0, dT = 0 below the ramp and T = 1, dT = 0
            // above it, so the three regions need no branch.
            const float qc = fmaxf(qs, 1e-7f), uc = fmaxf(k.fl_w - qs, 1e-7f);
            float r1, r2;
            if (ONE_RCP) {     // one reciprocal serves both: with R = 1/(q u), r1 = u R and r2 = q R
                const float R = Mth<false>::rcp_(qc * uc);
                r1 = uc * R;
                r2 = qc * R;
            } else {
                r1 = Mth<false>::rcp_(qc);
                r2 = Mth<false>::rcp_(uc);
            }
            const float rr = fmaf(r1, r1, r2 * r2);
            const float T = Mth<false>::rcp_(1.0f + Mth<false>::ex2_(r1 - r2));
            const float dT = fmaf(-T, T, T) * (k.fl_c * rr);          // T (1 - T) as one FMA
            gx = fmaf(w.wfence, copysignf(fmaf(dT, ax, T), x), gx);
        }
    }
}

// Feature vector phi[K] at one world state, in the reference's order and op order.
// SCALED: the other cars' coordinates come from a FAST slab (already divided by the half-width).
template <int LT, bool PRECISE, bool SCALED, bool LIN = false>
__device__ __forceinline__ void feature_values(const KParams &k, float x, float y, float v, float sn,
                                               const float *oth, int jstride, int cstride,
                                               float (&phi)[OCD_MAX_LANES + 4], float tf = 0.0f) {
    const int L = LT > 0 ? LT : k.L;
    {
        const float e = __fsub_rn(__fmul_rn(v, sn), k.ts);
        phi[0] = fminf(__fmul_rn(e, e), k.bound);
    }
    float fmin_ = 0.0f;
#pragma unroll
    for (int i = 0; i < (LT > 0 ? LT : OCD_MAX_LANES); ++i)
        if (i < L) {
            const float dx = x - k.lane_x[i];
            const float f = __fmul_rn(__fmul_rn(dx, dx), 10.0f);
            phi[1 + i] = f;
            fmin_ = (i == 0) ? f : fminf(fmin_, f);
        }
    float best = 0.0f;
    for (int j = 0; j < k.NO; ++j) {
        float iw, bx, dbx, by, dby;
        const float cxj = LIN ? fmaf(tf, oth[j * jstride + cstride], oth[j * jstride]) : oth[j * jstride];
        const float cyj = LIN ? fmaf(tf, oth[j * jstride + 3 * cstride], oth[j * jstride + 2 * cstride])
                              : oth[j * jstride + cstride];
        bump_n<PRECISE>(bump_offset<PRECISE, SCALED>(x, cxj, OCD_BUMP_HX, OCD_BUMP_IX, iw), bx, dbx);
        bump_n<PRECISE>(bump_offset<PRECISE, SCALED>(y, cyj, OCD_BUMP_HY, OCD_BUMP_IY, iw), by, dby);
        const float val = __fmul_rn(bx, by);
        best = (j == 0) ? val : fmaxf(best, val);
    }
    float fen, dfen;
    fence_vg<PRECISE>(k, x, fen, dfen);
    // write the tail with static indices only (phi stays in registers)
#pragma unroll
    for (int i = 1; i <= OCD_MAX_LANES; ++i)
        if (i == L) {
            phi[1 + i] = fmin_;
            if (i + 2 < OCD_MAX_LANES + 4) phi[2 + i] = best;
            if (i + 3 < OCD_MAX_LANES + 4) phi[3 + i] = fen;
        }
}

// w . phi summed in feature order (linear_reward_car.py:53).  w[k] at w[k*ws].
template <int LT, bool PRECISE, bool SCALED, bool LIN = false>
__device__ __forceinline__ float reward_value(const KParams &k, const float *w, int ws, float x, float y,
                                              float v, float sn, const float *oth, int jstride, int cstride,
                                              float tf = 0.0f) {
    float phi[OCD_MAX_LANES + 4];
    feature_values<LT, PRECISE, SCALED, LIN>(k, x, y, v, sn, oth, jstride, cstride, phi, tf);
    float r = 0.0f;
#pragma unroll
    for (int i = 0; i < OCD_MAX_LANES + 4; ++i)
        if (i < k.K) r = __fadd_rn(r, __fmul_rn(w[i * ws], phi[i]));
    return r;
}

// ---------------------------------------------------------------------------------------------
// One (problem, start): forward rollout + feature gradients, reverse adjoint sweep, SGD update.
//   HT > 0: horizon known at compile time, everything unrolled into registers.
//   HT == 0: runtime horizon (<= OCD_MAX_H), per-step arrays in local memory.
// `oth` points at this problem's column of the block's other-car slab laid out
// [H][NO][2][P]: coordinate c of car j at step t is oth[((t*NO + j)*2 + c)*P].
// ---------------------------------------------------------------------------------------------
template <int HT>
struct Traj {
    static constexpr int HM = HT > 0 ? HT : OCD_MAX_H;
    float ua[HM], uw[HM];                      // controls (acceleration, angular velocity)
    __device__ __forceinline__ float acc(int t) const { return ua[t]; }
    __device__ __forceinline__ float ang(int t) const { return uw[t]; }
};

// Controls of one thread kept in shared memory (used by the segmented solver for runtime / long
// horizons, where 2H controls do not fit the register file).  Each thread owns 2H consecutive floats; rows are
// 2 x (odd) floats apart (seg_u_stride), so the 64-bit accesses of a warp to the same step are conflict-free,
// while inside a segment every access is base + immediate.
struct SmemTraj {
    float *p;      // this thread's controls: (acc_t, ang_t) at p[2t], p[2t+1]
    __device__ __forceinline__ float acc(int t) const { return p[2 * t]; }
    __device__ __forceinline__ float ang(int t) const { return p[2 * t + 1]; }
    __device__ __forceinline__ void set(int t, float a, float w) const {
        p[2 * t] = a;
        p[2 * t + 1] = w;
    }
};
// Row strides (in floats) of a thread's controls and checkpoints.  A control pair is read and written as one
// 64-bit access, a checkpoint as one 128-bit access, so rows are 8 / 16 bytes aligned and the stride is 2 x odd /
// 4 x odd: the 16 (8) threads such an access is split over then hit 32 different banks.
__host__ __device__ inline int seg_u_stride(int H) { return 2 * (H | 1); }
// one checkpoint per segment but the first (which starts at the initial state); never zero floats
__host__ __device__ inline int seg_ck_stride(int H, int SEG) {
    const int nseg = (H + SEG - 1) / SEG;
    return 4 * ((nseg > 1 ? nseg - 1 : 1) | 1);
}

// Forward half of one iteration: roll the robot out and, at every new state, take the gradient of w.phi.
// Fills the saved per-step values the reverse sweep needs.  VM as in feature_grad.
// SF (step fence; the wide form with three or more other cars): an opaque, always-true branch at the top of every
// step keeps the steps in separate basic blocks.  The cars of one step still interleave, but the scheduler no
// longer overlaps the five steps, which with many cars needs more registers than there are (544 bytes of spill
// at 128 registers, none with the fence: 8.45 vs 11.0 ms with six cars, and 9.57 ms for the throughput form).
template <int HT, int NOT_, int LT, bool PRECISE, int VM, bool SF = false>
__device__ __forceinline__ void forward_sweep(const KParams &k, const GradW &w, float x0, float y0, float v0,
                                              float th0, float sn0, float cs0, const float *oth, int P,
                                              const Traj<HT> &u, float *sv, float *sc, float *ss, float *sd,
                                              float *gx, float *gy, float *gv, float *gth, bool &flag) {
    const int H = HT > 0 ? HT : k.H;
    const int NO = NOT_ > 0 ? NOT_ : k.NO;
    float x = x0, y = y0, v = v0, th = th0, sn = sn0, cs = cs0;
#pragma unroll(HT > 0 ? HT : 1)
    for (int t = 0; t < H; ++t) {
        if (SF) {
            int one;
            asm volatile("mov.u32 %0, 1;" : "=r"(one));
            if (one == 0) continue;
        }
        const float ac = fmaxf(fminf(u.ua[t], 4.0f), -8.0f);
        const float oc = fmaxf(fminf(u.uw[t], 4.0f), -4.0f);
        const float total = fmaf(-k.mu, v * v, ac);
        const float dist = fmaf(total, k.hdt2, v * k.dt);
        sv[t] = v; sc[t] = cs; ss[t] = sn; sd[t] = dist;
        x = fmaf(cs, dist, x);
        y = fmaf(sn, dist, y);
        v = fmaf(total, k.dt, v);
        th = fmaf(oc, k.dt, th);
        Mth<PRECISE>::sincos_(th, sn, cs);
        feature_grad<NOT_, LT, PRECISE, false, VM>(k, w, x, y, v, sn, cs, oth + (size_t)t * NO * 2 * P, 2 * P, P,
                                                   gx[t], gy[t], gv[t], gth[t], 0.0f, flag);
    }
    sv[H] = v; sc[H] = cs; ss[H] = sn;          // the state after the last step (sv/sc/ss[t] = before step t)
}

// LAT: the latency variant for small batches (few warps per SM, nothing to hide latency with): the forward
// sweep is straight-line code -- no votes, no rare-path branches -- and is simply run again with the exact
// rules in the rare case that some lane needed one.
template <int HT, int NOT_, int LT, bool PRECISE, bool UPDATE, int LAT = 0>
__device__ __forceinline__ void sgd_iteration(const KParams &k, const GradW &w, float x0, float y0, float v0,
                                              float th0, float sn0, float cs0, const float *oth, int P,
                                              Traj<HT> &u, float *ga_out, float *gw_out) {
    constexpr int HM = Traj<HT>::HM;
    const int H = HT > 0 ? HT : k.H;
    float sv[HM + 1], sc[HM + 1], ss[HM + 1], sd[HM];   // saved v_t, cos th_t, sin th_t (t = 0..H), d_t
    float gx[HM], gy[HM], gv[HM], gth[HM];     // feature gradient at s_{t+1} (FAST: gy, gv hold factors, see RAWG)
    bool flag = false;
    if (LAT != 0 && !PRECISE) {
        forward_sweep<HT, NOT_, LT, PRECISE, 1, (LAT == 2 && NOT_ >= 3)>(k, w, x0, y0, v0, th0, sn0, cs0, oth, P, u, sv, sc,
                                                                         ss, sd, gx, gy, gv, gth, flag);
        if (__any_sync(OCD_FULL, flag))
            forward_sweep<HT, NOT_, LT, PRECISE, 0>(k, w, x0, y0, v0, th0, sn0, cs0, oth, P, u, sv, sc, ss, sd, gx, gy,
                                                    gv, gth, flag);
    } else {
        forward_sweep<HT, NOT_, LT, PRECISE, 0>(k, w, x0, y0, v0, th0, sn0, cs0, oth, P, u, sv, sc, ss, sd, gx, gy, gv,
                                                gth, flag);
    }
    float lx = 0.0f, ly = 0.0f, lv = 0.0f, lth = 0.0f;
    const float c1 = -2.0f * k.mu * k.dt, c2 = -k.mu * k.dt2;
    const float lra = k.lr * k.hdt2, lrv = k.lr * k.dt;
#pragma unroll(HT > 0 ? HT : 1)
    for (int tt = 0; tt < H; ++tt) {
        const int t = H - 1 - tt;
        const float mx = gx[t] + lx;
        const float my = PRECISE ? gy[t] + ly : fmaf(w.wcy, gy[t], ly);
        const float mv = PRECISE ? gv[t] + lv : fmaf(gv[t], ss[t + 1], lv);
        const float mth = PRECISE ? gth[t] + lth : fmaf(__fmul_rn(gv[t], sv[t + 1]), sc[t + 1], lth);
        const float ld = fmaf(sc[t], mx, ss[t] * my);
        const float a = u.ua[t], om = u.uw[t];
        const bool in_a = (a >= -8.0f) && (a <= 4.0f);      // d clip / d a, inclusive (TF masks)
        const bool in_w = fabsf(om) <= 4.0f;
        lv = fmaf(fmaf(c1, sv[t], 1.0f), mv, fmaf(c2, sv[t], k.dt) * ld);
        lth = fmaf(sd[t], fmaf(sc[t], my, -(ss[t] * mx)), mth);
        lx = mx;
        ly = my;
        if (UPDATE && !PRECISE) {               // u <- u + lr * dR/du, constants folded
            u.ua[t] = in_a ? fmaf(lra, ld, fmaf(lrv, mv, a)) : a;
            u.uw[t] = in_w ? fmaf(lrv, mth, om) : om;
        } else {
            const float ga = in_a ? fmaf(k.hdt2, ld, k.dt * mv) : 0.0f;
            const float gw = in_w ? k.dt * mth : 0.0f;
            if (UPDATE) {                       // u <- u - lr * d(-R)/du          naive_planner.py:153
                u.ua[t] = fmaf(k.lr, ga, a);
                u.uw[t] = fmaf(k.lr, gw, om);
            } else {
                ga_out[t] = ga;
                gw_out[t] = gw;
            }
        }
    }
}

// R(u) = sum_t w . phi(s_{t+1}), value only, reference op order (naive_planner.py:43-77).
// The slab is a FAST (scaled) one exactly when PRECISE is false.
template <int HT, int LT, bool PRECISE, typename Controls, bool LIN = false>
__device__ __forceinline__ float rollout_reward(const KParams &k, const float *wraw, int ws, float x0, float y0,
                                                float v0, float th0, const float *oth, int P,
                                                const Controls &u) {
    const int H = HT > 0 ? HT : k.H;
    float x = x0, y = y0, v = v0, th = th0;
    float r = 0.0f;
#pragma unroll(HT > 0 ? HT : 1)
    for (int t = 0; t < H; ++t) {
        dynamics_step<PRECISE>(x, y, v, th, u.acc(t), u.ang(t), k.dt, k.dt2, k.mu);
        float sn, cs;
        Mth<PRECISE>::sincos_(th, sn, cs);
        r = __fadd_rn(r, LIN ? reward_value<LT, PRECISE, !PRECISE, LIN>(k, wraw, ws, x, y, v, sn, oth, 4 * P, P,
                                                                         (float)(t + 1))
                             : reward_value<LT, PRECISE, !PRECISE, false>(k, wraw, ws, x, y, v, sn,
                                                                          oth + (size_t)t * k.NO * 2 * P, 2 * P, P));
    }
    return r;
}

// Start s of the multi-start set (naive_planner.py:107-118).
template <int HT>
__device__ __forceinline__ void init_start(const KParams &k, int s, float cur_speed, Traj<HT> &u) {
    const int H = HT > 0 ? HT : k.H;
    const float a0 = (s >= 3) ? __fmul_rn(k.mu, __fmul_rn(cur_speed, cur_speed)) : 0.0f;
    const int m = s % 3;
    const float w0 = (m == 0) ? 0.0f : ((m == 1) ? -k.turn : k.turn);
#pragma unroll(HT > 0 ? HT : 1)
    for (int t = 0; t < H; ++t) {
        u.ua[t] = a0;
        u.uw[t] = w0;
    }
}

// The complete solve for one (problem, start): n_iter SGD iterations, then the final loss.
template <int HT, int NOT_, int LT, bool PRECISE, int LAT = 0>
__device__ __forceinline__ float solve_start(const KParams &k, const GradW &w, const float *wraw, int ws,
                                             float x0, float y0, float v0, float th0, const float *oth, int P,
                                             Traj<HT> &u) {
    float sn0, cs0;
    Mth<PRECISE>::sincos_(th0, sn0, cs0);
#pragma unroll 1
    for (int it = 0; it < k.n_iter; ++it)
        sgd_iteration<HT, NOT_, LT, PRECISE, true, LAT>(k, w, x0, y0, v0, th0, sn0, cs0, oth, P, u, nullptr, nullptr);
    return -rollout_reward<HT, LT, PRECISE, Traj<HT>>(k, wraw, ws, x0, y0, v0, th0, oth, P, u);
}

// ---------------------------------------------------------------------------------------------
// Medium compile-time horizons (HT = 9..24, FAST): the register file holds the controls and the saved
// (v, cos, sin) of every step, but not the four other values the reverse sweep needs per step -- the step
// length d_t and the feature gradient (gx, hy, ke).  Those four travel through shared memory as ONE float4 per
// step: one STS.128 in the forward sweep, one LDS.128 in the reverse sweep (row of HT float4 per thread, rows an
// odd number of float4 apart: conflict-free quarter-warps, base + immediate addressing).  Nothing is recomputed
// -- the segmented kernels pay a second dynamics pass with its two MUFU per step -- and the arithmetic is the
// register-resident kernels' own, statement for statement.
// ---------------------------------------------------------------------------------------------
#ifndef OCD_Q_SF
#define OCD_Q_SF 1
#endif
__host__ __device__ inline int q_stride(int H) { return H | 1; }     // float4 per thread row

// OCD_Q_CS = 1 (tuning): the saved (cos, sin) of every step travel through shared memory too (one float2 per step,
// rows of HT + 1 float2 behind the float4 rows), which leaves only v_t and the controls in registers.
#ifndef OCD_Q_CS
#define OCD_Q_CS 1
#endif
__host__ __device__ inline int q2_stride(int H) { return (H - 1) | 1; }                 // float2 per thread row: steps 1 .. H-1 (entry t at index t-1)
__host__ __device__ inline int q_thread_floats(int H) { return 4 * q_stride(H) + (OCD_Q_CS ? 2 * q2_stride(H) : 0); }

template <int HT, int NOT_, int LT, int VM, bool SF, bool LIN = false>
__device__ __forceinline__ void forward_sweep_q(const KParams &k, const GradW &w, float x0, float y0, float v0,
                                                float th0, float sn0, float cs0, const float *oth, int P,
                                                const Traj<HT> &u, float *sv, float *sc, float *ss, float4 *q,
                                                float2 *q2, bool &flag) {
    constexpr int NO = NOT_;
    float x = x0, y = y0, v = v0, th = th0, sn = sn0, cs = cs0;
#pragma unroll
    for (int t = 0; t < HT; ++t) {
        if (SF && t % (OCD_Q_SF > 0 ? OCD_Q_SF : 1) == 0) {     // a fence every OCD_Q_SF steps: groups of that many steps interleave
            int one;
            asm volatile("mov.u32 %0, 1;" : "=r"(one));
            if (one == 0) continue;
        }
        const float ac = fmaxf(fminf(u.ua[t], 4.0f), -8.0f);
        const float oc = fmaxf(fminf(u.uw[t], 4.0f), -4.0f);
        const float total = fmaf(-k.mu, v * v, ac);
        const float dist = fmaf(total, k.hdt2, v * k.dt);
        sv[t] = v;
        if (OCD_Q_CS) {
            if (t > 0) q2[t] = make_float2(cs, sn);           // (cos, sin) of th_0 stay in registers
        } else {
            sc[t] = cs; ss[t] = sn;
        }
        x = fmaf(cs, dist, x);
        y = fmaf(sn, dist, y);
        v = fmaf(total, k.dt, v);
        th = fmaf(oc, k.dt, th);
        Mth<false>::sincos_(th, sn, cs);
        float gx, hy, ke, unused;
        if (LIN)        // the slab holds (x0, dx, y0, dy) per other car: the position after t+1 steps is x0 + (t+1) dx
            feature_grad<NOT_, LT, false, true, VM, true, true>(k, w, x, y, v, sn, cs, oth, 4 * P, P, gx, hy, ke, unused,
                                                                (float)(t + 1), flag);
        else
            feature_grad<NOT_, LT, false, false, VM>(k, w, x, y, v, sn, cs, oth + (size_t)t * NO * 2 * P, 2 * P, P, gx, hy,
                                                     ke, unused, 0.0f, flag);
        q[t] = make_float4(dist, gx, hy, ke);
    }
    sv[HT] = v; sc[HT] = cs; ss[HT] = sn;
}

template <int HT, int NOT_, int LT, int LAT, bool LIN = false>
__device__ __forceinline__ void sgd_iteration_q(const KParams &k, const GradW &w, float x0, float y0, float v0,
                                                float th0, float sn0, float cs0, const float *oth, int P,
                                                Traj<HT> &u, float4 *q, float2 *q2) {
    float sv[HT + 1], sc[HT + 1], ss[HT + 1];
    bool flag = false;
    if (LAT != 0) {
        forward_sweep_q<HT, NOT_, LT, 1, (OCD_Q_SF != 0 || (LAT == 2 && NOT_ >= 3)), LIN>(k, w, x0, y0, v0, th0, sn0, cs0, oth,
                                                                                         P, u, sv, sc, ss, q, q2, flag);
        if (__any_sync(OCD_FULL, flag))
            forward_sweep_q<HT, NOT_, LT, 0, false, LIN>(k, w, x0, y0, v0, th0, sn0, cs0, oth, P, u, sv, sc, ss, q, q2, flag);
    } else {
        forward_sweep_q<HT, NOT_, LT, 0, false, LIN>(k, w, x0, y0, v0, th0, sn0, cs0, oth, P, u, sv, sc, ss, q, q2, flag);
    }
    float lx = 0.0f, ly = 0.0f, lv = 0.0f, lth = 0.0f;
    const float c1 = -2.0f * k.mu * k.dt, c2 = -k.mu * k.dt2;
    const float lra = k.lr * k.hdt2, lrv = k.lr * k.dt;
    float cn = sc[HT], snn = ss[HT];                       // (cos, sin) of th_{t+1}
#pragma unroll
    for (int tt = 0; tt < HT; ++tt) {
        const int t = HT - 1 - tt;
        const float4 g = q[t];                               // (d_t, gx, hy, ke) at s_{t+1}
        float ct, st;                                        // (cos, sin) of th_t
        if (OCD_Q_CS) {
            if (t > 0) {
                const float2 c = q2[t];
                ct = c.x; st = c.y;
            } else {
                ct = cs0; st = sn0;
            }
        } else {
            ct = sc[t]; st = ss[t];
        }
        const float mx = g.y + lx;
        const float my = fmaf(w.wcy, g.z, ly);
        const float mv = fmaf(g.w, snn, lv);
        const float mth = fmaf(__fmul_rn(g.w, sv[t + 1]), cn, lth);
        const float ld = fmaf(ct, mx, st * my);
        const float a = u.ua[t], om = u.uw[t];
        const bool in_a = (a >= -8.0f) && (a <= 4.0f);
        const bool in_w = fabsf(om) <= 4.0f;
        lv = fmaf(fmaf(c1, sv[t], 1.0f), mv, fmaf(c2, sv[t], k.dt) * ld);
        lth = fmaf(g.x, fmaf(ct, my, -(st * mx)), mth);
        lx = mx;
        ly = my;
        cn = ct; snn = st;
        u.ua[t] = in_a ? fmaf(lra, ld, fmaf(lrv, mv, a)) : a;
        u.uw[t] = in_w ? fmaf(lrv, mth, om) : om;
    }
}

template <int HT, int NOT_, int LT, int LAT, bool LIN = false>
__device__ __forceinline__ float solve_start_q(const KParams &k, const GradW &w, const float *wraw, int ws, float x0,
                                               float y0, float v0, float th0, const float *oth, int P, Traj<HT> &u,
                                               float4 *q, float2 *q2) {
    float sn0, cs0;
    Mth<false>::sincos_(th0, sn0, cs0);
#pragma unroll 1
    for (int it = 0; it < k.n_iter; ++it) {
#if defined(OCD_Q_SYNC) && OCD_Q_SYNC
        __syncthreads();        // keep the block's warps on the same stretch of the unrolled sweep (instruction fetch)
#endif
        sgd_iteration_q<HT, NOT_, LT, LAT, LIN>(k, w, x0, y0, v0, th0, sn0, cs0, oth, P, u, q, q2);
    }
    return -rollout_reward<HT, LT, false, Traj<HT>, LIN>(k, wraw, ws, x0, y0, v0, th0, oth, P, u);
}

// ---------------------------------------------------------------------------------------------
// Time-parallel solve: the lowest-latency path, for batches so small that even the latency variant
// leaves the GPU idle (the reference's real configurations: 45-180 episodes).  TG consecutive lanes share one
// (problem, start) -- TG = 8 for horizons up to 8, 16 for horizons up to 16 (tp_lanes) -- and lane t owns horizon
// step t (HT <= TG; spare lanes shadow the last step).  Per iteration the lanes all-gather the controls (2 HT
// shuffles), every lane rolls the cheap dynamics for the whole horizon, evaluates the expensive feature gradient for
// ITS step only, the gradients are all-gathered (3 HT shuffles) and every lane runs the short reverse sweep, keeping
// the update of its own control.  Same formulas in the same order as sgd_iteration, so the result is bit
// for bit the one of the other kernels; the dependent chain per iteration shrinks from H feature
// evaluations to one.
// ---------------------------------------------------------------------------------------------
static constexpr int kTG = 8;                 // lanes per (problem, start) for horizons up to 8
static constexpr int kTGMax = 16;             // ... and for horizons 9 .. 16
__host__ __device__ constexpr int tp_lanes(int HT) { return HT <= kTG ? kTG : kTGMax; }

template <int HT, int NOT_, int LT>
__device__ __forceinline__ float solve_start_tp(const KParams &k, const GradW &w, const float *wraw, int ws,
                                                float x0, float y0, float v0, float th0, const float *oth, int P,
                                                int t, float a_init, float w_init, Traj<HT> &u) {
    static_assert(HT > 0 && HT <= kTGMax, "time-parallel solve: one lane per horizon step");
    constexpr int NO = NOT_;
    constexpr int TG = tp_lanes(HT);
    const int tt = t < HT ? t : HT - 1;
    float ua = a_init, uw = w_init;                      // this lane's control: step tt
    float sn0, cs0;
    Mth<false>::sincos_(th0, sn0, cs0);
    const float c1 = -2.0f * k.mu * k.dt, c2 = -k.mu * k.dt2;
    const float lra = k.lr * k.hdt2, lrv = k.lr * k.dt;
    const float *omine = oth + (size_t)tt * NO * 2 * P;
#pragma unroll 1
    for (int it = 0; it < k.n_iter; ++it) {
        float A[HT], W[HT], sv[HT + 1], sc[HT + 1], ss[HT + 1], sd[HT];    // [j]: at the state before step j
        {   // the lanes exchange the CLIPPED controls (the rollout uses nothing else; the raw ones only mask the update)
            const float ac_own = fmaxf(fminf(ua, 4.0f), -8.0f);
            const float oc_own = fmaxf(fminf(uw, 4.0f), -4.0f);
#pragma unroll
            for (int j = 0; j < HT; ++j) {
                A[j] = __shfl_sync(OCD_FULL, ac_own, j, TG);
                W[j] = __shfl_sync(OCD_FULL, oc_own, j, TG);
            }
        }
        float v = v0, th = th0, sn = sn0, cs = cs0;
        // this lane's own state (after step tt) freezes once its step has passed: the position is accumulated under the
        // predicate j <= tt (same FMAs in the same order as the other kernels' rollout, no selects)
        float mx_ = x0, my_ = y0, mv_ = v0, msn = sn0, mcs = cs0;
#pragma unroll
        for (int j = 0; j < HT; ++j) {
            const float total = fmaf(-k.mu, v * v, A[j]);
            const float dist = fmaf(total, k.hdt2, v * k.dt);
            sv[j] = v; sc[j] = cs; ss[j] = sn; sd[j] = dist;
            const bool upto = j <= tt;
            if (upto) {
                mx_ = fmaf(cs, dist, mx_);
                my_ = fmaf(sn, dist, my_);
            }
            v = fmaf(total, k.dt, v);
            th = fmaf(W[j], k.dt, th);
            Mth<false>::sincos_(th, sn, cs);
            if (upto) {
                mv_ = v; msn = sn; mcs = cs;
            }
        }
        sv[HT] = v; sc[HT] = cs; ss[HT] = sn;
        float gx, hy, ke, unused;
        bool flag = false;
        feature_grad<NOT_, LT, false, false, 1>(k, w, mx_, my_, mv_, msn, mcs, omine, 2 * P, P, gx, hy, ke, unused,
                                                      0.0f, flag);
        if (__any_sync(OCD_FULL, flag))                  // rare: a lane needs an exact tie rule
            feature_grad<NOT_, LT, false, false, 0>(k, w, mx_, my_, mv_, msn, mcs, omine, 2 * P, P, gx, hy, ke,
                                                          unused, 0.0f, flag);
        float GX[HT], HY[HT], KE[HT];
#pragma unroll
        for (int j = 0; j < HT; ++j) {
            GX[j] = __shfl_sync(OCD_FULL, gx, j, TG);
            HY[j] = __shfl_sync(OCD_FULL, hy, j, TG);
            KE[j] = __shfl_sync(OCD_FULL, ke, j, TG);
        }
        float lx = 0.0f, ly = 0.0f, lv = 0.0f, lth = 0.0f;
        float ld_ = 0.0f, mv_t = 0.0f, mth_t = 0.0f;         // the three adjoint terms of this lane's own step
#pragma unroll
        for (int jj = 0; jj < HT; ++jj) {
            const int j = HT - 1 - jj;
            // the gradient at s_{j+1} plus the adjoint, fused exactly as the single-thread kernels fuse them.  A lane
            // walks the chain down to its own step and stops there (the steps below it run under a false predicate), so
            // what the three terms hold after the loop are the values of its own step -- no select chain
            // In C (what every other kernel's reverse sweep writes, statement for statement):
            //   if (j >= tt) { mx = GX[j] + lx;  my = fmaf(wcy, HY[j], ly);  mv = fmaf(KE[j], ss[j+1], lv);
            //                  mth = fmaf(KE[j] * sv[j+1], sc[j+1], lth);  ld = fmaf(sc[j], mx, ss[j] * my);
            //                  lv = fmaf(fmaf(c1, sv[j], 1), mv, fmaf(c2, sv[j], dt) * ld);
            //                  lth = fmaf(sd[j], fmaf(sc[j], my, -(ss[j] * mx)), mth);  lx = mx;  ly = my; }
            // written as predicated PTX because ptxas turns the C form into a divergent branch per step (BSSY / BRA /
            // BSYNC); explicit .rn on every mul / add keeps ptxas from contracting them, so the roundings are the C form's.
            asm("{\n\t"
                ".reg .pred p;\n\t"
                ".reg .f32 t0, t1, t2, t3, t4, t5;\n\t"        // temporaries are written unconditionally (no old value to keep)
                "setp.ne.s32 p, %21, 0;\n\t"
                "@p add.rn.f32 %0, %7, %0;\n\t"               // lx := mx = GX + lx
                "@p fma.rn.f32 %1, %17, %8, %1;\n\t"          // ly := my = wcy HY + ly
                "@p fma.rn.f32 %5, %9, %10, %2;\n\t"          // mv = KE ss[j+1] + lv
                "mul.rn.f32 t0, %9, %11;\n\t"                 // KE sv[j+1]
                "@p fma.rn.f32 %6, t0, %12, %3;\n\t"          // mth = (KE sv[j+1]) sc[j+1] + lth
                "mul.rn.f32 t1, %14, %1;\n\t"                 // ss[j] my
                "@p fma.rn.f32 %4, %13, %0, t1;\n\t"          // ld = sc[j] mx + ss[j] my
                "fma.rn.f32 t2, %18, %16, 0f3F800000;\n\t"    // c1 sv[j] + 1
                "fma.rn.f32 t3, %19, %16, %20;\n\t"           // c2 sv[j] + dt
                "mul.rn.f32 t3, t3, %4;\n\t"                  // (c2 sv[j] + dt) ld
                "@p fma.rn.f32 %2, t2, %5, t3;\n\t"           // lv
                "mul.rn.f32 t4, %14, %0;\n\t"                 // ss[j] mx
                "neg.f32 t4, t4;\n\t"
                "fma.rn.f32 t5, %13, %1, t4;\n\t"             // sc[j] my - ss[j] mx
                "@p fma.rn.f32 %3, %15, t5, %6;\n\t"          // lth = sd[j] (...) + mth
                "}"
                : "+f"(lx), "+f"(ly), "+f"(lv), "+f"(lth), "+f"(ld_), "+f"(mv_t), "+f"(mth_t)
                : "f"(GX[j]), "f"(HY[j]), "f"(KE[j]), "f"(ss[j + 1]), "f"(sv[j + 1]), "f"(sc[j + 1]), "f"(sc[j]),
                  "f"(ss[j]), "f"(sd[j]), "f"(sv[j]), "f"(w.wcy), "f"(c1), "f"(c2), "f"(k.dt), "r"((int)(j >= tt)));
        }
        // every lane walks the whole adjoint chain, but updates only its own control (same formula, same operands as
        // the single-thread kernels' update of step tt)
        const bool in_a = (ua >= -8.0f) && (ua <= 4.0f);
        const bool in_w = fabsf(uw) <= 4.0f;
        ua = in_a ? fmaf(lra, ld_, fmaf(lrv, mv_t, ua)) : ua;
        uw = in_w ? fmaf(lrv, mth_t, uw) : uw;
    }
#pragma unroll
    for (int j = 0; j < HT; ++j) {
        u.ua[j] = __shfl_sync(OCD_FULL, ua, j, TG);
        u.uw[j] = __shfl_sync(OCD_FULL, uw, j, TG);
    }
    // Final loss -R(u), R = sum_t w . phi(s_{t+1}) accumulated in step order (rollout_reward, naive_planner.py:43-77).
    // Every lane evaluates the reward of ITS step only: it reaches that state with rollout_reward's own dynamics
    // (dynamics_step's FMAs, the position accumulated under j <= tt), the lane values are all-gathered and added in
    // step order -- the same additions in the same order, but one feature evaluation on the dependent chain, not HT.
    {
        const float ac_own = fmaxf(fminf(ua, 4.0f), -8.0f);
        const float oc_own = fmaxf(fminf(uw, 4.0f), -4.0f);
        float v = v0, th = th0, sn = sn0, cs = cs0;
        float xq = x0, yq = y0, vq = v0, snq = sn0;
#pragma unroll
        for (int j = 0; j < HT; ++j) {
            const float aj = __shfl_sync(OCD_FULL, ac_own, j, TG);
            const float wj = __shfl_sync(OCD_FULL, oc_own, j, TG);
            const float total = fmaf(-k.mu, v * v, aj);
            const float dist = fmaf(total, k.hdt2, v * k.dt);
            const bool upto = j <= tt;
            if (upto) {
                xq = fmaf(cs, dist, xq);
                yq = fmaf(sn, dist, yq);
            }
            v = fmaf(total, k.dt, v);
            th = fmaf(wj, k.dt, th);
            Mth<false>::sincos_(th, sn, cs);
            if (upto) {
                vq = v; snq = sn;
            }
        }
        const float rv = reward_value<LT, false, true, false>(k, wraw, ws, xq, yq, vq, snq, omine, 2 * P, P);
        float r = 0.0f;
#pragma unroll
        for (int j = 0; j < HT; ++j) r = __fadd_rn(r, __shfl_sync(OCD_FULL, rv, j, TG));
        return -r;
    }
}

// ---------------------------------------------------------------------------------------------
// Runtime / long horizons: the segmented adjoint.
// The register file cannot hold 10 values per step for H = 15..64, and spilling them to local
// memory is what makes a naive runtime-H kernel slow.  Instead the horizon is cut into segments of
// SEG steps.  Pass 1 rolls the dynamics alone (no features) and checkpoints the state at every
// segment start; pass 2 walks the segments backwards, re-running each segment forward from its
// checkpoint WITH the feature gradients (register arrays of SEG entries) and then sweeping it in
// reverse.  The expensive part -- the feature gradient -- is still evaluated exactly once per step,
// so the cost over the register-resident kernel is one extra dynamics step (~15 %).  Controls
// [H][2] and checkpoints [nseg][4] live in shared memory, thread index fastest.
// ---------------------------------------------------------------------------------------------
// Forward half of pass 2 of one segment: from the segment's start state, with the feature gradients.
template <int SEG, int NOT_, int LT, bool PRECISE, bool LIN, bool FULL, int VM, bool SF = false, bool FR = false>
__device__ __forceinline__ void seg_forward(const KParams &k, const GradW &w, float x, float y, float v, float th,
                                            const float *os, int ostep, int P, const float *us, int rem, float tbase,
                                            float *ua, float *uw, float *sv, float *sc, float *ss, float *sd, float *gx,
                                            float *gy, float *gv, float *gth, bool &flag) {
    float sn, cs;
    Mth<PRECISE>::sincos_(th, sn, cs);
    const float *ot = os;
#pragma unroll
    for (int i = 0; i < SEG; ++i, ot += ostep) {
        if (SF) {                                  // step fence, see forward_sweep
            int one;
            asm volatile("mov.u32 %0, 1;" : "=r"(one));
            if (one == 0) continue;
        }
        if (FULL || i < rem) {
            const float2 uu = *reinterpret_cast<const float2 *>(us + 2 * i);
            ua[i] = uu.x;
            uw[i] = uu.y;
            const float ac = fmaxf(fminf(ua[i], 4.0f), -8.0f);
            const float oc = fmaxf(fminf(uw[i], 4.0f), -4.0f);
            const float total = fmaf(-k.mu, v * v, ac);
            const float dist = fmaf(total, k.hdt2, v * k.dt);
            sv[i] = v; sc[i] = cs; ss[i] = sn; sd[i] = dist;
            x = fmaf(cs, dist, x);
            y = fmaf(sn, dist, y);
            v = fmaf(total, k.dt, v);
            th = fmaf(oc, k.dt, th);
            Mth<PRECISE>::sincos_(th, sn, cs);
            feature_grad<NOT_, LT, PRECISE, LIN, VM, !PRECISE, FR>(k, w, x, y, v, sn, cs, ot, LIN ? 4 * P : 2 * P, P,
                                                                   gx[i], gy[i], gv[i], gth[i], tbase + (float)(i + 1),
                                                                   flag);
            sv[i + 1] = v; sc[i + 1] = cs; ss[i + 1] = sn;      // the state after step i (overwritten by step i+1's own save)
        }
    }
}

// Pass 2 of one segment: the forward half above, then the reverse sweep with the SGD update.  FULL: all SEG
// steps exist (no per-step predicates); otherwise the first `rem`.  LAT (small batches, FAST): the forward half
// runs straight-line (vote mode 1) and is repeated with the exact rules if a lane asked for one.
// Step fences in the segmented wide form: measured, they help with two or more other cars (H=15, 6 cars: 8.40 ->
// 7.65 ms) and cost with one (4.74 -> 4.84 ms), so they follow the car count.
template <int SEG, int NOT_, int LT, bool PRECISE, bool LIN, bool FULL, int LAT, bool FR = false>
__device__ __forceinline__ void seg_pass2(const KParams &k, const GradW &w, float x, float y, float v, float th,
                                          const float *os, int ostep, int P, float *us, int rem, float tbase,
                                          float (&lam)[4]) {
    float ua[SEG], uw[SEG], sv[SEG + 1], sc[SEG + 1], ss[SEG + 1], sd[SEG], gx[SEG], gy[SEG], gv[SEG], gth[SEG];
    bool flag = false;
    if (LAT != 0 && !PRECISE) {
        constexpr bool SF = LAT == 2 && NOT_ != 1;
        seg_forward<SEG, NOT_, LT, PRECISE, LIN, FULL, 1, SF, FR>(k, w, x, y, v, th, os, ostep, P, us, rem, tbase, ua, uw,
                                                                  sv, sc, ss, sd, gx, gy, gv, gth, flag);
        if (__any_sync(OCD_FULL, flag))
            seg_forward<SEG, NOT_, LT, PRECISE, LIN, FULL, 0, false, FR>(k, w, x, y, v, th, os, ostep, P, us, rem, tbase, ua,
                                                                         uw, sv, sc, ss, sd, gx, gy, gv, gth, flag);
    } else {
        seg_forward<SEG, NOT_, LT, PRECISE, LIN, FULL, 0, false, FR>(k, w, x, y, v, th, os, ostep, P, us, rem, tbase, ua, uw,
                                                                     sv, sc, ss, sd, gx, gy, gv, gth, flag);
    }
    float lx = lam[0], ly = lam[1], lv = lam[2], lth = lam[3];
    const float c1 = -2.0f * k.mu * k.dt, c2 = -k.mu * k.dt2;
    const float lra = k.lr * k.hdt2, lrv = k.lr * k.dt;
#pragma unroll
    for (int ii = 0; ii < SEG; ++ii) {
        const int i = SEG - 1 - ii;
        if (FULL || i < rem) {
            const float mx = gx[i] + lx;
            const float my = PRECISE ? gy[i] + ly : fmaf(w.wcy, gy[i], ly);
            const float mv = PRECISE ? gv[i] + lv : fmaf(gv[i], ss[i + 1], lv);
            const float mth = PRECISE ? gth[i] + lth : fmaf(__fmul_rn(gv[i], sv[i + 1]), sc[i + 1], lth);
            const float ld = fmaf(sc[i], mx, ss[i] * my);
            const float a = ua[i], om = uw[i];
            const bool in_a = (a >= -8.0f) && (a <= 4.0f);
            const bool in_w = fabsf(om) <= 4.0f;
            lv = fmaf(fmaf(c1, sv[i], 1.0f), mv, fmaf(c2, sv[i], k.dt) * ld);
            lth = fmaf(sd[i], fmaf(sc[i], my, -(ss[i] * mx)), mth);
            lx = mx;
            ly = my;
            if (PRECISE) {
                const float ga = in_a ? fmaf(k.hdt2, ld, k.dt * mv) : 0.0f;
                const float gw = in_w ? k.dt * mth : 0.0f;
                *reinterpret_cast<float2 *>(us + 2 * i) =         // u <- u - lr * d(-R)/du
                    make_float2(fmaf(k.lr, ga, a), fmaf(k.lr, gw, om));
            } else {                                           // the same with the constants folded
                *reinterpret_cast<float2 *>(us + 2 * i) =
                    make_float2(in_a ? fmaf(lra, ld, fmaf(lrv, mv, a)) : a, in_w ? fmaf(lrv, mth, om) : om);
            }
        }
    }
    lam[0] = lx; lam[1] = ly; lam[2] = lv; lam[3] = lth;
}

// HC > 0: the horizon is a compile-time constant (segment count and the length of the last segment fold away).
// FR: the register-resident kernels' folded lane term and single-reciprocal bump / fence (feature_grad's FOLD).
template <int SEG, int NOT_, int LT, bool PRECISE, bool LIN, int LAT = 0, int HC = 0, bool FR = false>
__device__ __forceinline__ void sgd_iteration_seg(const KParams &k, const GradW &w, float x0, float y0, float v0,
                                                  float th0, const float *oth, int P, const SmemTraj &u, float *ck) {
    const int H = HC > 0 ? HC : k.H;
    const int NO = NOT_ > 0 ? NOT_ : k.NO;
    const int nseg = (H + SEG - 1) / SEG;
    {   // pass 1: dynamics only; the state at the start of segments 1.. is checkpointed (segment 0 starts at
        // the initial state; all segments but the last are full)
        float x = x0, y = y0, v = v0, th = th0;
        const float *us = u.p;
        float *c = ck;
#pragma unroll 1
        for (int sg = 0; sg < nseg - 1; ++sg, us += 2 * SEG, c += 4) {
#pragma unroll
            for (int i = 0; i < SEG; ++i) {
                float sn, cs;
                Mth<PRECISE>::sincos_(th, sn, cs);
                const float2 uu = *reinterpret_cast<const float2 *>(us + 2 * i);
                const float ac = fmaxf(fminf(uu.x, 4.0f), -8.0f);
                const float oc = fmaxf(fminf(uu.y, 4.0f), -4.0f);
                const float total = fmaf(-k.mu, v * v, ac);
                const float dist = fmaf(total, k.hdt2, v * k.dt);
                x = fmaf(cs, dist, x);
                y = fmaf(sn, dist, y);
                v = fmaf(total, k.dt, v);
                th = fmaf(oc, k.dt, th);
            }
            *reinterpret_cast<float4 *>(c) = make_float4(x, y, v, th);
        }
    }
    float lam[4] = {0.0f, 0.0f, 0.0f, 0.0f};                   // adjoint of (x, y, v, th)
    const int ostep = LIN ? 0 : NO * 2 * P;                    // slab floats per horizon step
    float *us = u.p + 2 * SEG * (nseg - 1);
    const float *c = ck + 4 * (nseg - 2);                      // checkpoint of the last segment (sg - 1)
    const float *os = oth + (size_t)SEG * (nseg - 1) * ostep;
    float tbase = (float)(SEG * (nseg - 1));                   // steps before this segment
    const int rem = H - SEG * (nseg - 1);                      // steps in the last segment (it may be short)
    int sg = nseg - 1;
    if (rem != SEG) {      // a short last segment is handled ahead of the loop: the loop body carries no predicates
        float x = x0, y = y0, v = v0, th = th0;
        if (sg > 0) {
            const float4 st = *reinterpret_cast<const float4 *>(c);
            x = st.x; y = st.y; v = st.z; th = st.w;
        }
        seg_pass2<SEG, NOT_, LT, PRECISE, LIN, false, LAT, FR>(k, w, x, y, v, th, os, ostep, P, us, rem, tbase, lam);
        --sg; us -= 2 * SEG; c -= 4; os -= SEG * ostep; tbase -= (float)SEG;
    }
#pragma unroll 1
    for (; sg >= 0; --sg, us -= 2 * SEG, c -= 4, os -= SEG * ostep, tbase -= (float)SEG) {
        float x = x0, y = y0, v = v0, th = th0;
        if (sg > 0) {
            const float4 st = *reinterpret_cast<const float4 *>(c);
            x = st.x; y = st.y; v = st.z; th = st.w;
        }
        seg_pass2<SEG, NOT_, LT, PRECISE, LIN, true, LAT, FR>(k, w, x, y, v, th, os, ostep, P, us, SEG, tbase, lam);
    }
}

// The complete segmented solve for one (problem, start): start controls, n_iter iterations, final loss.
template <int SEG, int NOT_, int LT, bool PRECISE, bool LIN, int LAT = 0, int HC = 0, bool FR = false>
__device__ __forceinline__ float solve_start_seg(const KParams &k, const GradW &w, const float *wraw, int ws,
                                                 float x0, float y0, float v0, float th0, const float *oth, int P,
                                                 int s, float cur_speed, const SmemTraj &u, float *ck) {
    {
        const float a0 = (s >= 3) ? __fmul_rn(k.mu, __fmul_rn(cur_speed, cur_speed)) : 0.0f;
        const int m = s % 3;
        const float w0 = (m == 0) ? 0.0f : ((m == 1) ? -k.turn : k.turn);
        for (int t = 0; t < k.H; ++t) u.set(t, a0, w0);
    }
#pragma unroll 1
    for (int it = 0; it < k.n_iter; ++it)
        sgd_iteration_seg<SEG, NOT_, LT, PRECISE, LIN, LAT, HC, FR>(k, w, x0, y0, v0, th0, oth, P, u, ck);
    return -rollout_reward<0, LT, PRECISE, SmemTraj, LIN>(k, wraw, ws, x0, y0, v0, th0, oth, P, u);
}

}  // namespace ocd
