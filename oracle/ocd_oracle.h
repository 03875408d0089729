/*
 * ocd_oracle.h -- CPU restatement ("oracle") of the MPC hot path of avikj/L4DC-MPC-OCD.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker / the timed CPU baseline.  The product path
 * (l4dc-mpc-ocd_b200/, include/) never links, imports or calls anything in this directory.
 *
 * PARITY STATUS.  The reference is pure Python on TensorFlow 2.1; TensorFlow is not
 * installable here, so the reference cannot be executed as shipped.  The oracle is pinned
 * two ways (see DESIGN.md "Oracle"):
 *   1. against the reference's own known-answer tests (dynamics dt=1 cases, _f /
 *      smooth_threshold / smooth_bump doctests, the two planner KATs);
 *   2. against golden vectors in tests/golden/ produced by running the reference's
 *      UNMODIFIED Python sources (naive_planner.py, world.py, car/ *.py, merging.py,
 *      mpc_ord.py, the scenario constructors) on top of a torch-backed stand-in for the
 *      handful of TensorFlow primitives they call (oracle/tf_shim/, generator script
 *      tests/golden/make_golden.py).
 * What stays unpinned: the numerics of real TensorFlow kernels (1-ulp differences in
 * sin/cos/exp, reduce_sum order) and the third-party cma / scipy.truncnorm / TFP L-BFGS
 * behaviour -- "parity unpinned" for those.
 *
 * Every function exists in two precisions, suffix _f32 (mirrors the reference's float32
 * op order; compiled with -ffp-contract=off) and _f64 (same formulas in double).
 * All arrays are plain C row-major.  Car 0 is the planning ("robot") car.
 */
#ifndef OCD_ORACLE_H
#define OCD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCDO_MAX_LANES 4
#define OCDO_MAX_CARS  8
#define OCDO_MAX_PLAN  16
#define OCDO_MAX_H     64
#define OCDO_MAX_S     6
#define OCDO_LBFGS_M   4     /* correction pairs                                        */
#define OCDO_LBFGS_LS  6     /* backtracking trials                                     */
#define OCDO_LBFGS_MAX_H 16

/* Planner / world constants.  Doubles here; the _f32 functions cast each one to float at
 * the point of use, which is what TF does with Python-float constants. */
typedef struct {
    int32_t H;           /* planning horizon                       naive_planner.py:24     */
    int32_t C;           /* number of cars, robot = car 0                                  */
    int32_t L;           /* number of lanes (K = L + 4 features)   merging.py:61-65        */
    int32_t n_iter;      /* SGD steps per start                    naive_planner.py:151    */
    int32_t num_lanes;   /* fence threshold = 0.05*num_lanes       merging.py:80           */
    int32_t other_mode;  /* 0 constant velocity, 1 known controls  naive_planner.py:53-66  */
    int32_t extra_inits; /* 3 extra starts with a0 = mu*v^2        naive_planner.py:112-116*/
    int32_t optimizer;   /* 0 SGD (the reference); 1 the engine's opt-in L-BFGS (no reference counterpart that runs) */
    double  lr;          /* SGD learning rate                      naive_planner.py:28     */
    double  dt;          /* world.dt                               world.py:18             */
    double  friction;    /* robot friction                         car.py:33               */
    double  target_speed;/*                                        merging.py:29           */
    double  lane_x[OCDO_MAX_LANES]; /* x of each lane median       world.py:149-159        */
} ocdo_params;

/* Scripted cars + replanning world, for the episode loop. */
typedef struct {
    int32_t n_other;                          /* C - 1                                                  */
    int32_t critical_t;                       /* 0 = plain CarWorld; else replanning_world.py:29-36     */
    int32_t kind[OCDO_MAX_CARS];              /* 0 FixedControl/FixedVelocity, 1 FixedPlan              */
    int32_t plan_len[OCDO_MAX_CARS];
    double  init_state[OCDO_MAX_CARS][4];
    double  friction[OCDO_MAX_CARS];          /* 0 for FixedVelocityCar, 0.2 default for FixedPlanCar   */
    double  control[OCDO_MAX_CARS][2];        /* fixed control / default_control                        */
    double  plan[OCDO_MAX_CARS][OCDO_MAX_PLAN][2];
    double  teleport_state[4];                /* [10,0,0,0]                                             */
} ocdo_scenario;

#define OCDO_DECL(SUF, REAL)                                                                          \
    void ocdo_dynamics_step_##SUF(const REAL s[4], const REAL u[2], double dt, double mu,            \
                                  REAL out[4]);                                                       \
    REAL ocdo_f_##SUF(REAL x, REAL shape);                                                            \
    REAL ocdo_smooth_threshold_##SUF(REAL z, double threshold, double width, double c);              \
    REAL ocdo_smooth_bump_##SUF(REAL z, REAL start, REAL end);                                        \
    void ocdo_features_##SUF(const ocdo_params *p, const REAL *world /*[C][4]*/,                     \
                             REAL *phi /*[K]*/, REAL *jac /*[K][4] d phi / d robot state, or NULL*/);\
    int  ocdo_mpc_reward_##SUF(const ocdo_params *p, const REAL *init_world /*[C][4]*/,              \
                               const REAL *controls /*[H][2]*/,                                       \
                               const REAL *other_controls /*[C][H][2] (row 0 unused) or NULL*/,      \
                               const REAL *w /*[K]*/, REAL *R, REAL *grad /*[H][2] or NULL*/);        \
    int  ocdo_generate_plan_##SUF(const ocdo_params *p, const REAL *init_world,                      \
                                  const REAL *other_controls, const REAL *w, REAL cur_speed,         \
                                  REAL *plan /*[H][2]*/, REAL *losses /*[S]*/, int32_t *best,        \
                                  REAL *all_plans /*[S][H][2] or NULL*/);                            \
    int  ocdo_generate_plan_batch_##SUF(const ocdo_params *p, int64_t B,                             \
                                  const REAL *init_world /*[B][C][4]*/,                              \
                                  const REAL *other_controls /*[B][C][H][2] or NULL*/,               \
                                  const REAL *w /*[B][K]*/, REAL *plan /*[B][H][2]*/,                \
                                  REAL *losses /*[B][S]*/, int32_t *best /*[B]*/, int nthreads);     \
    int  ocdo_episode_##SUF(const ocdo_params *p, const ocdo_scenario *sc,                           \
                            const REAL robot_init[4], const REAL *w_plan /*[K]*/,                     \
                            const REAL *w_true /*[K]*/, int32_t unlucky_idx, int32_t T,              \
                            REAL *ret, REAL *traj_controls /*[T][2] or NULL*/,                        \
                            int32_t *traj_best /*[T] or NULL*/,                                       \
                            REAL *traj_states /*[T][C][4] past states, or NULL*/,                     \
                            REAL *step_rewards /*[T] or NULL*/);                                      \
    int  ocdo_episode_batch_##SUF(const ocdo_params *p, const ocdo_scenario *sc, int64_t B,          \
                            const REAL *robot_init /*[B][4]*/, const REAL *w_plan /*[B][K]*/,        \
                            const REAL *w_true /*[K]*/, const int32_t *unlucky_idx /*[B] or NULL*/,  \
                            int32_t T, REAL *ret /*[B]*/, int nthreads);

OCDO_DECL(f32, float)
OCDO_DECL(f64, double)

int ocdo_num_starts(const ocdo_params *p);
int ocdo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
