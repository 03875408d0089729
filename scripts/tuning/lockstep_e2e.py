"""Wall clock of optimize_cmaes_lockstep (R independent finite_horizon CMA-ES runs, n_inits 5 each) per generation,
against the launch alone: how much of a lock-step generation is host bookkeeping.   python scripts/tuning/lockstep_e2e.py [R] [gens]"""
import cProfile, pstats, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from l4dc_mpc_ocd_b200.interact_drive.reward_design.mpc_ord import MPC_ORD, finite_horizon_env, optimize_cmaes_lockstep
R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 20
def make(R):
    runs = []
    for r in range(R):
        car, world, inits = finite_horizon_env(horizon=5, env_seeds=[1000 + 5 * r + i for i in range(5)], debug=False)
        runs.append(MPC_ORD(world, car, inits, designer_horizon=15, verbose=False))
    return runs
optimize_cmaes_lockstep(make(R), list(range(1, R + 1)), sigma0=0.05, maxiter=3)      # warm-up at the same batch size (graph capture, buffers)
runs = make(R)
torch.cuda.synchronize()
t0 = time.perf_counter()
optimize_cmaes_lockstep(runs, list(range(1, R + 1)), sigma0=0.05, maxiter=gens)
dt = time.perf_counter() - t0
print("R=%d runs, %d generations: %.2f ms per generation, %.0f candidate-evals/s end to end" % (R, gens, 1e3 * dt / (gens + 1), R * 9 * gens / dt))
if "--profile" in sys.argv:
    runs = make(R)
    pr = cProfile.Profile(); pr.enable()
    optimize_cmaes_lockstep(runs, list(range(1, R + 1)), sigma0=0.05, maxiter=gens)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumtime").print_stats(25)
