"""Host side of a lock-step generation alone (no GPU): optimize_cmaes_lockstep with a stand-in host context that returns
made-up episode returns at once.  What is left is the Python bookkeeping.   python scripts/tuning/lockstep_host.py [R] [gens] [--profile]"""
import cProfile, pstats, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import l4dc_mpc_ocd_b200.runtime as RT
from l4dc_mpc_ocd_b200.interact_drive.reward_design.mpc_ord import MPC_ORD, finite_horizon_env, optimize_cmaes_lockstep
class Fake:
    def episodes_soa(self, p, sc, robot_init, plan_weights, true_weights, T, weight_idx=None, unlucky_idx=None, final_world=False, **_):
        B = robot_init.shape[-1]
        w = np.asarray(plan_weights)[:, np.asarray(weight_idx)]
        ret = (-(w[0] - 0.3) ** 2 - 0.1 * np.asarray(robot_init)[0] ** 2).astype(np.float32)
        return ret, np.zeros((p.C, 4, B), np.float32)
RT.get_host_context = lambda device=None: Fake()
R = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 64
gens = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
def make():
    out = []
    for r in range(R):
        car, world, inits = finite_horizon_env(horizon=5, env_seeds=[1000 + 5 * r + i for i in range(5)], debug=False)
        out.append(MPC_ORD(world, car, inits, designer_horizon=15, verbose=False))
    return out
optimize_cmaes_lockstep(make(), list(range(1, R + 1)), sigma0=0.05, maxiter=3)
runs = make()
t0 = time.perf_counter()
optimize_cmaes_lockstep(runs, list(range(1, R + 1)), sigma0=0.05, maxiter=gens)
dt = time.perf_counter() - t0
print("R=%d: host side %.2f ms per generation" % (R, 1e3 * dt / (gens + 1)))
if "--profile" in sys.argv:
    runs = make()
    pr = cProfile.Profile(); pr.enable()
    optimize_cmaes_lockstep(runs, list(range(1, R + 1)), sigma0=0.05, maxiter=gens)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
