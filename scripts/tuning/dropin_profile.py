"""Where the host time of MPC_ORD.eval_weights_batch goes (cProfile over 2000 calls, finite_horizon 9 x 5)."""
import cProfile, pstats, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from l4dc_mpc_ocd_b200.interact_drive.reward_design.mpc_ord import MPC_ORD, finite_horizon_env
car, world, inits = finite_horizon_env(horizon=5, env_seeds=[1, 2, 3, 4, 5], debug=False)
ord_ = MPC_ORD(world, car, inits, designer_horizon=15, verbose=False)
rng = np.random.default_rng(0)
W = np.asarray(car.weights)[None] + 0.05 * rng.normal(size=(9, 7))
for _ in range(10): ord_.eval_weights_batch(W)
n = 2000
t0 = time.perf_counter()
for _ in range(n): ord_.eval_weights_batch(W)
print("plain: %.1f us per call" % (1e6 * (time.perf_counter() - t0) / n))
pr = cProfile.Profile(); pr.enable()
for _ in range(n): ord_.eval_weights_batch(W)
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
