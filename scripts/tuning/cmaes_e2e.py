import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
import l4dc_mpc_ocd_b200 as ocd
from l4dc_mpc_ocd_b200.interact_drive.reward_design.mpc_ord import MPC_ORD, finite_horizon_env
car, world, inits = finite_horizon_env(horizon=5, env_seeds=[1, 2, 3, 4, 5], debug=False)
ord_ = MPC_ORD(world, car, inits, designer_horizon=15, verbose=False)
rng = np.random.default_rng(0)
W = np.asarray(car.weights)[None] + 0.05 * rng.normal(size=(9, 7))
for _ in range(3): ord_.eval_weights_batch(W)
torch.cuda.synchronize()
t0 = time.perf_counter(); n = 50
for _ in range(n): r = ord_.eval_weights_batch(W)
dt = (time.perf_counter() - t0) / n
print("eval_weights_batch (9 candidates x 5 inits x 15 steps): %.3f ms per generation -> %.0f candidate-evals/s" % (dt * 1e3, 9 / dt))
t0 = time.perf_counter()
res = ord_.optimize_cmaes(seed=1, sigma0=0.05, maxiter=30)
dt = time.perf_counter() - t0
print("optimize_cmaes 30 generations: %.1f ms total, %.3f ms per generation" % (dt * 1e3, dt * 1e3 / 30))
