"""Run-sharded lock-step CMA-ES under torchrun: time of the local optimisation and of the end-of-run exchange.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/tuning/lockstep_shard.py [runs_per_gpu] [gens]"""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo')
rank, ws, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from l4dc_mpc_ocd_b200 import parallel as P
from l4dc_mpc_ocd_b200.interact_drive.reward_design.mpc_ord import MPC_ORD, finite_horizon_env, optimize_cmaes_lockstep
P.bind_rank_cpus(lr, ws)
rpg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
gens = int(sys.argv[2]) if len(sys.argv) > 2 else 15
R = rpg * ws
def make():
    out = []
    for r in range(R):
        car, world, inits = finite_horizon_env(horizon=5, env_seeds=[1000 + 5 * r + i for i in range(5)], debug=False)
        out.append(MPC_ORD(world, car, inits, designer_horizon=15, verbose=False))
    return out
seeds = list(range(1, R + 1))
optimize_cmaes_lockstep(make(), seeds, sigma0=0.05, shard_runs=True, maxiter=2)
runs = make()
mine = list(range(rank, R, ws))
torch.cuda.synchronize()
if ws > 1: dist.barrier()
t0 = time.perf_counter()
with P.local_only():
    xs = optimize_cmaes_lockstep([runs[i] for i in mine], [seeds[i] for i in mine], 0.05, maxiter=gens)
t1 = time.perf_counter()
packed = P.all_gather_objects([(i, x, runs[i]._export_state()) for i, x in zip(mine, xs)])
t2 = time.perf_counter()
for part in packed:
    for i, x, st in part:
        if i not in mine: runs[i]._import_state(st)
t3 = time.perf_counter()
print("rank %d: local optimisation %.2f ms per generation (%d runs), exchange %.1f ms, import %.1f ms" % (rank, 1e3 * (t1 - t0) / (gens + 1), len(mine), 1e3 * (t2 - t1), 1e3 * (t3 - t2)), flush=True)
if ws > 1: dist.destroy_process_group()
