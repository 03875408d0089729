"""Latency of the time-parallel kernels at the reference's real sizes: one CMA-ES generation of finite_horizon
(45 episodes x 15 control steps) and local_opt (90), and a single solve of 45 problems.  Median of many launches.
    [OCD_B200_LIB=scratch/libocd_<variant>.so] python scripts/tuning/tp_lat.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import l4dc_mpc_ocd_b200 as ocd
import oracle as O
from l4dc_mpc_ocd_b200 import synthetic
eng = ocd.Engine(0)
def timed(fn, reps=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))
def ep(name, B, T=15):
    spec = O.scenario_params(name)
    op = spec.params
    p = ocd.PlannerParams(H=op.H, C=op.C, lane_x=tuple(op.lane_x), n_iter=op.n_iter, num_lanes=op.num_lanes, other_mode=op.other_mode, target_speed=op.target_speed, lr=op.lr)
    s = spec.scenario
    sc = ocd.Scenario(init_state=s.init_state, kind=s.kind, friction=s.friction, control=s.control, plan=s.plan, critical_t=s.critical_t, teleport_state=s.teleport_state)
    rng = np.random.default_rng(12)
    ri = np.tile(spec.example_init.astype(np.float32), (B, 1))
    ri[:, 0] += rng.uniform(-0.04, 0.04, B).astype(np.float32); ri[:, 1] += rng.uniform(-0.05, 0.05, B).astype(np.float32); ri[:, 2] += rng.uniform(-0.1, 0.1, B).astype(np.float32)
    wt = (spec.designer_weights / np.linalg.norm(spec.designer_weights)).astype(np.float32)
    cand = wt[None] + 0.05 * rng.normal(size=(9, p.K)).astype(np.float32); cand /= np.linalg.norm(cand, axis=1, keepdims=True)
    widx = (np.arange(B) % 9).astype(np.int32)
    ul = rng.integers(1, p.C, B).astype(np.int32) if name == "replanning" else None
    dev = eng.device
    t = lambda a, dt=torch.float32: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    robot, W, tw, wi, u = t(ri.T), t(cand.T), t(wt), t(widx, torch.int32), t(ul, torch.int32)
    out = eng.episodes_soa(p, sc, robot, W, 9, tw, T, weight_idx=wi, unlucky_idx=u)
    f = lambda: eng.episodes_soa(p, sc, robot, W, 9, tw, T, weight_idx=wi, unlucky_idx=u, out=out)
    med, mn = timed(f)
    return round(med, 4), round(mn, 4), float(out["returns"].double().sum())
def solve(B):
    p = ocd.PlannerParams(H=5, C=2, lr=0.1, n_iter=100)
    b = synthetic.make_batch(B, C=2, seed=99)
    world = torch.as_tensor(b["world"], device=eng.device).permute(1, 2, 0).contiguous()
    w = torch.as_tensor(b["weights"], device=eng.device).t().contiguous()
    idx = torch.as_tensor(b["weight_idx"], device=eng.device)
    out = eng.solve_soa(p, world, w, w.shape[1], idx)
    med, mn = timed(lambda: eng.solve_soa(p, world, w, w.shape[1], idx, out=out))
    return round(med, 4), round(mn, 4), float(out["losses"].double().sum())
print("lib", os.environ.get("OCD_B200_LIB", "stock"))
print("finite_horizon 45 episodes: median ms, min ms, checksum", ep("finite_horizon", 45), flush=True)
print("local_opt 90 episodes:", ep("local_opt", 90), flush=True)
print("replanning 90 episodes, T=20:", ep("replanning", 90, 20), flush=True)
for B in (45, 360):
    print("solve H=5, B=%d:" % B, solve(B), flush=True)
