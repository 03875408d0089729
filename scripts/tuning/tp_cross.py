"""Where the time-parallel form stops winning: whole finite_horizon / replanning episodes and H = 5 solves at growing batch
sizes, OCD_KERNEL_FORM=tp against latency (and auto).  Median of 20 launches.   python scripts/tuning/tp_cross.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import l4dc_mpc_ocd_b200 as ocd
import oracle as O
from l4dc_mpc_ocd_b200 import synthetic
eng = ocd.Engine(0)
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return round(float(np.median(ts)), 4)
def forms(fn):
    out = {}
    for f in ("tp", "latency", ""):
        if f: os.environ["OCD_KERNEL_FORM"] = f
        else: os.environ.pop("OCD_KERNEL_FORM", None)
        out[f or "auto"] = timed(fn)
    return out
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
    nc = max(1, B // 5)
    cand = wt[None] + 0.05 * rng.normal(size=(nc, p.K)).astype(np.float32); cand /= np.linalg.norm(cand, axis=1, keepdims=True)
    widx = (np.arange(B) % nc).astype(np.int32)
    ul = rng.integers(1, p.C, B).astype(np.int32) if name == "replanning" else None
    dev = eng.device
    t = lambda a, dt=torch.float32: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    robot, W, tw, wi, u = t(ri.T), t(cand.T), t(wt), t(widx, torch.int32), t(ul, torch.int32)
    out = eng.episodes_soa(p, sc, robot, W, nc, tw, T, weight_idx=wi, unlucky_idx=u)
    return forms(lambda: eng.episodes_soa(p, sc, robot, W, nc, tw, T, weight_idx=wi, unlucky_idx=u, out=out))
def solve(B, H=5):
    p = ocd.PlannerParams(H=H, C=2, lr={5: 0.1, 15: 0.02}[H], n_iter=100)
    b = synthetic.make_batch(B, C=2, seed=99)
    world = torch.as_tensor(b["world"], device=eng.device).permute(1, 2, 0).contiguous()
    w = torch.as_tensor(b["weights"], device=eng.device).t().contiguous()
    idx = torch.as_tensor(b["weight_idx"], device=eng.device)
    out = eng.solve_soa(p, world, w, w.shape[1], idx)
    return forms(lambda: eng.solve_soa(p, world, w, w.shape[1], idx, out=out))
for B in (270, 540, 1080, 1620, 2160, 2880, 4320, 5760):
    print("finite_horizon episodes B=%d (tp warps %d):" % (B, 3 * ((B + 3) // 4)), ep("finite_horizon", B), flush=True)
for B in (540, 1080, 2160, 2880):
    print("replanning episodes T=20 B=%d:" % B, ep("replanning", B, 20), flush=True)
for B in (540, 1080, 1620, 2160, 2880, 4320):
    print("solve H=5 B=%d:" % B, solve(B), flush=True)
for B in (720, 1440, 2160, 2880):
    print("solve H=15 B=%d:" % B, solve(B, 15), flush=True)
