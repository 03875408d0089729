import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import l4dc_mpc_ocd_b200 as ocd
import oracle as O
from l4dc_mpc_ocd_b200 import synthetic
eng = ocd.Engine(0)
def ep(name, B, form, T=15):
    if form: os.environ["OCD_KERNEL_FORM"] = form
    else: os.environ.pop("OCD_KERNEL_FORM", None)
    spec = O.scenario_params(name)
    op = spec.params
    p = ocd.PlannerParams(H=op.H, C=op.C, lane_x=tuple(op.lane_x), n_iter=op.n_iter, num_lanes=op.num_lanes, other_mode=op.other_mode, target_speed=op.target_speed, lr=op.lr)
    s = spec.scenario
    sc = ocd.Scenario(init_state=s.init_state, kind=s.kind, friction=s.friction, control=s.control, plan=s.plan, critical_t=s.critical_t, teleport_state=s.teleport_state)
    rng = np.random.default_rng(12)
    ri = np.tile(spec.example_init.astype(np.float32), (B, 1))
    ri[:, 0] += rng.uniform(-0.04, 0.04, B).astype(np.float32); ri[:, 1] += rng.uniform(-0.05, 0.05, B).astype(np.float32); ri[:, 2] += rng.uniform(-0.1, 0.1, B).astype(np.float32)
    wt = (spec.designer_weights / np.linalg.norm(spec.designer_weights)).astype(np.float32)
    cand = wt[None] + 0.05 * rng.normal(size=(B // 8, p.K)).astype(np.float32); cand /= np.linalg.norm(cand, axis=1, keepdims=True)
    widx = (np.arange(B) // 8).astype(np.int32)
    ul = rng.integers(1, p.C, B).astype(np.int32) if name == "replanning" else None
    eng.episodes(p, sc, ri, cand, wt, T, weight_idx=widx, unlucky_idx=ul)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.episodes(p, sc, ri, cand, wt, T, weight_idx=widx, unlucky_idx=ul); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for name in ("finite_horizon", "local_opt", "replanning"):
    for B in (45, 4096, 65536, 262144):
        print(name, B, {f or "auto": round(ep(name, B, f), 3) for f in ("", "throughput", "latency")}, flush=True)
