#!/usr/bin/env python
"""A/B of the split form of k_solve (one-warp blocks + k_pick) against the three-warp-block wide form, same problems,
same staging rows (all_plans given in both arms).  Prints the times and whether the results are bit-identical.

    python scripts/tuning/ab_split.py [--B 1048576] [--reps 10]
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import l4dc_mpc_ocd_b200 as ocd                  # noqa: E402
from l4dc_mpc_ocd_b200 import synthetic         # noqa: E402

LR = {5: 0.1, 15: 0.02, 50: 0.0003}


def run(eng, p, world, w, idx, reps, split):
    os.environ["OCD_SPLIT"] = "1" if split else "0"
    out = eng.solve_soa(p, world, w, w.shape[1], idx, all_plans=True)
    for _ in range(2):
        eng.solve_soa(p, world, w, w.shape[1], idx, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.solve_soa(p, world, w, w.shape[1], idx, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--points", default="5:2,15:2,50:2,5:3,5:6,15:6,50:6")
    args = ap.parse_args()
    eng = ocd.Engine(0)
    for pt in args.points.split(","):
        H, C = (int(v) for v in pt.split(":"))
        p = ocd.PlannerParams(H=H, C=C, lr=LR.get(H, 0.1), n_iter=100)
        b = synthetic.make_batch(args.B, C=C, seed=99)
        world = torch.as_tensor(b["world"], device=eng.device).permute(1, 2, 0).contiguous()
        w = torch.as_tensor(b["weights"], device=eng.device).t().contiguous()
        idx = torch.as_tensor(b["weight_idx"], device=eng.device)
        reps = max(2, args.reps // (1 if H == 5 else (3 if H == 15 else 10)))
        ms0, o0 = run(eng, p, world, w, idx, reps, False)
        o0 = {k: v.clone() for k, v in o0.items()}
        ms1, o1 = run(eng, p, world, w, idx, reps, True)
        same = {k: bool(torch.equal(o0[k].view(torch.int32), o1[k].view(torch.int32))) for k in o0}
        fl = synthetic.flops_per_solve(H, C, 3)
        print(json.dumps(dict(H=H, C=C, B=args.B, ms_block=round(ms0, 4), ms_split=round(ms1, 4),
                              gain=round(ms0 / ms1, 4), tflops_split=round(fl * args.B / ms1 / 1e9, 2),
                              bit_identical=same)), flush=True)


if __name__ == "__main__":
    main()
