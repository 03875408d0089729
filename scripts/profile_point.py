#!/usr/bin/env python
"""One point of the synthetic sweep, a few launches -- the command ncu wraps for the sweep kernels.

    python scripts/profile_point.py --H 15 --C 2 --B 262144 [--reps 3]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import l4dc_mpc_ocd_b200 as ocd                  # noqa: E402
from l4dc_mpc_ocd_b200 import synthetic         # noqa: E402

LR = {5: 0.1, 15: 0.02, 50: 0.0003}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--H", type=int, default=15)
    ap.add_argument("--C", type=int, default=2)
    ap.add_argument("--B", type=int, default=262144)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--n_iter", type=int, default=100)
    args = ap.parse_args()
    eng = ocd.Engine(0)
    p = ocd.PlannerParams(H=args.H, C=args.C, lr=LR.get(args.H, 0.1), n_iter=args.n_iter)
    b = synthetic.make_batch(args.B, C=args.C, seed=99)
    world = torch.as_tensor(b["world"], device=eng.device).permute(1, 2, 0).contiguous()
    w = torch.as_tensor(b["weights"], device=eng.device).t().contiguous()
    idx = torch.as_tensor(b["weight_idx"], device=eng.device)
    out = eng.solve_soa(p, world, w, w.shape[1], idx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        eng.solve_soa(p, world, w, w.shape[1], idx, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    fl = synthetic.flops_per_solve(args.H, args.C, 3) * (args.n_iter / 100.0)
    print(json.dumps(dict(H=args.H, C=args.C, B=args.B, ms=ms, solves_per_s=args.B / (ms * 1e-3),
                          tflops=fl * args.B / (ms * 1e-3) / 1e12)))


if __name__ == "__main__":
    main()
