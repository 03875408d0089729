#!/usr/bin/env python
"""BASELINE configs[4]: the synthetic sweep -- B in 4K..1M independent MPC problems, horizon 5/15/50,
2..6 cars -- on one GPU.  Prints one JSON line per point and writes them to the file given by --out.

    python scripts/sweep.py --out gpurun_out/sweep.json [--quick]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import l4dc_mpc_ocd_b200 as ocd                  # noqa: E402
from l4dc_mpc_ocd_b200 import synthetic         # noqa: E402

LR = {5: 0.1, 15: 0.02, 50: 0.0003}              # the reference's lr=0.1 is only stable at its own H=5/6 (bench.py: SWEEP_LR)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--horizons", type=int, nargs="*", default=[5, 15, 50])
    ap.add_argument("--cars", type=int, nargs="*", default=[2, 3, 4, 5, 6])
    ap.add_argument("--sizes", type=int, nargs="*", default=[4096, 16384, 65536, 262144, 1048576])
    args = ap.parse_args()
    eng = ocd.Engine(0)
    peak = eng.fp32_peak(8192)
    rows = []
    for H in args.horizons:
        for C in args.cars:
            for B in args.sizes:
                if args.quick and B > 65536:
                    continue
                p = ocd.PlannerParams(H=H, C=C, lr=LR.get(H, 0.1))
                b = synthetic.make_batch(B, C=C, seed=99)
                world = torch.as_tensor(b["world"], device=eng.device).permute(1, 2, 0).contiguous()
                w = torch.as_tensor(b["weights"], device=eng.device).t().contiguous()
                idx = torch.as_tensor(b["weight_idx"], device=eng.device)
                out = eng.solve_soa(p, world, w, w.shape[1], idx)
                torch.cuda.synchronize()
                reps = 3 if H * B >= 15 * 1048576 else 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    eng.solve_soa(p, world, w, w.shape[1], idx, out=out)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                fl = synthetic.flops_per_solve(H, C, 3)
                finite = bool(torch.isfinite(out["losses"]).all().item())
                nominal = 148 * 128 * 2 * 1.965e9
                row = dict(H=H, C=C, B=B, lr=p.lr, ms=ms, solves_per_s=B / (ms * 1e-3), form=ocd.kernel_form(p, B),
                           tflops=fl * B / (ms * 1e-3) / 1e12, frac_of_measured_fp32=fl * B / (ms * 1e-3) / peak,
                           frac_of_nominal_fp32=fl * B / (ms * 1e-3) / nominal, all_losses_finite=finite)
                rows.append(row)
                print(json.dumps(row), flush=True)
    if args.out:
        json.dump(dict(fp32_peak_tflops=peak / 1e12, rows=rows), open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
