#!/usr/bin/env python
"""Small batches at H = 15: the 16-lane time-parallel form against the latency form (OCD_KERNEL_FORM), ms per launch."""
import json, os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import l4dc_mpc_ocd_b200 as ocd
from l4dc_mpc_ocd_b200 import synthetic

eng = ocd.Engine(0)
for H, C in ((15, 2), (15, 3), (5, 2)):
    p = ocd.PlannerParams(H=H, C=C, lr=0.02 if H == 15 else 0.1)
    for B in (45, 90, 180, 360, 720, 1440):
        b = synthetic.make_batch(B, C=C, seed=5)
        world = torch.as_tensor(b["world"], device=eng.device).permute(1, 2, 0).contiguous()
        w = torch.as_tensor(b["weights"], device=eng.device).t().contiguous()
        idx = torch.as_tensor(b["weight_idx"], device=eng.device)
        row = dict(H=H, C=C, B=B)
        for form in ("tp", "latency", "auto"):
            if form == "auto":
                os.environ.pop("OCD_KERNEL_FORM", None)
            else:
                os.environ["OCD_KERNEL_FORM"] = form
            out = eng.solve_soa(p, world, w, w.shape[1], idx)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                eng.solve_soa(p, world, w, w.shape[1], idx, out=out)
            e1.record()
            torch.cuda.synchronize()
            row[form + "_ms"] = round(e0.elapsed_time(e1) / 20, 4)
        row["auto_form"] = ocd.kernel_form(p, B)
        print(json.dumps(row), flush=True)
