import os, sys, json, torch
sys.path.insert(0, '/root/repo')
import l4dc_mpc_ocd_b200 as ocd
from l4dc_mpc_ocd_b200 import synthetic
eng = ocd.Engine(0)
def run(H, C, B, form, lr):
    if form: os.environ["OCD_KERNEL_FORM"] = form
    else: os.environ.pop("OCD_KERNEL_FORM", None)
    p = ocd.PlannerParams(H=H, C=C, lr=lr)
    b = synthetic.make_batch(B, C=C, seed=99)
    world = torch.as_tensor(b["world"], device=eng.device).permute(1, 2, 0).contiguous()
    w = torch.as_tensor(b["weights"], device=eng.device).t().contiguous()
    idx = torch.as_tensor(b["weight_idx"], device=eng.device)
    out = eng.solve_soa(p, world, w, w.shape[1], idx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.solve_soa(p, world, w, w.shape[1], idx, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for H, C, lr in ((5, 2, 0.1), (15, 2, 0.03), (50, 2, 0.003)):
    for B in (512, 1024, 2048, 4096, 8192, 12288, 16384, 24576, 32768, 65536):
        row = {f or "auto": round(run(H, C, B, f, lr), 4) for f in ("", "throughput", "latency") + (("tp",) if H <= 8 else ())}
        print(H, C, B, row, flush=True)
