import os, sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import l4dc_mpc_ocd_b200 as ocd
from l4dc_mpc_ocd_b200 import synthetic
B = 1 << 20
p = ocd.PlannerParams()
hb = synthetic.make_batch(B, seed=1234)
ctx = ocd.HostContext(0)
def pinned(a):
    buf = ocd.HostContext.pinned_empty(a.shape, a.dtype); buf[...] = a; return buf
h_world = pinned(np.ascontiguousarray(hb["world"].transpose(1, 2, 0)))
h_w = pinned(np.ascontiguousarray(hb["weights"].T)); h_idx = pinned(hb["weight_idx"])
h_out = dict(plan=ocd.HostContext.pinned_empty((p.H, 2, B)), losses=ocd.HostContext.pinned_empty((p.S, B)),
             best=ocd.HostContext.pinned_empty((B,), np.int32))
scheds = ["", "1,1,1,1,1,1,1,1,1,1,1,1,1,1,1,1", "1,2,3,4,4,4,4,4,3,2,1", "1,2,4,8,8,4,2,1","1"]
for sc in scheds:
    if sc: os.environ["OCD_HOST_CHUNKS"] = sc
    else: os.environ.pop("OCD_HOST_CHUNKS", None)
    for _ in range(2): ctx.solve_soa(p, h_world, h_w, weight_idx=h_idx, out=h_out)
    t0 = time.perf_counter()
    for _ in range(8): ctx.solve_soa(p, h_world, h_w, weight_idx=h_idx, out=h_out)
    t = (time.perf_counter() - t0) / 8
    print(f"{sc or 'default':40s} {t*1e3:.3f} ms  {B/t:.4e} solves/s", flush=True)
