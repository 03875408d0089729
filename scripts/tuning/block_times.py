"""Per-block start / end times of one Q-kernel launch (variant library built with -DOCD_BLOCK_TIMES): who straggles?"""
import ctypes, os, sys, json
import numpy as np, torch
sys.path.insert(0, '/root/repo')
os.environ["OCD_B200_LIB"] = "scratch/libocd_btimes.so"
os.environ["OCD_KERNEL_FORM"] = "wide"
import l4dc_mpc_ocd_b200 as ocd
from l4dc_mpc_ocd_b200 import synthetic, _native as N
eng = ocd.Engine(0)
for B in (23680, 47360, 65536):
    p = ocd.PlannerParams(H=15, C=2, lr=0.02)
    b = synthetic.make_batch(B, C=2, seed=99)
    world = torch.as_tensor(b["world"], device=eng.device).permute(1, 2, 0).contiguous()
    w = torch.as_tensor(b["weights"], device=eng.device).t().contiguous()
    idx = torch.as_tensor(b["weight_idx"], device=eng.device)
    out = eng.solve_soa(p, world, w, w.shape[1], idx)
    torch.cuda.synchronize()
    eng.solve_soa(p, world, w, w.shape[1], idx, out=out)
    torch.cuda.synchronize()
    buf = np.zeros(3 * 4096, np.uint64)
    print("B", B)
    get = N.lib.ocd_debug_block_times
    get(buf.ctypes.data_as(ctypes.c_void_p))
    nb = (B + 31) // 32
    t = buf.reshape(-1, 3)[:min(nb, 4096)].astype(np.float64)
    t0 = t[:, 0].min()
    dur = (t[:, 1] - t[:, 0]) / 1e3
    print(json.dumps(dict(blocks=int(len(t)), total_us=float((t[:, 1].max() - t0) / 1e3), start_spread_us=float((t[:, 0].max() - t0) / 1e3),
                          dur_us_min=float(dur.min()), dur_us_med=float(np.median(dur)), dur_us_max=float(dur.max()),
                          sms=int(len(np.unique(t[:, 2]))), blocks_per_sm_max=int(np.bincount(t[:, 2].astype(int)).max()))))
    if B == 23680:
        sm = t[:, 2].astype(int)
        per_sm = np.array([dur[sm == i].mean() for i in range(148)])
        spread_in_sm = np.array([dur[sm == i].max() - dur[sm == i].min() for i in range(148)])
        print(" per-SM mean duration (us), SM 0..147:", np.round(per_sm, 0).astype(int).tolist())
        print(" spread inside an SM (us): median %.1f max %.1f" % (np.median(spread_in_sm), spread_in_sm.max()))
        print(" block index -> SM of the first 20 blocks:", sm[:20].tolist())
    order = np.argsort(t[:, 0])
    print(" first-wave durations (us) by start order, every 74th block:", np.round(dur[order][::74][:12], 1).tolist())
