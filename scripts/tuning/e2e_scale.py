#!/usr/bin/env python
"""End-to-end (host arrays in, host arrays out) scaling over the GPUs of one box: what limits N = 8?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/tuning/e2e_scale.py [--no-bind] [--problems B]

Per rank: B problems of the bench shape through ocd_solve_batch_host (pinned arrays: whole plans / first control only;
pageable arrays), every rank at once, max over ranks.  Rank 0 prints one JSON line per variant.  With --no-bind the
ranks do not partition the host's cores (round 1's behaviour)."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import l4dc_mpc_ocd_b200 as ocd                     # noqa: E402
from l4dc_mpc_ocd_b200 import parallel, synthetic   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problems", type=int, default=1 << 20)
    ap.add_argument("--no-bind", action="store_true")
    ap.add_argument("--reps", type=int, default=8)
    args = ap.parse_args()
    rank, lr, ws = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(lr)
    cpus = len(os.sched_getaffinity(0))
    if not args.no_bind:
        cpus = parallel.bind_rank_cpus(lr, int(os.environ.get("LOCAL_WORLD_SIZE", ws)))
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dev = torch.device("cuda", lr)
    B = args.problems
    p = ocd.PlannerParams()
    b = synthetic.make_batch(B, seed=1234 + rank)
    ctx = ocd.HostContext(lr)

    def pinned(a):
        buf = ocd.HostContext.pinned_empty(a.shape, a.dtype)
        buf[...] = a
        return buf

    world = pinned(np.ascontiguousarray(b["world"].transpose(1, 2, 0)))
    w = pinned(np.ascontiguousarray(b["weights"].T))
    idx = pinned(b["weight_idx"])
    full = dict(plan=ocd.HostContext.pinned_empty((p.H, 2, B)), losses=ocd.HostContext.pinned_empty((p.S, B)),
                best=ocd.HostContext.pinned_empty((B,), np.int32))
    first = dict(first=ocd.HostContext.pinned_empty((2, B)), losses=full["losses"], best=full["best"])
    only = dict(first=first["first"])
    pg = (np.array(world), np.array(w), np.array(idx))
    pg_out = {k: np.zeros_like(v) for k, v in full.items()}

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, reps):
        fn(); fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt = time.perf_counter() - t0
        if ws > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt / reps

    variants = {
        "pinned_full_plans": lambda: ctx.solve_soa(p, world, w, weight_idx=idx, out=full),
        "pinned_first_control_losses_best": lambda: ctx.solve_first_soa(p, world, w, weight_idx=idx, out=first),
        "pinned_first_control_only": lambda: ctx.solve_first_soa(p, world, w, weight_idx=idx, out=only, losses=False, best=False),
        "pageable_full_plans": lambda: ctx.solve_soa(p, pg[0], pg[1], weight_idx=pg[2], out=pg_out),
    }
    rg = (np.array(world), np.array(w), np.array(idx))
    rg_out = {k: np.zeros_like(v) for k, v in full.items()}
    for a in list(rg) + list(rg_out.values()):
        ocd.HostContext.register(a)
    variants["registered_full_plans"] = lambda: ctx.solve_soa(p, rg[0], rg[1], weight_idx=rg[2], out=rg_out)
    # device-resident reference point
    eng = ocd.Engine(lr)
    dw = torch.as_tensor(b["world"], device=dev).permute(1, 2, 0).contiguous()
    dwt = torch.as_tensor(b["weights"], device=dev).t().contiguous()
    di = torch.as_tensor(b["weight_idx"], device=dev)
    out = eng.solve_soa(p, dw, dwt, dwt.shape[1], di)

    def resident():
        eng.solve_soa(p, dw, dwt, dwt.shape[1], di, out=out)
        torch.cuda.synchronize(dev)
    variants = {"device_resident": resident, **variants}
    for name, fn in variants.items():
        ms = 1e3 * timed(fn, args.reps if "pageable" not in name else 3)
        if rank == 0:
            print(json.dumps(dict(variant=name, n_gpus=ws, bound_cpus_per_rank=cpus, bind=not args.no_bind, ms_per_step=round(ms, 3),
                                  solves_per_s=B * ws / (ms * 1e-3), host_threads_env=os.environ.get("OCD_HOST_THREADS"))), flush=True)
    ctx.close()
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
