#!/usr/bin/env python
"""What the host sustains: every rank copies pinned host <-> device buffers in both directions at once (plain
cudaMemcpyAsync on two streams), all ranks together; rank 0 prints per-rank and aggregate GB/s.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/tuning/dma_probe.py"""
import json, os, time
import torch
import torch.distributed as dist

rank, lr, ws = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if ws > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
for mb_in, mb_out, label in ((44, 59, "the bench's bytes per step (44 MB in, 59 MB out)"), (64, 0, "H2D only"), (0, 64, "D2H only"), (64, 64, "64 MB each way")):
    hin = torch.empty(max(mb_in, 1) << 20, dtype=torch.uint8, pin_memory=True)
    hout = torch.empty(max(mb_out, 1) << 20, dtype=torch.uint8, pin_memory=True)
    din, dout = torch.empty_like(hin, device=dev), torch.empty_like(hout, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def step():
        if mb_in:
            with torch.cuda.stream(s1):
                din.copy_(hin, non_blocking=True)
        if mb_out:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        step()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    if ws > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        gbs = (mb_in + mb_out) * (1 << 20) * reps / dt / 1e9
        print(json.dumps(dict(case=label, n_gpus=ws, ms_per_step=round(1e3 * dt / reps, 3), gb_per_s_per_rank=round(gbs, 1),
                              gb_per_s_aggregate=round(gbs * ws, 1))), flush=True)
if ws > 1:
    dist.destroy_process_group()
