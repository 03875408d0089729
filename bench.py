#!/usr/bin/env python
"""bench.py -- batched MPC solve throughput of the B200 engine (and the CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--problems B]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one batch: ONE ocd_solve_batch launch that runs
NaivePlanner.generate_plan (3 starts x 100 SGD iterations + final losses + argmin) for B
independent problems of the finite_horizon shape (H=5, 2 cars, 3 lanes).  Rank r solves its own
B problems (weak scaling, no data-path collective: the problems are independent); the value is
all ranks' solves divided by the slowest rank's device time.

Extra keys on the JSON line: roofline (FP32-pipe bound; the peak is measured in this job with a
dependent-free FMA kernel, since MEASURED_PEAKS.json has no FP32 figure), hbm (algorithmic bytes
vs the measured copy bandwidth, to show the path is nowhere near memory-bound), e2e (same metric
through the host-buffer C ABI, copies inside), cpu_baseline (the CPU oracle on this box's cores),
cmaes (candidate-evals/s of the finite_horizon cmaes --n_inits 5 generation, one episode launch),
horizons (solves/s at H = 15 and H = 50, the other horizons BASELINE's metric names).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SHAPE = dict(H=5, C=2, L=3, S=3, n_iter=100)
METRIC, UNIT = "mpc_solves_per_sec", "solves/s"


def _config(B, n_gpus):
    return {"workload": "synthetic sweep (BASELINE configs[4]) at the finite_horizon shape: "
                        f"{B} independent MPC problems per GPU, H=5, 2 cars, 3 lanes, 3 starts x 100 SGD iterations",
            "problems_per_gpu": B, "n_gpus": n_gpus, "horizon": 5, "cars": 2, "lanes": 3, "starts": 3,
            "n_iter": 100, "lr": 0.1, "inits_per_candidate": 5, "parallelism": f"problem-sharded x{n_gpus}",
            "l2_policy": "4 rotating input/output sets (>= 350 MB in total, L2 is 126 MB)"}


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def _visible_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Host cores this process may use.  (torchrun exports OMP_NUM_THREADS=1; the oracle's OpenMP loop
    takes its thread count as an argument, so that default does not throttle the CPU arm.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_arm(problems: int, threads: int | None = None, seed: int = 1234):
    """Times the CPU oracle (oracle/: C restatement of the reference's planner, OpenMP over
    problems) on `problems` synthetic problems of the bench shape.  -> (solves/s, seconds, threads)"""
    import oracle as O
    from l4dc_mpc_ocd_b200 import synthetic
    threads = threads or host_threads()
    batch = synthetic.make_batch(problems, seed=seed)
    w = batch["weights"][batch["weight_idx"]]
    p = O.OracleParams()
    t0 = time.perf_counter()
    O.generate_plan_batch(p, batch["world"], w, nthreads=threads)
    dt = time.perf_counter() - t0
    return problems / dt, dt, threads


def reference_main(args, rank: int, world_size: int):
    """--impl reference: the reference planner's CPU implementation on this box's host cores.
    TensorFlow (the reference's substrate) is not installable here, so the timed code is the
    oracle port of NaivePlanner.generate_plan, all host threads (cpu_baseline.kind = "port")."""
    if rank != 0:
        return
    threads = host_threads()
    rate, _, _ = cpu_arm(max(256, 32 * threads), threads)          # calibrate
    per_step = int(max(256, min(rate * 1.5, 2_000_000)))           # about 1.5 s of CPU work per step
    for _ in range(args.warmup):
        cpu_arm(max(256, per_step // 8), threads)
    t_tot = 0.0
    for i in range(args.steps):
        _, dt, _ = cpu_arm(per_step, threads, seed=1234 + i)
        t_tot += dt
    value = per_step * args.steps / t_tot
    cfg = _config(args.problems, args.gpus)
    sample = f"{per_step} synthetic problems of the bench shape per step x {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--problems", type=int, default=1 << 20, help="MPC problems per GPU per step")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / cpu_baseline / cmaes legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_main(args, rank, world_size)
        return

    import torch
    import torch.distributed as dist
    import l4dc_mpc_ocd_b200 as ocd
    from l4dc_mpc_ocd_b200 import synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world_size > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = ocd.Engine(local_rank)
    dev = eng.device
    B, K, W = args.problems, args.steps, args.warmup
    p = ocd.PlannerParams()                  # finite_horizon shape, fast-math kernels

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # inputs resident in HBM, SoA; 4 rotating sets so that no step finds its data in L2
    NSETS = 4
    sets = []
    for i in range(NSETS):
        b = synthetic.make_batch(B, seed=1234 + 17 * rank + i)
        world = torch.as_tensor(b["world"], device=dev).permute(1, 2, 0).contiguous()
        weights = torch.as_tensor(b["weights"], device=dev).t().contiguous()
        idx = torch.as_tensor(b["weight_idx"], device=dev)
        out = dict(plan=torch.empty((p.H, 2, B), dtype=torch.float32, device=dev),
                   losses=torch.empty((p.S, B), dtype=torch.float32, device=dev),
                   best=torch.empty((B,), dtype=torch.int32, device=dev))
        sets.append((world, weights, idx, out, b))
    Bw = sets[0][1].shape[1]

    def step(i):
        world, weights, idx, out, _ = sets[i % NSETS]
        eng.solve_soa(p, world, weights, Bw, idx, out=out)

    fp32_peak = eng.fp32_peak(8192)          # roofline denominator, measured before the run
    for i in range(W):
        step(i)
    barrier()
    sampler = ClockSampler(_visible_index(local_rank))
    sampler.start()
    launches0 = eng.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(W + i)
    e1.record()
    barrier()
    ms_total = max_over_ranks(float(e0.elapsed_time(e1)))
    clocks = sampler.stop()
    launches = eng.kernel_launches - launches0
    value = B * K * world_size / (ms_total * 1e-3)
    ms_per_step = ms_total / K

    # roofline of the dominant (only) kernel: k_solve
    flops = synthetic.flops_per_solve(SHAPE["H"], SHAPE["C"], SHAPE["L"], SHAPE["S"], SHAPE["n_iter"])
    hbm_bytes = synthetic.hbm_bytes_per_solve(SHAPE["H"], SHAPE["C"], SHAPE["L"], SHAPE["S"])
    own_ms = float(e0.elapsed_time(e1)) / K
    achieved_tf = flops * B / (own_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
    except Exception:
        pass
    traffic = None
    try:
        traffic = json.load(open(ROOT / "profiles" / "traffic.json")).get("k_solve_bytes_per_launch")
    except Exception:
        pass
    nominal_tf = 148 * 128 * 2 * (clocks.get("sm_max_mhz") or 1965) * 1e6 / 1e12
    roofline = {"bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
                "frac": achieved_tf / (fp32_peak / 1e12), "traffic": traffic,
                "bound_note": "BASELINE's north_star names the FP32 pipe as this path's roofline: no stage is a dense "
                              "contraction (no tensor roofline) and HBM carries 0.3 % of its measured bandwidth (see hbm)",
                "peak_source": "dependent-free FFMA kernel measured in this job (MEASURED_PEAKS.json has no FP32 figure)",
                "nominal_peak": nominal_tf, "frac_of_nominal": achieved_tf / nominal_tf,
                "flops_per_solve": flops, "kernel": "k_solve<5,1,3,fast,wide>", "kernel_ms": own_ms}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm = {"achieved": hbm_bytes * B / (own_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
           "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback",
           "bytes_per_solve": hbm_bytes}
    hbm["frac"] = hbm["achieved"] / hbm_peak

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": _config(B, world_size), "clocks": clocks,
            "gpu_launches": launches, "roofline": roofline, "hbm": hbm}

    # ---- e2e: the same solves through the host-buffer C ABI (ocd_solve_batch_host) ----------------
    # inputs start in pinned host memory, results end in pinned host memory; every step pays the
    # host->device copy of its inputs and the device->host copy of plans, losses and winners.
    if not args.no_extras:
        ctx = ocd.HostContext(local_rank)
        hb = sets[0][4]

        def pinned(a):
            buf = ocd.HostContext.pinned_empty(a.shape, a.dtype)
            buf[...] = a
            return buf

        h_world = pinned(np.ascontiguousarray(hb["world"].transpose(1, 2, 0)))
        h_w = pinned(np.ascontiguousarray(hb["weights"].T))
        h_idx = pinned(hb["weight_idx"])
        h_out = dict(plan=ocd.HostContext.pinned_empty((p.H, 2, B)), losses=ocd.HostContext.pinned_empty((p.S, B)),
                     best=ocd.HostContext.pinned_empty((B,), np.int32))
        ke = max(3, min(K, 10))
        for _ in range(2):
            ctx.solve_soa(p, h_world, h_w, weight_idx=h_idx, out=h_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            r = ctx.solve_soa(p, h_world, h_w, weight_idx=h_idx, out=h_out)
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        # the same call on ordinary pageable numpy arrays (staged through the context's pinned area)
        # (outputs too: ordinary numpy arrays reused across steps, as a steady-state caller would)
        pg_world, pg_w, pg_idx = np.array(h_world), np.array(h_w), np.array(h_idx)
        pg_out = {k: np.zeros_like(v) for k, v in h_out.items()}
        ctx.solve_soa(p, pg_world, pg_w, weight_idx=pg_idx, out=pg_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.solve_soa(p, pg_world, pg_w, weight_idx=pg_idx, out=pg_out)
        t_pageable = max_over_ranks(time.perf_counter() - t0) / 3
        line["e2e"] = {"value": B * ke * world_size / t_e2e, "unit": UNIT,
                       "h2d_bytes_per_step": int(h_world.nbytes + h_w.nbytes + h_idx.nbytes),
                       "d2h_bytes_per_step": int(r["plan"].nbytes + r["losses"].nbytes + r["best"].nbytes),
                       "steps": ke, "ms_per_step": 1e3 * t_e2e / ke,
                       "api": "ocd_solve_batch_host: pinned host arrays in, pinned host arrays out, "
                              "ramped column chunks over an H2D stream, two compute lanes and a D2H stream, so copies overlap the solve",
                       "pageable_host_arrays_value": B * world_size / t_pageable}
        line["gpu_launches"] = launches
        ctx.close()

        # ---- candidate-evals/s: one CMA-ES generation of finite_horizon cmaes --n_inits 5 -------------
        line["cmaes"] = cmaes_leg(eng, ocd, dist if distributed else None, rank, world_size, max_over_ranks, barrier)

        # ---- the other horizons of the metric (H = 5..50): two points of BASELINE configs[4] -------------------
        line["horizons"] = horizons_leg(eng, ocd, synthetic, rank, world_size, max_over_ranks, barrier, fp32_peak)

    # ---- CPU baseline beside it (rank 0, N=1 only) -------------------------------------------------
    if not args.no_extras and world_size == 1:
        threads = host_threads()
        rate, _, _ = cpu_arm(max(256, 32 * threads), threads)
        n = int(max(512, min(rate * 12.0, 4_000_000)))         # about 12 s of CPU work
        v, dt, _ = cpu_arm(n, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{n} problems of the bench shape, oracle C port with OpenMP, {dt:.1f} s"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def horizons_leg(eng, ocd, synthetic, rank, world_size, max_over_ranks, barrier, fp32_peak):
    """H = 15 and H = 50 (2 cars, the sweep's learning rates), device-resident, a few launches each: the
    segmented-adjoint kernels.  Same accounting as the headline: all ranks' solves / slowest rank's time."""
    import torch
    rows = []
    for H, B, lr in ((15, 262144, 0.03), (50, 65536, 0.003)):
        p = ocd.PlannerParams(H=H, C=2, lr=lr)
        b = synthetic.make_batch(B, seed=99 + rank)
        world = torch.as_tensor(b["world"], device=eng.device).permute(1, 2, 0).contiguous()
        w = torch.as_tensor(b["weights"], device=eng.device).t().contiguous()
        idx = torch.as_tensor(b["weight_idx"], device=eng.device)
        out = eng.solve_soa(p, world, w, w.shape[1], idx)
        barrier()
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.solve_soa(p, world, w, w.shape[1], idx, out=out)
        e1.record()
        barrier()
        ms = max_over_ranks(float(e0.elapsed_time(e1))) / reps
        fl = synthetic.flops_per_solve(H, 2, 3)
        rows.append({"horizon": H, "problems_per_gpu": B, "lr": lr, "ms_per_launch": ms,
                     "solves_per_sec": B * world_size / (ms * 1e-3),
                     "frac_of_measured_fp32": fl * B / (ms * 1e-3) / fp32_peak})
    return rows


def cmaes_leg(eng, ocd, dist, rank, world_size, max_over_ranks, barrier):
    """BASELINE configs[1]: finite_horizon cmaes --n_inits 5.  One generation = popsize 9 candidates
    x 5 initial states x 15 control steps, evaluated by ONE ocd_episode_batch launch; with N ranks
    the 45 episodes are sharded and the per-episode returns all-gathered (NCCL)."""
    import torch
    pop, n_inits, T = 9, 5, 15
    rng = np.random.default_rng(2024)
    w_true = np.array([-5, 0., 0., 0., -6., -50, -50])
    w_true = (w_true / np.linalg.norm(w_true)).astype(np.float32)
    cand = w_true[None] + 0.05 * rng.normal(size=(pop, 7))
    cand = (cand / np.linalg.norm(cand, axis=1, keepdims=True)).astype(np.float32)
    inits = np.stack([rng.uniform(-0.1, 0.1, n_inits), rng.uniform(-0.95, -0.85, n_inits),
                      rng.uniform(0.7, 0.9, n_inits), np.full(n_inits, np.pi / 2)], 1).astype(np.float32)
    ri = np.tile(inits, (pop, 1))
    widx = np.repeat(np.arange(pop), n_inits).astype(np.int32)
    B = pop * n_inits
    per = (B + world_size - 1) // world_size
    lo, hi = rank * per, min(B, (rank + 1) * per)
    p = ocd.PlannerParams()
    sc = ocd.Scenario(init_state=[[0.0, -0.6, 0.5, np.pi / 2]], kind=[0], friction=[0.0], control=[[0.0, 0.0]])
    dev = eng.device
    # shard padded to `per` problems so the all-gather is regular
    sel = np.arange(lo, lo + per) % B
    ris = torch.as_tensor(np.ascontiguousarray(ri[sel].T), device=dev)
    w = torch.as_tensor(np.ascontiguousarray(cand.T), device=dev)
    idx = torch.as_tensor(widx[sel], device=dev)
    tw = torch.as_tensor(w_true, device=dev)
    out = dict(returns=torch.empty((per,), dtype=torch.float32, device=dev))
    gathered = torch.empty((per * world_size,), dtype=torch.float32, device=dev)

    def generation():
        eng.episodes_soa(p, sc, ris, w, pop, tw, T, weight_idx=idx, out=out)
        if dist is not None:
            dist.all_gather_into_tensor(gathered, out["returns"])
        else:
            gathered.copy_(out["returns"])
        return gathered[:B].reshape(pop, n_inits).sum(1)

    for _ in range(3):
        generation()
    barrier()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        cand_returns = generation()
    e1.record()
    barrier()
    ms = max_over_ranks(float(e0.elapsed_time(e1))) / reps
    return {"workload": "finite_horizon cmaes --n_inits 5: one generation = 9 candidates x 5 inits x 15 control steps",
            "candidate_evals_per_sec": pop / (ms * 1e-3), "ms_per_generation": ms,
            "mpc_solves_per_generation": B * T, "episodes_per_rank": per,
            "collective": "all_gather of per-episode returns (NCCL)" if dist is not None else "none (1 GPU)",
            "first_candidate_return": float(cand_returns[0].item())}


if __name__ == "__main__":
    main()
