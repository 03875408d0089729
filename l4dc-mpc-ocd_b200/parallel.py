"""Multi-GPU plumbing: one process per GPU, problems sharded on the batch axis.

Every (candidate x initial state x sample) episode -- and every start inside it -- is independent
for its whole length, so the only exchange step of the path is one all-gather of the per-episode
float32 returns per CMA-ES generation (4 bytes per episode; latency-bound).  Ranks then hold the
same returns and run the identical optimiser update, so nothing else is communicated."""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np
import torch
import torch.distributed as dist


_local_only = 0


def world() -> Tuple[int, int]:
    """(rank, world_size) of the default process group; (0, 1) outside torch.distributed, and inside `local_only()`."""
    if not _local_only and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def group() -> Tuple[int, int]:
    """(rank, world_size) of the default process group, `local_only()` or not."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class local_only:
    """Inside this context the sharding helpers see a world of one: work whose UNITS are spread over the ranks by the
    caller (independent optimisation runs, a rank's own subset each) is evaluated entirely on this rank's GPU, with no
    collective -- the caller exchanges results once, at the end."""

    def __enter__(self):
        global _local_only
        _local_only += 1
        return self

    def __exit__(self, *exc):
        global _local_only
        _local_only -= 1
        return False


def all_gather_objects(obj) -> list:
    """Every rank's `obj`, in rank order, on every rank (pickled; for end-of-run results, not for the data path)."""
    rank, ws = group()
    if ws == 1:
        return [obj]
    out = [None] * ws
    dist.all_gather_object(out, obj)
    return out


def shard_bounds(B: int, rank: int, world_size: int) -> Tuple[int, int, int]:
    """Contiguous shard [lo, hi) of B problems for `rank`, and the padded per-rank length `per`
    (every rank launches `per` problems so that the all-gather is regular)."""
    per = (B + world_size - 1) // world_size if B else 0
    lo = min(B, rank * per)
    hi = min(B, lo + per)
    return lo, hi, per


def shard_indices(B: int, rank: int, world_size: int) -> np.ndarray:
    """Problem indices this rank evaluates: its shard, padded by repeating the last problem of the
    batch (padding results are dropped after the gather)."""
    lo, hi, per = shard_bounds(B, rank, world_size)
    idx = np.arange(lo, lo + per)
    return np.minimum(idx, B - 1) if B else idx


def gather_returns(local: torch.Tensor, B: int) -> torch.Tensor:
    """local [per] on every rank -> the first B entries of the rank-ordered concatenation."""
    rank, ws = world()
    if ws == 1:
        return local[:B]
    out = torch.empty((local.numel() * ws,), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, local.contiguous())
    else:                                   # gloo (CPU tests)
        parts = [torch.empty_like(local) for _ in range(ws)]
        dist.all_gather(parts, local.contiguous())
        out = torch.cat(parts)
    return out[:B]


def sharded_returns(evaluate: Callable[[np.ndarray], torch.Tensor], B: int) -> torch.Tensor:
    """Run `evaluate(indices) -> returns [len(indices)]` on this rank's shard of range(B) and
    all-gather: every rank gets all B returns."""
    rank, ws = world()
    idx = shard_indices(B, rank, ws)
    return gather_returns(evaluate(idx), B)


def gather_rows(local: torch.Tensor, B: int) -> torch.Tensor:
    """local [per, F] on every rank -> the first B rows of the rank-ordered concatenation (every rank gets all)."""
    rank, ws = world()
    if ws == 1:
        return local[:B]
    flat = gather_returns(local.contiguous().reshape(-1), B * local.shape[1])
    return flat.reshape(B, local.shape[1])


def sharded_rows(evaluate: Callable[[np.ndarray], torch.Tensor], B: int) -> torch.Tensor:
    """Like sharded_returns for `evaluate(indices) -> [len(indices), F]`: per-episode rows (return and final world
    state), so that every rank ends with the same data and leaves the same Python-object state behind."""
    rank, ws = world()
    idx = shard_indices(B, rank, ws)
    return gather_rows(evaluate(idx), B)


def _visible_gpu_index(local_rank: int) -> int:
    import os
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            return local_rank
    return local_rank


def gpu_local_cpus(gpu_index: int):
    """CPUs on the NUMA node the GPU hangs off (NVML's ideal CPU affinity), or None when NVML cannot say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        import os
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        return cpus or None
    except Exception:
        return None


def plan_rank_cpus(allowed, local_world_size: int, affinities):
    """Pure planning step of bind_rank_cpus (unit-tested on CPU): `allowed` = CPUs the process may run on,
    `affinities[r]` = NUMA-local CPUs of rank r's GPU (or None).  Every rank gets a slice of ITS GPU's node, the node
    being shared evenly among the ranks whose GPUs hang off it; ranks whose node is unknown (or has fewer allowed CPUs
    than ranks) share what is left / everything, contiguously.  -> list of CPU lists, one per rank."""
    allowed = sorted(allowed)
    if local_world_size <= 1 or len(allowed) < local_world_size:
        return [list(allowed) for _ in range(max(1, local_world_size))]
    aset = set(allowed)
    keys = []
    for r in range(local_world_size):
        a = affinities[r] if affinities and r < len(affinities) else None
        node = tuple(sorted(set(a) & aset)) if a else ()
        keys.append(node)
    out = [None] * local_world_size
    for node in set(keys):
        ranks = [r for r in range(local_world_size) if keys[r] == node]
        if node and len(node) >= len(ranks) and len(node) < len(allowed):
            per = len(node) // len(ranks)
            for i, r in enumerate(ranks):
                out[r] = list(node[i * per:(i + 1) * per])
    if any(o is None for o in out):
        # no usable topology for some rank: contiguous slices of all allowed CPUs, by rank, for everyone (round 1's
        # rule; mixing the two rules could hand the same core to two ranks)
        per = len(allowed) // local_world_size
        out = [allowed[r * per:(r + 1) * per] for r in range(local_world_size)]
    return out


def bind_rank_cpus(local_rank: int, local_world_size: int) -> int:
    """One process per GPU on one host: give this rank its own slice of the CPUs the process may run on
    (sched_setaffinity) -- a slice of the NUMA node ITS GPU hangs off (NVML affinity), shared evenly with the other
    ranks on that node -- so that the ranks' staging-copy threads do not fight over the same cores and the pinned
    buffers this rank allocates afterwards (first touch) are node-local to the GPU's PCIe root.  The engine sizes its
    copy pool from the affinity mask.  Returns the number of CPUs this rank owns.  No-op when there is one rank,
    fewer CPUs than ranks, or no affinity support."""
    import os
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
    if local_world_size <= 1 or len(cpus) < local_world_size:
        return len(cpus)
    aff = [gpu_local_cpus(_visible_gpu_index(r)) for r in range(local_world_size)]
    mine = plan_rank_cpus(cpus, local_world_size, aff)[local_rank]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return len(cpus)
    return len(mine)
