"""The N>1 path on CPU: two gloo ranks shard a batch of episodes, each evaluates its shard (here
with the CPU oracle standing in for the GPU kernel -- the plumbing under test is the sharding and
the all-gather of returns), and every rank must end up with the full, correctly ordered vector."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle as O
    from l4dc_mpc_ocd_b200 import parallel
    spec = O.scenario_params("finite_horizon")
    w = (spec.designer_weights / np.linalg.norm(spec.designer_weights)).astype(np.float32)
    rng = np.random.default_rng(5)
    ri = np.tile(spec.example_init, (B, 1)).astype(np.float32)
    ri[:, 0] += rng.uniform(-0.02, 0.02, B).astype(np.float32)
    assert parallel.world() == (rank, ws)

    def evaluate(idx):
        return torch.from_numpy(O.episode_batch(spec.params, spec.scenario, ri[idx], w, w, 4, nthreads=1))

    full = parallel.sharded_returns(evaluate, B).numpy()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full)
    if rank == 0:
        np.save(os.path.join(out_dir, "ref.npy"), O.episode_batch(spec.params, spec.scenario, ri, w, w, 4, nthreads=1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_returns(tmp_path):
    B, ws = 11, 2            # not divisible: the last rank pads
    mp.spawn(_worker, args=(ws, _free_port(), B, str(tmp_path)), nprocs=ws, join=True)
    ref = np.load(tmp_path / "ref.npy")
    for r in range(ws):
        got = np.load(tmp_path / f"r{r}.npy")
        assert got.shape == (B,)
        np.testing.assert_array_equal(got, ref)
