"""Multi-GPU path on real devices (skipped on boxes with fewer than 2 GPUs): the CLI under torchrun
shards every generation's episodes over the ranks, all-gathers the returns with NCCL, and must produce
exactly the history a single-GPU run produces."""
import os
import pickle
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run_cli(tmp, nproc, tag):
    out = tmp / f"hist_{tag}.pkl"
    code = (
        "import sys, pickle; sys.path.insert(0, %r)\n"
        "from l4dc_mpc_ocd_b200.experiments import run_mpc_ord as r\n"
        "import torch.distributed as dist\n"
        "m = r.main(['replanning', 'cmaes', '--n_inits', '3', '--seed', '1', '--opt_seed', '5', '--max_evals', '16',\n"
        "            '--quiet', '--no_save'])[0]\n"
        "rank = dist.get_rank() if dist.is_initialized() else 0\n"
        "if rank == 0:\n"
        "    pickle.dump([(np_w.tolist(), float(v)) for np_w, v in m.history], open(%r, 'wb'))\n"
        "if dist.is_initialized():\n"
        "    dist.barrier(); dist.destroy_process_group()\n" % (str(ROOT), str(out)))
    script = tmp / f"run_{tag}.py"
    script.write_text(code)
    if nproc == 1:
        cmd = [sys.executable, str(script)]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    env = dict(os.environ)
    env.pop("RANK", None), env.pop("WORLD_SIZE", None), env.pop("LOCAL_RANK", None)
    subprocess.run(cmd, check=True, cwd=str(ROOT), env=env, timeout=600, capture_output=True)
    return pickle.load(open(out, "rb"))


def test_two_gpu_cmaes_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    one = _run_cli(tmp_path, 1, "one")
    two = _run_cli(tmp_path, 2, "two")
    assert len(one) == len(two) == 1 + 2 * 9        # designer weights + two generations of popsize 9 (N=6)
    for (w1, v1), (w2, v2) in zip(one, two):
        np.testing.assert_array_equal(w1, w2)
        assert v1 == v2                               # same kernel, same problems: bit-identical returns
