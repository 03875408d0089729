#!/usr/bin/env python
"""Fixed-seed CMA-ES trajectories on the sphere and a rotated ellipsoid, N = 6 and 7 (the weight dimensions of the
three scenarios) -> tests/golden/cmaes_trajectories.json.

pycma (the reference's optimiser, interact_drive/reward_design/mpc_ord.py:41) is not installable on this image, so
these are REGRESSION fixtures of the restatement in l4dc-mpc-ocd_b200/cmaes.py, not pycma output: they pin the
restatement's sampling order, update equations and termination against accidental change.  The strategy parameters
are pinned separately against the closed forms of Hansen's tutorial (tests/test_host_logic_cpu.py).

    python tests/golden/make_cmaes_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from l4dc_mpc_ocd_b200 import cmaes        # noqa: E402


def sphere(x):
    return float(np.dot(x, x))


def ellipsoid(x):
    n = len(x)
    c = np.cos(0.3), np.sin(0.3)
    y = np.array(x, dtype=np.float64)
    y[0], y[1] = c[0] * x[0] - c[1] * x[1], c[1] * x[0] + c[0] * x[1]
    return float(sum((10.0 ** (3.0 * i / (n - 1))) * y[i] ** 2 for i in range(n)))


OBJ = {"sphere": sphere, "ellipsoid": ellipsoid}


def trajectory(name, N, seed, gens=25, sigma0=0.05):
    es = cmaes.CMAES(list(np.linspace(-0.5, 0.5, N)), sigma0, seed=seed)
    rows = []
    for _ in range(gens):
        pop = es.ask()
        fit = [OBJ[name](x) for x in pop]
        es.tell(fit)
        rows.append(dict(first_candidate=pop[0].tolist(), best_f=min(fit), sigma=es.sigma, mean=es.mean.tolist()))
    return dict(objective=name, N=N, seed=seed, sigma0=sigma0, generations=rows)


if __name__ == "__main__":
    out = [trajectory(n, N, seed) for n in ("sphere", "ellipsoid") for N in (6, 7) for seed in (1, 12345)]
    with open(Path(__file__).resolve().parent / "cmaes_trajectories.json", "w") as f:
        json.dump(out, f)
    print("wrote", len(out), "trajectories")
