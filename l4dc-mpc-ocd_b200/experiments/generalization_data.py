"""Generalisation evaluation: the core of experiments/generalization_data.py:27,64-107 of the reference.

The reference re-evaluates, on 32 held-out initial states (environment seeds 2**32-i-1), the best
weights of every CMA-ES history (n_inits in {1,3,5,7} x 6-7 seeds), with four successive
`multiprocessing.Pool(8)` maps, one worker per test state.  Here all (weights x test states x samples)
episodes are one launch."""
from __future__ import annotations

import pickle
from typing import Dict, Iterable, Sequence

import numpy as np

from ..interact_drive.reward_design.mpc_ord import MPC_ORD
from . import run_mpc_ord

NUM_CMAES_EVALS = 85          # evaluations of each history the reference keeps (generalization_data.py:15)


def make_test_env(scenario: str, n_test: int = 40):
    """(car, world, test_inits): the held-out initial states of the reference (seeds 2**32-i-1)."""
    env = run_mpc_ord.envs[scenario]
    return env['make_env'](env_seeds=[2 ** 32 - i - 1 for i in range(n_test)])


def best_weights(history: Sequence, num_evals: int = NUM_CMAES_EVALS) -> np.ndarray:
    """The weights a history would have selected after its first `num_evals` evaluations."""
    return np.asarray(max(list(history)[:num_evals], key=lambda a: a[1])[0])


def evaluate_on_test_inits(scenario: str, weights: Dict, n_test: int = 32, verbose: bool = False) -> Dict:
    """weights: {key: weight vector} (the reference keys by (n_inits, seed)).
    -> {test_init_index: {key: return}} like the per-init result pickles of the reference, where return is
    `MPC_ORD.eval_weights_for_init` (summed over the samples)."""
    env = run_mpc_ord.envs[scenario]
    car, world, test_inits = make_test_env(scenario, max(n_test, 1))
    bord = MPC_ORD(world, car, [], env['eval_horizon'], num_samples=env['num_eval_samples'], verbose=verbose)
    keys = list(weights)
    ret = bord.episode_returns([weights[k] for k in keys], test_inits[:n_test])      # [n_w, n_test, n_samples]
    per_init = ret.sum(axis=2)
    return {i: {k: float(per_init[j, i]) for j, k in enumerate(keys)} for i in range(n_test)}


def load_histories(files: Dict) -> Dict:
    """{key: path of a history pickle written by MPC_ORD.save_history} -> {key: best weights}."""
    out = {}
    for k, fn in files.items():
        with open(fn, 'rb') as f:
            out[k] = best_weights(pickle.load(f))
    return out
