"""CPU oracle for the L4DC-MPC-OCD hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and there only as the checker or the
timed CPU baseline.  The product (``l4dc-mpc-ocd_b200/``) never imports it.

Parity status: pinned against the reference's known-answer tests and against golden
vectors produced by the reference's unmodified Python on a torch-backed TensorFlow
stand-in (``oracle/tf_shim``); unpinned against real TensorFlow numerics, ``cma``,
``scipy.stats.truncnorm`` (version-dependent) and TFP L-BFGS.  See ``ocd_oracle.h``.
"""
from .oracle import (  # noqa: F401
    OracleParams,
    OracleScenario,
    build,
    lib,
    dynamics_step,
    smooth_f,
    smooth_threshold,
    smooth_bump,
    features,
    mpc_reward,
    generate_plan,
    generate_plan_batch,
    episode,
    episode_batch,
    max_threads,
    scenario_params,
    SCENARIOS,
)
