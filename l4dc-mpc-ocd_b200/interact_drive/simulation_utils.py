"""Car dynamics: mirror of interact_drive/simulation_utils.py:9-21, 24-123, 321-326 of the reference.

States are (x, y, vel, angle), controls (acc, ang_vel).  Every function runs the engine's dynamics
kernel (`ocd_dynamics_step_batch`) and returns host float32 arrays."""
from __future__ import annotations

import numpy as np

from ..runtime import as_f32, get_engine


def car_dynamics_step(x, y, v, angle, acc, ang_vel, dt, friction):
    """One step of the point-mass car with control clipping and quadratic friction
    (reference simulation_utils.py:9-21).  Scalars or equally shaped arrays; returns the four
    updated coordinates."""
    xs = [np.asarray(as_f32(a)) for a in (x, y, v, angle, acc, ang_vel)]
    shape = np.broadcast(*xs).shape
    st = np.stack([np.broadcast_to(a, shape).reshape(-1) for a in xs[:4]], axis=1)
    u = np.stack([np.broadcast_to(a, shape).reshape(-1) for a in xs[4:]], axis=1)
    fr = as_f32(friction)
    fr = float(fr) if fr.ndim == 0 else np.broadcast_to(fr, shape).reshape(-1)
    out = get_engine().dynamics(st, u, float(dt), fr).cpu().numpy()
    res = tuple(out[:, i].reshape(shape) for i in range(4))
    return tuple(np.float32(r) if r.ndim == 0 else r for r in res)


def batched_next_car_state(state, controls, dt, friction=0.2):
    """state [B, 4], controls [B, 2] -> next state [B, 4] (reference simulation_utils.py:24-70)."""
    state, controls = as_f32(state), as_f32(controls)
    if state.ndim != 2 or state.shape[-1] != 4:
        raise ValueError("Expected state to have shape (B, 4), got {} instead".format(state.shape))
    if controls.ndim != 2 or controls.shape != (state.shape[0], 2):
        raise ValueError("Expected controls to have shape (B, 2), got {} instead".format(controls.shape))
    fr = as_f32(friction)
    return get_engine().dynamics(state, controls, float(dt), float(fr) if fr.ndim == 0 else fr).cpu().numpy()


def next_car_state(state, controls, dt, friction=0.2):
    """state (4,), controls (2,) -> next state (4,) (reference simulation_utils.py:73-123;
    malformed shapes raise ValueError like :110-115)."""
    state, controls = as_f32(state), as_f32(controls)
    if state.shape != (4,):
        raise ValueError("Expected state to have shape (4,), got {} instead".format(state.shape))
    if controls.shape != (2,):
        raise ValueError("Expected controls to have shape (2,), got {} instead".format(controls.shape))
    return get_engine().dynamics(state[None], controls[None], float(dt), float(as_f32(friction))).cpu().numpy()[0]


def get_dynamics_fn(friction):
    """-> f(state, control, dt) with the friction baked in (reference simulation_utils.py:321-326)."""
    fr = float(as_f32(friction))

    def dynamics_fn(state, control, dt):
        return next_car_state(state, control, dt, fr)

    return dynamics_fn
