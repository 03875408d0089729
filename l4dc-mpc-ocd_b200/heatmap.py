"""Reward heat-map: the grid of `reward_fn` evaluations the reference's visualizer draws
(interact_drive/visualizer.py:211-271 -- there 128 x 128 serial Python calls per frame), as ONE launch
of the feature kernel.  No rendering: the values are returned as an array."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import engine as _eng
from .runtime import as_f32, get_engine


def reward_heatmap(car, min_coord: Sequence[float], max_coord: Sequence[float], size=(128, 128),
                   weights: Optional[Sequence[float]] = None, zero_motion: bool = False) -> np.ndarray:
    """vals[j, i] = reward of `car` if it stood at (x_i, y_j) with its current speed and heading (or
    zero speed and heading, the reference's `car is None` branch), every other car where it is.
    Grid exactly as the reference's: linspace(min + 1e-6, max - 1e-6, size) per axis."""
    world = car.env
    xs = np.linspace(min_coord[0] + 1e-6, max_coord[0] - 1e-6, size[0])
    ys = np.linspace(min_coord[1] + 1e-6, max_coord[1] - 1e-6, size[1])
    order = [car.index] + [i for i in range(len(world.cars)) if i != car.index]
    base = np.stack([as_f32(world.cars[i].state, (4,)) for i in order])        # [C, 4], car first
    n = size[0] * size[1]
    states = np.broadcast_to(base, (n,) + base.shape).copy()
    gx, gy = np.meshgrid(xs, ys)                                               # [size1, size0]: y rows, x columns
    states[:, 0, 0] = gx.reshape(-1)
    states[:, 0, 1] = gy.reshape(-1)
    if zero_motion:
        states[:, 0, 2:] = 0.0
    p = _eng.PlannerParams(C=base.shape[0], lane_x=world.lane_medians(), num_lanes=int(car.num_lanes),
                           target_speed=float(car.target_speed), math_mode=_eng.MATH_PRECISE)
    phi = get_engine().features(p, states).cpu().numpy()                       # [n, K]
    w = car.weights_f32 if weights is None else as_f32(weights)
    return (phi @ w).reshape(size[1], size[0])
