"""Synthetic MPC problem batches of the shape BASELINE.json's sweep config names.

Robot: x~U(-0.1,0.1), y~U(-0.95,-0.85), v~U(0.7,1.1), theta=pi/2.  Other cars: x in the lane
medians, y = robot_y + U(0.1,0.6) (inside / near the collision bump), v~U(0.4,1.0), theta=pi/2.
Weights: N(0,1)^K normalised to unit L2 -- one vector per ``inits_per_candidate`` problems, selected
through ``weight_idx`` exactly like a CMA-ES population evaluated on several initial conditions.
"""
from __future__ import annotations

import numpy as np

LANES3 = (-0.1, 0.0, 0.1)


def make_batch(B: int, C: int = 2, lane_x=LANES3, inits_per_candidate: int = 5, seed: int = 1234):
    """-> dict(world [B, C, 4] f32, weights [Bw, K] f32, weight_idx [B] i32)."""
    rng = np.random.default_rng(seed)
    K = len(lane_x) + 4
    world = np.empty((B, C, 4), np.float32)
    world[:, 0, 0] = rng.uniform(-0.1, 0.1, B)
    world[:, 0, 1] = rng.uniform(-0.95, -0.85, B)
    world[:, 0, 2] = rng.uniform(0.7, 1.1, B)
    world[:, :, 3] = np.float32(np.pi / 2)
    lanes = np.asarray(lane_x, np.float32)
    for j in range(1, C):
        world[:, j, 0] = lanes[rng.integers(0, len(lanes), B)]
        world[:, j, 1] = world[:, 0, 1] + rng.uniform(0.1, 0.6, B).astype(np.float32)
        world[:, j, 2] = rng.uniform(0.4, 1.0, B)
    Bw = max(1, (B + inits_per_candidate - 1) // inits_per_candidate)
    w = rng.normal(size=(Bw, K))
    w /= np.linalg.norm(w, axis=1, keepdims=True)
    idx = (np.arange(B) // inits_per_candidate).astype(np.int32)
    return dict(world=world, weights=w.astype(np.float32), weight_idx=idx)


def make_other_controls(B: int, C: int, H: int, seed: int = 4321) -> np.ndarray:
    """Known controls of the other cars for other_mode=1: [B, C-1, H, 2]."""
    rng = np.random.default_rng(seed)
    return (rng.normal(size=(B, C - 1, H, 2)) * np.array([0.7, 2.0])).astype(np.float32)


def flops_per_solve(H: int, C: int, L: int, S: int = 3, n_iter: int = 100) -> float:
    """ALGORITHMIC FLOPs of one generate_plan: S*H*(n_iter*F_step + F_fwd)."""
    f_fwd = 44 + 6 * L + 20 * C
    f_step = 104 + 9 * L + 34 * C
    return float(S * H * (n_iter * f_step + f_fwd))


def hbm_bytes_per_solve(H: int, C: int, L: int, S: int = 3, inits_per_candidate: int = 5,
                        other_mode: int = 0) -> float:
    """ALGORITHMIC HBM bytes of one solve: inputs read once, outputs written once."""
    K = L + 4
    return 4.0 * (4 * C + K / inits_per_candidate + 1 + 2 * H * (C - 1) * other_mode + 2 * H + S + 1)
