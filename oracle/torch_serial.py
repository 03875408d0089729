"""float32 torch-CPU *autograd* restatement of the reference planner, driven the way the reference
drives TensorFlow.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): it is the `cpu_ref_serial`
baseline BASELINE.md section 3 names and an AD-independent cross-check of the C oracle's closed-form adjoint.

The reference solves ONE problem at a time in a Python loop with one tape-gradient call per SGD step
(interact_drive/planner/naive_planner.py:151-153) inside one process (`Pool(1)` unless --one_by_one,
experiments/run_mpc_ord.py:83-86).  TensorFlow is not installable on this image, so the same graph is
written against torch tensors here, op for op; nothing below is vectorised over problems on purpose.

Reference lines restated (paths relative to /root/reference):
  car_dynamics_step      interact_drive/simulation_utils.py:9-21
  _f / threshold / bump  interact_drive/math_utils.py:28-31, 92-95, 169-178
  features               experiments/merging.py:51-83, interact_drive/world.py:216-218
  reward_fn              interact_drive/car/linear_reward_car.py:49-55
  mpc_reward             interact_drive/planner/naive_planner.py:32-79
  generate_plan          interact_drive/planner/naive_planner.py:107-164
TF gradient conventions that differ from torch's defaults are spelled out where they occur.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np
import torch

F32 = torch.float32


def _c(x) -> torch.Tensor:
    return torch.tensor(x, dtype=F32)


def _tf_minimum(a, b):
    # tf.minimum routes the whole gradient to the first argument where a <= b (torch.minimum halves it on ties)
    return torch.where(a <= b, a, b)


def _tf_maximum(a, b):
    return torch.where(a >= b, a, b)


def car_dynamics_step(x, y, v, angle, acc, ang_vel, dt: float, friction: float):
    acc = _tf_maximum(_tf_minimum(acc, _c(4.0)), _c(-2 * 4.0))
    ang_vel = _tf_maximum(_tf_minimum(ang_vel, _c(4.0)), _c(-4.0))
    total_acc = acc - _c(friction) * v ** 2
    distance = v * _c(dt) + _c(0.5) * total_acc * _c(dt ** 2)
    return (x + torch.cos(angle) * distance, y + torch.sin(angle) * distance, v + total_acc * _c(dt),
            angle + ang_vel * _c(dt))


def _f(x, shape):
    x_clipped = torch.where(x > 0, x, torch.zeros_like(x) + 0.01)
    return torch.where(x > 0, torch.exp(-1 / (shape * x_clipped)), torch.zeros_like(x))


def smooth_threshold(threshold: float, width: float, c: float = 5.0):
    shape = _c(c / width)
    lo = _c(threshold - width)

    def t(x):
        x_diff = x - lo
        return _f(x_diff, shape) / (_f(x_diff, shape) + _f(_c(width) - x_diff, shape))

    return t


def smooth_bump(start, end):
    def bmp(x):
        width = (end - start) / 2
        center = (start + end) / 2
        x_norm = (x - center) / width
        cond = x_norm ** 2 < 1
        x_norm_clipped = torch.where(cond, x_norm, torch.zeros_like(x_norm))
        return torch.where(cond, torch.exp(-1 / (1 - x_norm_clipped ** 2) + 1), torch.zeros_like(x_norm))

    return bmp


class Problem:
    """One planner: the constants NaivePlanner / ThreeLaneTestCar / the world hold."""

    def __init__(self, H=5, lane_x: Sequence[float] = (-0.1, 0.0, 0.1), num_lanes=3, target_speed=1.0, dt=0.1,
                 friction=0.2, lr=0.1, n_iter=100, extra_inits=False):
        self.H, self.lane_x, self.num_lanes = H, [float(l) for l in lane_x], num_lanes
        self.ts = _c(np.float32(target_speed))
        self.dt, self.friction, self.lr, self.n_iter, self.extra_inits = dt, friction, lr, n_iter, extra_inits
        self.fence = smooth_threshold(0.05 * num_lanes, width=0.05)

    def features(self, state: torch.Tensor) -> torch.Tensor:
        """state [C, 4] (row 0 = the planning car) -> phi [L + 4]."""
        car = state[0]
        feats = []
        velocity = car[2] * torch.sin(car[3])
        feats.append(_tf_minimum((velocity - self.ts) ** 2, 4 * self.ts ** 2))
        lane_dists = []
        for lx in self.lane_x:      # StraightLane.dist2median with n = (-1, 0): ((x - p0) * -1 + (y - p1) * 0) ** 2
            r = (car[0] - _c(lx)) * _c(-1.0) + (car[1] - _c(0.0)) * _c(0.0)
            lane_dists.append(r ** 2 * 10)
        feats.extend(lane_dists)
        feats.append(torch.amin(torch.stack(lane_dists), dim=0))      # reduce_min: even split among ties (amin does too)
        coll = []
        for j in range(1, state.shape[0]):
            o = state[j]
            xb = smooth_bump(o[0] - 0.08, o[0] + 0.08)
            yb = smooth_bump(o[1] - 0.15, o[1] + 0.15)
            coll.append(xb(car[0]) * yb(car[1]))
        feats.append(torch.amax(torch.stack(coll), dim=0))
        feats.append((self.fence(car[0]) + self.fence(-car[0])) * torch.abs(car[0]))
        return torch.stack(feats)

    def mpc_reward(self, init_state: torch.Tensor, controls: Sequence[torch.Tensor], weights: torch.Tensor,
                   other_controls: Optional[torch.Tensor] = None) -> torch.Tensor:
        """init_state [C, 4]; controls: H tensors of shape (2,); other_controls [C-1, H, 2] or None."""
        world, dt = init_state, self.dt
        r = _c(0.0)
        for t in range(self.H):
            u = controls[t]
            new = []
            for i in range(world.shape[0]):
                x = world[i]
                if i == 0:
                    new.append(torch.stack(car_dynamics_step(x[0], x[1], x[2], x[3], u[0], u[1], dt, self.friction)))
                elif other_controls is not None:
                    v, ang = x[2], x[3]
                    a, w = other_controls[i - 1][t][0], other_controls[i - 1][t][1]
                    d = v * _c(dt) + _c(0.5) * a * _c(dt ** 2)
                    new.append(x + torch.stack([torch.cos(ang) * d, torch.sin(ang) * d, a * _c(dt), w * _c(dt)]))
                else:
                    v, ang = x[2], x[3]
                    new.append(x + torch.stack([torch.cos(ang) * v * _c(dt), torch.sin(ang) * v * _c(dt), _c(0.0), _c(0.0)]))
            world = torch.stack(new)
            r = r + torch.sum(weights * self.features(world))
        return r

    def generate_plan(self, world_state, weights, other_controls=None, cur_speed: Optional[float] = None):
        """-> (plan [H, 2], losses [S], best): the multi-start fixed-budget SGD of naive_planner.py:107-164."""
        init = torch.as_tensor(np.asarray(world_state, np.float32))
        w = torch.as_tensor(np.asarray(weights, np.float32))
        oc = None if other_controls is None else torch.as_tensor(np.asarray(other_controls, np.float32))
        turn = 5 * 0.13
        starts = [(0.0, 0.0), (0.0, -turn), (0.0, turn)]
        if self.extra_inits:
            v = float(init[0, 2]) if cur_speed is None else float(cur_speed)
            a0 = self.friction * v ** 2
            starts += [(a0, 0.0), (a0, -turn), (a0, turn)]
        losses, opts = [], []
        for a0, w0 in starts:
            planned = [torch.tensor([a0, w0], dtype=F32, requires_grad=True) for _ in range(self.H)]
            for _ in range(self.n_iter):                      # optimizer.minimize(loss, planned_controls): one tape per step
                loss = -self.mpc_reward(init, planned, w, oc)
                grads = torch.autograd.grad(loss, planned)
                with torch.no_grad():
                    for c, g in zip(planned, grads):
                        c -= _c(self.lr) * g                  # Keras SGD, momentum 0
            with torch.no_grad():
                losses.append(float(-self.mpc_reward(init, planned, w, oc)))
            opts.append(np.stack([c.detach().numpy().copy() for c in planned]))
        best = losses.index(min(losses))                      # Python min: first minimum
        return opts[best], np.asarray(losses, np.float32), best


def time_serial_solves(n: int, seed: int = 1234):
    """Time `n` bench-shape solves the reference's way: one at a time, one thread.  -> (solves/s, seconds)."""
    import importlib.util
    import time
    from pathlib import Path
    spec = importlib.util.spec_from_file_location(
        "_ocd_synthetic", Path(__file__).resolve().parent.parent / "l4dc-mpc-ocd_b200" / "synthetic.py")
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    b = syn.make_batch(n, seed=seed)
    prob = Problem()
    old = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        t0 = time.perf_counter()
        for i in range(n):
            prob.generate_plan(b["world"][i], b["weights"][b["weight_idx"][i]])
        dt = time.perf_counter() - t0
    finally:
        torch.set_num_threads(old)
    return n / dt, dt
