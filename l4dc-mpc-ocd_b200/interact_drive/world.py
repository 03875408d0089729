"""Worlds and lanes: mirror of interact_drive/world.py of the reference (CarWorld :9-109,
ThreeLaneCarWorld :143-152, TwoLaneCarWorld :155-159, StraightLane :162-218)."""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import numpy as np


class StraightLane(object):
    """Lane whose median is the segment p -> q, of width w (reference world.py:162-218)."""

    def __init__(self, p: Tuple[float, float], q: Tuple[float, float], w: float):
        self.p, self.q, self.w = np.asarray(p, dtype=float), np.asarray(q, dtype=float), w
        d = self.q - self.p
        self.m = d / np.linalg.norm(d)                     # along the lane
        self.n = np.asarray([-self.m[1], self.m[0]])       # across the lane

    def shifted(self, n_lanes: int) -> "StraightLane":
        off = self.n * self.w * n_lanes
        return StraightLane(self.p + off, self.q + off, self.w)

    def dist2median(self, point) -> float:
        """Squared distance of (x, y, ...) to the median (reference world.py:206-218)."""
        r = (point[0] - self.p[0]) * self.n[0] + (point[1] - self.p[1]) * self.n[1]
        return r ** 2

    def on_road(self, point):
        raise NotImplementedError

    @property
    def median_x(self) -> float:
        """x of the median.  The engine's lane features assume lanes parallel to the y axis, which is
        what every world of the reference builds."""
        if abs(self.n[1]) > 1e-12:
            raise ValueError("the batched MPC engine supports lanes parallel to the y axis only")
        return float(self.p[0])


class CarWorld(object):
    """Container of cars and lanes with the two-phase step of the reference (world.py:79-109): every
    car fixes its control from the same past state, then every car integrates."""

    def __init__(self, dt: float = 0.1, lanes: Optional[List] = None, obstacles: Optional[List] = None,
                 visualizer_args: Optional[dict] = None, **kwargs):
        self.cars = []
        self.dt = dt
        self.lanes = [] if lanes is None else lanes
        self.obstacles = [] if obstacles is None else obstacles
        self.visualizer_args = dict() if visualizer_args is None else visualizer_args
        self.visualizer = None

    def add_car(self, car):
        car.index = len(self.cars)
        self.cars.append(car)

    def add_cars(self, cars: Iterable):
        for car in cars:
            self.add_car(car)

    @property
    def state(self):
        return [c.state for c in self.cars]

    @state.setter
    def state(self, new_state: Iterable):
        for c, x in zip(self.cars, new_state):
            c.state = x

    def reset(self):
        for car in self.cars:
            car.reset()

    def step(self, dt: Optional[float] = None):
        """-> (past_state, controls, state)."""
        past_state = self.state
        if dt is None:
            dt = self.dt
        for car in self.cars:
            if not car.control_already_determined_for_current_step:
                car.set_next_control()
        for car in self.cars:
            car.step(dt)
        return past_state, [c.control for c in self.cars], self.state

    def render(self, mode: str = "human", heatmap_show=False):
        raise NotImplementedError("rendering (pyglet visualizer) is outside the batched MPC engine's scope")

    # -- engine description ---------------------------------------------------------------------
    def lane_medians(self) -> Tuple[float, ...]:
        return tuple(lane.median_x for lane in self.lanes)


class ThreeLaneCarWorld(CarWorld):
    """Three straight lanes at x = -0.1, 0, +0.1 (reference world.py:143-152)."""

    def __init__(self, dt=0.1, **kwargs):
        lane = StraightLane((0.0, -5.), (0.0, 10.), 0.1)
        super().__init__(dt=dt, lanes=[lane.shifted(1), lane, lane.shifted(-1)], **kwargs)


class TwoLaneCarWorld(CarWorld):
    """Two straight lanes at x = -0.05, +0.05 (reference world.py:155-159)."""

    def __init__(self, dt=0.1, **kwargs):
        lane = StraightLane((-0.05, -5.), (-0.05, 10.), 0.1)
        super().__init__(dt=dt, lanes=[lane, lane.shifted(-1)], **kwargs)
