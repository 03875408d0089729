"""Base car: mirror of interact_drive/car/car.py:12-123 of the reference.  State (x, y, vel, angle)
and control (acc, ang_vel) are host float32 vectors; integration runs the engine's dynamics kernel."""
from __future__ import annotations

import numpy as np

from ...runtime import as_f32
from ..simulation_utils import get_dynamics_fn


class Car(object):
    def __init__(self, env, init_state, color: str = "gray", opacity: float = 1.0, friction: float = 0.2,
                 index: int = 0, debug: bool = False, **kwargs):
        self.env = env
        self.friction = friction
        self.dynamics_fn = get_dynamics_fn(friction)
        self.init_state = as_f32(init_state, (4,))
        self.state = self.init_state
        self.debug = debug
        self.past_traj = []
        self.color, self.opacity = color, opacity
        self.index = index
        self.control = None
        self.control_already_determined_for_current_step = False

    # states are always stored as float32 copies, whatever the caller assigns
    @property
    def init_state(self):
        return self._init_state

    @init_state.setter
    def init_state(self, value):
        self._init_state = as_f32(value, (4,))

    @property
    def state(self):
        return self._state

    @state.setter
    def state(self, value):
        self._state = as_f32(value, (4,))

    def reset(self):
        self.state = self.init_state
        if self.debug:
            self.past_traj = []

    def step(self, dt):
        """Integrate one tick with the latched control (reference car.py:76-87)."""
        if self.debug:
            self.past_traj.append((self.state, self.control))
        self.control_already_determined_for_current_step = False
        self.state = self.dynamics_fn(self.state, self.control, dt)

    def reward_fn(self, world_state, self_control):
        raise NotImplementedError

    def _get_next_control(self):
        raise NotImplementedError

    def set_next_control(self, control=None):
        """Latch `control`, or ask the car for one if none was fixed yet this tick (reference
        car.py:109-123)."""
        if control is not None:
            self.control = as_f32(control, (2,))
        elif not self.control_already_determined_for_current_step:
            self.control = self._get_next_control()
        self.control_already_determined_for_current_step = True
