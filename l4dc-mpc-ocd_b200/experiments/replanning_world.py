"""The replanning scenario: mirror of experiments/replanning_world.py:11-95 of the reference."""
from typing import Optional

import numpy as np

from ..interact_drive.car import FixedPlanCar
from ..interact_drive.reward_design.mpc_ord import sample_init_state
from ..interact_drive.world import TwoLaneCarWorld
from .merging import ThreeLaneTestCar


class ReplanningCarWorld(TwoLaneCarWorld):
    """One of the two other cars vanishes (is teleported far away) at tick `critical_t`; which one
    alternates with every reset."""

    def __init__(self, dt=0.1, critical_t=4, **kwargs):
        super().__init__(dt=dt, **kwargs)
        self.critical_t = critical_t
        self.unlucky_car_idx = 1
        self.t = 0

    def reset(self):
        super().reset()
        self.unlucky_car_idx = 2 if self.unlucky_car_idx == 1 else 1
        self.t = 0

    def step(self, dt: Optional[float] = None):
        self.t += 1
        if self.t == self.critical_t:
            self.cars[self.unlucky_car_idx].state = np.array([10., 0., 0., 0.], np.float32)
        return super().step()


og_weights = np.array([-3, 0, 0, -2, -10, -10], dtype=np.float32)
og_weights /= np.linalg.norm(og_weights)
tuned_weights = np.array([-0.55899817, -0.4436692, -0.37245109, -0.19964276, -0.5438697, 0.12770044], dtype=np.float32)
tuned_weights /= np.linalg.norm(tuned_weights)


def setup_world(env_seeds=[1], debug=True):
    """Two-lane road; two FixedPlanCars start side by side ahead of the planning car and swerve to
    opposite sides; the planner is told their plans (check_plans).  -> (car, world, init_states)."""
    init_states = [sample_init_state(s, (-0.0, 0.02, (-0.005, 0.005)), (-0.9, 0.04, (-1., -0.8)), (1.0, 0.05, (0.8, 1.2)))
                   for s in env_seeds]
    world = ReplanningCarWorld()
    our_car = ThreeLaneTestCar(world, init_states[0], horizon=5, weights=og_weights, target_speed=1.2,
                               planner_args={'n_iter': 100}, check_plans=True, num_lanes=2, debug=debug)

    def swerving_car(turn):
        plan = [[0., 0.], [0.7, turn], [0., 0.], [0.0, -turn]]
        return FixedPlanCar(world, np.array([0., -0.7, 0.8, np.pi / 2]), plan=plan, default_control=[0.0, 0.0],
                            color='gray', opacity=0.8, debug=debug)

    world.add_cars([our_car, swerving_car(2.7), swerving_car(-2.7)])
    world.reset()
    return our_car, world, init_states
