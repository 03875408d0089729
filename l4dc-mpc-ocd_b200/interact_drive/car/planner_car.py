"""Mirror of interact_drive/car/planner_car.py:12-90 of the reference."""
import numpy as np

from .car import Car


class PlannerCar(Car):
    """Runs an MPC planner at every tick and applies the first control of the plan."""

    def __init__(self, env, init_state, horizon: int, color: str = "orange", opacity: float = 1.0,
                 friction: float = 0.2, planner_args: dict = None, check_plans: bool = False, **kwargs):
        super().__init__(env, init_state, color=color, opacity=opacity, friction=friction, **kwargs)
        self.horizon = horizon
        self.planner = None
        self.plan = []
        self.planner_args = {} if planner_args is None else planner_args
        self.check_plans = check_plans

    def initialize_planner(self, planner_args):
        from ..planner.naive_planner import NaivePlanner
        self.planner = NaivePlanner(self.env, self, self.horizon, **planner_args)

    def known_other_plans(self):
        """What `check_plans` feeds the planner (reference planner_car.py:58-80): for every other car
        its plan replayed FROM INDEX 0 (whatever its own clock says), then its default control, zeros
        for cars without a plan.  One [H, 2] array per car of the world; the own entry is zeros."""
        zero = np.zeros(2, np.float32)
        plans = []
        for i, other in enumerate(self.env.cars):
            rows = []
            for j in range(self.horizon):
                u = zero
                if i != self.index and getattr(other, "plan", None) is not None:
                    if j < len(other.plan):
                        u = other.plan[j]
                    elif getattr(other, "default_control", None) is not None:
                        u = other.default_control
                rows.append(np.asarray(u, np.float32))
            plans.append(np.stack(rows))
        return plans

    def _get_next_control(self):
        if self.planner is None:
            self.initialize_planner(self.planner_args)
        if self.check_plans:
            self.plan = self.planner.generate_plan(other_controls=self.known_other_plans())
        else:
            self.plan = self.planner.generate_plan()
        return np.array(self.plan[0], dtype=np.float32)
