"""Mirror of interact_drive/car/fixed_plan_car.py:10-43 of the reference."""
from ...runtime import as_f32
from .car import Car


class FixedPlanCar(Car):
    """Replays a list of controls, then its default control (friction stays at the Car default 0.2)."""

    def __init__(self, env, init_state, plan, default_control=None, color: str = "gray", opacity=1.0, **kwargs):
        super().__init__(env, init_state, color, opacity, **kwargs)
        self.control_already_determined_for_current_step = True
        self.plan = [as_f32(u, (2,)) for u in plan]
        self.default_control = None if default_control is None else as_f32(default_control, (2,))
        self.control = self.default_control
        self.t = 0

    def step(self, dt):
        super().step(dt)
        self.t += 1
        self.set_next_control(self.plan[self.t] if self.t < len(self.plan) else self.default_control)

    def _get_next_control(self):
        return self.control

    def reset(self):
        super().reset()
        self.t = 0
        self.set_next_control(self.plan[self.t])

    def reward_fn(self, world_state, self_control):
        return 0
