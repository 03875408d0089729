"""Mirror of interact_drive/car/fixed_velocity_car.py:18-24 of the reference."""
from .fixed_control_car import FixedControlCar


class FixedVelocityCar(FixedControlCar):
    """Keeps the velocity of its initial state: zero control, zero friction."""

    def __init__(self, env, init_state, color: str = "gray", opacity=1.0, **kwargs):
        kwargs.pop("friction", None)
        super().__init__(env, init_state, [0.0, 0.0], color, opacity, friction=0.0, **kwargs)
