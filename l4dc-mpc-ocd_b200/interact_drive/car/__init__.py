"""Car classes: mirror of interact_drive/car/ of the reference."""
from .car import Car
from .fixed_control_car import FixedControlCar
from .fixed_velocity_car import FixedVelocityCar
from .fixed_plan_car import FixedPlanCar
from .planner_car import PlannerCar
from .linear_reward_car import LinearRewardCar

__all__ = ["Car", "FixedControlCar", "FixedVelocityCar", "FixedPlanCar", "PlannerCar", "LinearRewardCar"]
