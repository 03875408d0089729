"""Planners: mirror of interact_drive/planner/ of the reference."""
from .car_planner import CarPlanner
from .naive_planner import NaivePlanner

__all__ = ["CarPlanner", "NaivePlanner"]
