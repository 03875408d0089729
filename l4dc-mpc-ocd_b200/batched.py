"""World -> engine description: turns a drop-in (world, planning car) pair into the POD structs the
episode kernel consumes, so that MPC_ORD can evaluate whole populations of weight vectors on whole
sets of initial states in one launch instead of stepping Python objects."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import engine as _eng
from .interact_drive.car import FixedControlCar, FixedPlanCar, PlannerCar


@dataclass
class WorldProgram:
    params: "_eng.PlannerParams"
    scenario: "_eng.Scenario"
    replanning: bool            # world toggles `unlucky_car_idx` on reset and teleports at critical_t


def compile_world(world, car, math_mode: Optional[int] = None) -> WorldProgram:
    """Describe `world` as seen from planning car `car`.

    Supported: the planning car is car 0 (as in every reference scenario) and is a ThreeLaneTestCar;
    every other car is a FixedControlCar / FixedVelocityCar / FixedPlanCar.  Anything else raises
    TypeError -- such worlds still run through CarWorld.step(), one launch per control step."""
    if car.index != 0 or world.cars[0] is not car:
        raise TypeError("compile_world: the planning car must be car 0 of the world")
    if not getattr(car, "engine_features", False) or not isinstance(car, PlannerCar):
        raise TypeError("compile_world: the planning car must be a ThreeLaneTestCar")
    args = dict(car.planner_args or {})
    unknown = set(args) - {"learning_rate", "n_iter", "extra_inits", "leaf_evaluation", "math_mode", "engine"}
    if unknown or args.get("leaf_evaluation") is not None:
        raise TypeError("compile_world: unsupported planner_args %s" % sorted(unknown))
    kind, fric, ctrl, plans, init = [], [], [], [], []
    for other in world.cars[1:]:
        if isinstance(other, FixedPlanCar):
            kind.append(1)
            plans.append([np.asarray(u, np.float32) for u in other.plan])
            dc = other.default_control
            ctrl.append(np.zeros(2, np.float32) if dc is None else np.asarray(dc, np.float32))
        elif isinstance(other, FixedControlCar):
            kind.append(0)
            plans.append([])
            ctrl.append(np.asarray(other.control, np.float32))
        else:
            raise TypeError("compile_world: car %d is a %s; only scripted cars can ride along a batched episode"
                            % (other.index, type(other).__name__))
        fric.append(float(other.friction))
        init.append(np.asarray(other.init_state, np.float32))
    mm = args.get("math_mode", _eng.MATH_FAST) if math_mode is None else math_mode
    params = _eng.PlannerParams(
        H=int(car.horizon), C=len(world.cars), lane_x=world.lane_medians(), n_iter=int(args.get("n_iter", 100)),
        num_lanes=int(car.num_lanes), other_mode=1 if car.check_plans else 0,
        extra_inits=bool(args.get("extra_inits", False)), math_mode=mm,
        lr=float(args.get("learning_rate", 0.1)), dt=float(world.dt), friction=float(car.friction),
        target_speed=float(car.target_speed))
    replanning = hasattr(world, "critical_t") and hasattr(world, "unlucky_car_idx")
    scenario = _eng.Scenario(init_state=init, kind=kind, friction=fric, control=ctrl, plan=plans,
                             critical_t=int(world.critical_t) if replanning else 0)
    return WorldProgram(params, scenario, replanning)


def unlucky_sequence(world, n_resets: int) -> List[int]:
    """The `unlucky_car_idx` values the next `n_resets` world.reset() calls would produce
    (ReplanningCarWorld toggles 1 <-> 2 on every reset), advancing the world's own toggle."""
    seq = []
    cur = world.unlucky_car_idx
    for _ in range(n_resets):
        cur = 2 if cur == 1 else 1
        seq.append(cur)
    world.unlucky_car_idx = cur
    return seq
