"""ThreeLaneTestCar: mirror of experiments/merging.py:20-83 of the reference -- the linear-in-features
reward every scenario optimises.  The feature arithmetic is the engine's (`ocd_features_batch`)."""
from __future__ import annotations

import numpy as np

from ..interact_drive.car import LinearRewardCar, PlannerCar
from ..interact_drive.world import CarWorld, ThreeLaneCarWorld   # noqa: F401  (re-exported like the reference)
from ..runtime import as_f32, get_engine
from .. import engine as _eng


class ThreeLaneTestCar(LinearRewardCar, PlannerCar):
    """Planning car with features [bounded squared speed error; 10 x squared distance to each lane;
    their minimum; collision bump (max over the other cars); fence]."""

    engine_features = True      # tells NaivePlanner that the kernels implement this car's features

    def __init__(self, env, init_state, horizon: int, weights, target_speed=1., color="orange", friction=0.2,
                 opacity=1.0, planner_args=None, debug=False, num_lanes=3, **kwargs):
        super().__init__(env, init_state, horizon=horizon, weights=weights, color=color, friction=friction,
                         opacity=opacity, planner_args=planner_args, debug=debug, **kwargs)
        self.target_speed = np.float32(target_speed)
        self.num_lanes = num_lanes

    def features(self, state, control=None):
        """state: one (4,) vector per car of the world (or [C, 4]); `control` is unused, as in the
        reference.  -> float32 [L + 4]."""
        st = np.stack([as_f32(s, (4,)) for s in state])
        order = [self.index] + [i for i in range(st.shape[0]) if i != self.index]
        st = st[order]
        if st.shape[0] == 1:
            raise ValueError("ThreeLaneTestCar.features needs at least one other car (the reference's "
                             "reduce_max over an empty collision list is undefined)")
        p = _eng.PlannerParams(C=st.shape[0], lane_x=self.env.lane_medians(), num_lanes=int(self.num_lanes),
                               target_speed=float(self.target_speed), math_mode=_eng.MATH_PRECISE)
        return get_engine().features(p, st[None]).cpu().numpy()[0]
