"""NaivePlanner: mirror of interact_drive/planner/naive_planner.py:15-164 of the reference.

The reference builds a tf.function for the H-step rollout reward and runs, for each of 3 (or 6)
fixed initial control sequences, `n_iter` Keras-SGD steps on its negation from Python, then keeps the
start with the smallest final loss.  Here the whole of that -- rollout, reward, reverse-mode gradient,
the SGD loop over all starts and the argmin -- is ONE kernel launch (`ocd_solve_batch`), for one world
(`generate_plan`, the reference's signature) or for a batch of worlds (`generate_plan_batch`).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from ... import engine as _eng
from ...runtime import as_f32, get_engine
from .car_planner import CarPlanner

_PHANTOM = np.array([1.0e3, 1.0e3, 0.0, np.pi / 2], np.float32)   # a car nobody can collide with


class NaivePlanner(CarPlanner):
    """MPC planner that assumes the other cars keep their velocity (or follow known controls)."""

    def __init__(self, world, car, horizon: int, learning_rate: float = 0.1, n_iter: int = 100,
                 leaf_evaluation=None, extra_inits=False, *, math_mode: Optional[int] = None, engine=None):
        super().__init__(world, car)
        if leaf_evaluation is not None:
            raise NotImplementedError("leaf_evaluation is never set by the reference's drivers and is not "
                                      "built into the kernels")
        self.leaf_evaluation = None
        self.learning_rate = learning_rate
        self.horizon = horizon
        self.n_iter = n_iter
        self.extra_inits = extra_inits
        self.math_mode = _eng.MATH_FAST if math_mode is None else math_mode
        self._engine = engine
        self.planned_controls = [np.zeros(2, np.float32) for _ in range(horizon)]
        self.reward_func = self.mpc_reward
        self.last_losses = None
        self.last_best = None

    # -- what the engine needs to know about the world ---------------------------------------------
    @property
    def engine(self):
        return self._engine if self._engine is not None else get_engine()

    def _order(self) -> List[int]:
        """Engine car order: the planning car first, the others in world order."""
        n = len(self.world.cars)
        me = self.car.index
        return [me] + [i for i in range(n) if i != me]

    def params(self, other_mode: int = 0) -> "_eng.PlannerParams":
        car = self.car
        if not getattr(car, "engine_features", False):
            raise TypeError("NaivePlanner: the reward features are built into the CUDA kernels; the planning "
                            "car must be a ThreeLaneTestCar (experiments/merging.py), got %s" % type(car).__name__)
        return _eng.PlannerParams(
            H=self.horizon, C=max(2, len(self.world.cars)), lane_x=self.world.lane_medians(), n_iter=self.n_iter,
            num_lanes=int(car.num_lanes), other_mode=other_mode, extra_inits=bool(self.extra_inits),
            math_mode=self.math_mode, lr=self.learning_rate, dt=self.world.dt, friction=float(car.friction),
            target_speed=float(car.target_speed))

    def _world_array(self, init_state) -> np.ndarray:
        """[C, 4] in engine order; a lone planning car gets a far-away phantom partner."""
        if init_state is None:
            init_state = self.world.state
        st = np.stack([as_f32(s, (4,)) for s in init_state])
        if st.shape[0] != len(self.world.cars):
            raise ValueError("init_state must hold one (4,) state per car of the world")
        st = st[self._order()]
        if st.shape[0] == 1:
            st = np.concatenate([st, _PHANTOM[None]])
        return st

    def _other_controls(self, other_controls) -> Optional[np.ndarray]:
        """reference layout (one [H, >=2] sequence per car, own entry ignored) -> [C-1, H, 2]."""
        if other_controls is None:
            return None
        if len(other_controls) != len(self.world.cars):
            raise ValueError("other_controls must hold one control sequence per car of the world")
        rows = [as_f32(other_controls[i])[: self.horizon, :2] if i != self.car.index else None
                for i in range(len(self.world.cars))]
        rows = [r for r in rows if r is not None]
        if not rows:
            rows = [np.zeros((self.horizon, 2), np.float32)]        # the phantom partner
        return np.stack(rows).reshape(len(rows), self.horizon, 2)

    def _weights(self, weights) -> np.ndarray:
        return self.car.weights_f32 if weights is None else as_f32(weights)

    # -- reference API -------------------------------------------------------------------------------
    def mpc_reward(self, init_state, controls, other_controls=None, weights=None):
        """Sum over the horizon of reward_fn(world_state_{t+1}, control_t) (reference :32-79)."""
        oc = self._other_controls(other_controls)
        p = self.params(other_mode=0 if oc is None else 1)
        u = np.stack([as_f32(c, (2,)) for c in controls])
        R = self.engine.reward(p, self._world_array(init_state)[None], u[None], self._weights(weights),
                               other_controls=None if oc is None else oc[None], grad=False)
        return np.float32(R.cpu().numpy()[0])

    def mpc_reward_and_grad(self, init_state, controls, other_controls=None, weights=None):
        """The reward and d reward / d controls [H, 2] -- what tf.GradientTape gives the reference."""
        oc = self._other_controls(other_controls)
        p = self.params(other_mode=0 if oc is None else 1)
        u = np.stack([as_f32(c, (2,)) for c in controls])
        R, G = self.engine.reward(p, self._world_array(init_state)[None], u[None], self._weights(weights),
                                  other_controls=None if oc is None else oc[None], grad=True)
        return np.float32(R.cpu().numpy()[0]), G.cpu().numpy()[0]

    def generate_plan(self, init_state=None, weights=None, other_controls=None, use_lbfgs=False):
        """-> list of H controls (2,) float32: the start with the smallest final loss after n_iter
        gradient steps (reference :81-164).  init_state None = the world's current state; weights None =
        the car's own normalised weights."""
        oc = self._other_controls(other_controls)
        p = self.params(other_mode=0 if oc is None else 1)
        if use_lbfgs:
            # The reference's branch (:127-149) hands each start to TFP's lbfgs_minimize with
            # max_iterations=200; it is never enabled by its drivers and mis-binds `weights` as
            # other_controls.  Here the flag selects the engine's own L-BFGS (same role, its own line
            # search -- see include/ocd_b200.h), with the arguments bound as the SGD branch binds them.
            if self.horizon > _eng.N.LBFGS_MAX_H:
                raise NotImplementedError("use_lbfgs: the L-BFGS kernel supports horizon <= %d" % _eng.N.LBFGS_MAX_H)
            p.optimizer, p.n_iter = _eng.OPT_LBFGS, 200
        cur_speed = [float(self.car.state[2])]                      # live speed, not init_state's (reference :114)
        res = self.engine.solve(p, self._world_array(init_state)[None], self._weights(weights),
                                other_controls=None if oc is None else oc[None], cur_speed=cur_speed)
        plan = res["plan"].cpu().numpy()[0]
        self.last_losses = res["losses"].cpu().numpy()[0]
        self.last_best = int(res["best"].cpu().numpy()[0])
        self.planned_controls = [plan[t].copy() for t in range(self.horizon)]
        return self.planned_controls

    # -- batched extension ---------------------------------------------------------------------------
    def feature_jacobian_batch(self, init_states, controls, other_controls=None):
        """init_states [B, C, 4] in WORLD car order, controls [B, H, 2] of the planning car ->
        (phi_sum [B, K], jac [B, K, H, 2]) as host arrays: the horizon-summed features and their Jacobian
        with respect to the controls (what the reference's IOC classes take from tf.GradientTape,
        reward_design/first_order_ioc.py:86-91, :213-240); K launches of the reward-gradient kernel."""
        st = as_f32(init_states)
        if st.ndim != 3 or st.shape[1] != len(self.world.cars) or st.shape[2] != 4:
            raise ValueError("init_states must have shape [B, %d, 4]" % len(self.world.cars))
        st = st[:, self._order()]
        if st.shape[1] == 1:
            st = np.concatenate([st, np.broadcast_to(_PHANTOM, (st.shape[0], 1, 4))], axis=1)
        p = self.params(other_mode=0 if other_controls is None else 1)
        phi, jac = self.engine.feature_jacobian(p, st, as_f32(controls), other_controls=other_controls)
        return phi.cpu().numpy(), jac.cpu().numpy()

    def feature_hessian_batch(self, init_states, controls, other_controls=None):
        """init_states [B, C, 4] in WORLD car order, controls [B, H, 2] of the planning car -> hess [B, K, 2H, 2H]
        (host array): the Hessian of every horizon-summed feature with respect to the flattened controls -- what the
        reference's LocalCIOC takes from `t.jacobian(gradients, controls)` (reward_design/second_order_ioc.py:143-147);
        one launch."""
        st = as_f32(init_states)
        if st.ndim != 3 or st.shape[1] != len(self.world.cars) or st.shape[2] != 4:
            raise ValueError("init_states must have shape [B, %d, 4]" % len(self.world.cars))
        st = st[:, self._order()]
        if st.shape[1] == 1:
            st = np.concatenate([st, np.broadcast_to(_PHANTOM, (st.shape[0], 1, 4))], axis=1)
        p = self.params(other_mode=0 if other_controls is None else 1)
        return self.engine.feature_hessian(p, st, as_f32(controls), other_controls=other_controls).cpu().numpy()

    def generate_plan_batch(self, init_states, weights=None, weight_idx=None, other_controls=None, cur_speed=None):
        """init_states [B, C, 4] in WORLD car order; weights [K] | [Bw, K] (+ weight_idx [B]);
        other_controls [B|1, C-1, H, 2] for the non-planning cars in world order.
        -> dict(plan [B, H, 2], losses [B, S], best [B]) as host arrays; one launch."""
        st = as_f32(init_states)
        if st.ndim != 3 or st.shape[1] != len(self.world.cars) or st.shape[2] != 4:
            raise ValueError("init_states must have shape [B, %d, 4]" % len(self.world.cars))
        st = st[:, self._order()]
        if st.shape[1] == 1:
            st = np.concatenate([st, np.broadcast_to(_PHANTOM, (st.shape[0], 1, 4))], axis=1)
        p = self.params(other_mode=0 if other_controls is None else 1)
        res = self.engine.solve(p, st, self._weights(weights), weight_idx=weight_idx,
                                other_controls=other_controls, cur_speed=cur_speed)
        return {k: v.cpu().numpy() for k, v in res.items()}
