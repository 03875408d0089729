"""First-order inverse optimal control: mirror of interact_drive/reward_design/first_order_ioc.py
(InverseLocallyOptimalControl :16-168, LinearInverseLocallyOptimalControl :171-314) and
inverse_optimal_control.py:9-48 of the reference.

Both classes fit reward weights w to an observed trajectory through the first-order optimality condition
d r / d u = 0.  For the linear reward r = w . phi that gradient is J^T w with J = d(sum_t phi)/du, and J does
not depend on w.  The reference re-derives the gradient with tf.GradientTape inside every one of its Adam
steps; here J comes from ONE batched launch (`ocd_feature_jacobian_batch`, all sliding windows of the
trajectory at once) and the optimisation over w is a K-dimensional host computation on that matrix.

What differs from the reference, and why:
  * the reward features are the ones built into the kernels (experiments/merging.py), so the planning
    car must be a ThreeLaneTestCar -- the reference's own IOC tests use a two-feature
    LinearTargetSpeedPlannerCar that the engine does not have;
  * `segment_loss` in the reference calls `reward_func(initial_state, controls, weights)`, which binds the
    weights to `other_controls` (naive_planner.py:32) -- bit-rot.  The intended binding is used here;
  * other cars are predicted by the planner's constant-velocity model, which is what
    `segment_jacobian` (:213-240) does by hand (zero control, zero friction).
Parity for these classes is therefore UNPINNED against the reference; the Jacobian itself is checked
against the CPU oracle (tests/test_gpu_ioc.py) and the optimisers by their own properties.
"""
from __future__ import annotations

from typing import Collection, List, Optional, Sequence, Tuple

import numpy as np

from ...runtime import as_f32
from ..car import PlannerCar
from ..planner import NaivePlanner


class InverseOptimalControl(object):
    """Abstract class for inverse optimal control algorithms (reference inverse_optimal_control.py:9-48)."""

    def __init__(self, car: PlannerCar, **kwargs):
        self.car = car
        self.world = self.car.env

    def rationalize(self, trajectory: List[Tuple]):
        raise NotImplementedError

    def rationalize_trajectories(self, trajectories: Collection[List[Tuple]]):
        raise NotImplementedError


def l2_normalize(w: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """tf.nn.l2_normalize: w / sqrt(max(sum w^2, eps))."""
    w = np.asarray(w, np.float64)
    return w / np.sqrt(max(float(np.sum(w * w)), eps))


def gradient_norm_loss(w: np.ndarray, blocks: Sequence[np.ndarray]) -> float:
    """sum over windows of || J_i^T w ||^2 (segment_loss, :62-91; compute_total_loss :93-113)."""
    return float(sum(np.sum((J.T @ w) ** 2) for J in blocks))


def adam_on_sphere(blocks: Sequence[np.ndarray], initial: np.ndarray, weight_norm: float, n_iter: int,
                   learning_rate: float = 0.1) -> np.ndarray:
    """The reference's optimisation loop (:115-138): Keras Adam (beta1 0.9, beta2 0.999, eps 1e-7) on the
    unnormalised weights of  L(theta) = sum_i || J_i^T (norm * theta / |theta|) ||^2.
    With M = sum_i J_i J_i^T the loss is w^T M w, so one K x K matrix replaces the tape."""
    K = blocks[0].shape[0]
    M = np.zeros((K, K))
    for J in blocks:
        M += J @ J.T
    theta = np.asarray(initial, np.float64).copy()
    m, v = np.zeros(K), np.zeros(K)
    b1, b2, eps = 0.9, 0.999, 1e-7
    for t in range(1, n_iter + 1):
        nrm = np.sqrt(max(float(theta @ theta), 1e-12))
        w = weight_norm * theta / nrm
        gw = 2.0 * (M @ w)                                   # dL/dw
        g = weight_norm * (gw - (theta @ gw) * theta / nrm ** 2) / nrm     # through the normalisation
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        lr_t = learning_rate * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta


def nullspace_weights(jacobian: np.ndarray) -> np.ndarray:
    """The unit w minimising || J^T w ||: the left singular vector of the smallest singular value
    (`_, u, _ = tf.linalg.svd(jacobian); u[:, -1]`, :284-287)."""
    u, _, _ = np.linalg.svd(np.asarray(jacobian, np.float64), full_matrices=True)
    return u[:, -1]


class InverseLocallyOptimalControl(InverseOptimalControl):
    """Minimises the squared norm of d r / d u over the (normalised) weights (reference :16-168)."""

    def __init__(self, car: PlannerCar, weight_norm: float = 1., initial_weights=None, **kwargs):
        super().__init__(car, **kwargs)
        if getattr(car, "planner", None) is None:
            car.initialize_planner(car.planner_args)
        if not isinstance(car.planner, NaivePlanner):
            raise NotImplementedError("Only NaivePlanners are supported.")
        K = len(np.asarray(car.weights))
        self.initial_weights = np.ones(K) if initial_weights is None else np.array(initial_weights, np.float64)
        self.unnorm_weights = self.initial_weights.copy()
        self.weight_norm = float(weight_norm)

    @property
    def weights(self) -> np.ndarray:
        return (self.weight_norm * l2_normalize(self.unnorm_weights)).astype(np.float32)

    @weights.setter
    def weights(self, new_value):
        self.unnorm_weights = np.array(new_value, np.float64)

    # -- the one device computation ---------------------------------------------------------------------
    def window_jacobians(self, trajectory: List[Tuple]) -> np.ndarray:
        """[n_windows, K, H, 2]: for every sliding window of `horizon` steps, the Jacobian of the summed
        features with respect to the window's controls, from its first world state (one launch)."""
        planner = self.car.planner
        H = planner.horizon
        n = len(trajectory) - H + 1
        if n < 1:
            raise ValueError("trajectory is shorter than the planning horizon")
        me = self.car.index
        states = np.stack([np.stack([as_f32(s, (4,)) for s in trajectory[i][0]]) for i in range(n)])
        controls = np.stack([np.stack([as_f32(trajectory[i + j][1][me], (2,)) for j in range(H)]) for i in range(n)])
        return planner.feature_jacobian_batch(states, controls)[1]

    def jacobian_blocks(self, trajectory: List[Tuple]) -> List[np.ndarray]:
        """Per window the columns the reference keeps: the first control only (to prevent double
        counting), every control for the last window (:100-112, :251-263)."""
        jac = self.window_jacobians(trajectory).astype(np.float64)
        n, K = jac.shape[0], jac.shape[1]
        return [jac[i, :, 0, :] if i < n - 1 else jac[i].reshape(K, -1) for i in range(n)]

    def segment_loss(self, weights, initial_state, controls, index: Optional[int] = None) -> float:
        """|| d r / d u ||^2 for one window (:62-91); index selects one control."""
        planner = self.car.planner
        jac = planner.feature_jacobian_batch(np.stack([as_f32(s, (4,)) for s in initial_state])[None],
                                             np.stack([as_f32(c, (2,)) for c in controls])[None])[1][0]
        J = jac.reshape(jac.shape[0], -1) if index is None else jac[:, index, :]
        return float(np.sum((J.astype(np.float64).T @ np.asarray(weights, np.float64)) ** 2))

    def compute_total_loss(self, weights, trajectory: List[Tuple]) -> float:
        return gradient_norm_loss(np.asarray(weights, np.float64), self.jacobian_blocks(trajectory))

    def rationalize(self, trajectory: List[Tuple], n_iter: int = 300) -> np.ndarray:
        self.unnorm_weights = adam_on_sphere(self.jacobian_blocks(trajectory), self.initial_weights,
                                             self.weight_norm, n_iter)
        return self.weights

    def rationalize_trajectories(self, trajectories: Collection[List[Tuple]], n_iter: int = 200) -> np.ndarray:
        blocks = [J for tr in trajectories for J in self.jacobian_blocks(tr)]
        self.unnorm_weights = adam_on_sphere(blocks, self.initial_weights, self.weight_norm, n_iter)
        return self.weights


class LinearInverseLocallyOptimalControl(InverseLocallyOptimalControl):
    """For linear rewards d r / d u = J^T w: take w from the SVD of J (reference :171-314)."""

    def __init__(self, car, **kwargs):
        if not isinstance(car, PlannerCar):
            raise ValueError("Car must also be a planner car.")
        super().__init__(car=car, **kwargs)

    def segment_jacobian(self, initial_state, controls, index: Optional[int] = None) -> List[np.ndarray]:
        """List of [K, 2] Jacobians, one per control (or only control `index`) (:193-240)."""
        planner = self.car.planner
        jac = planner.feature_jacobian_batch(np.stack([as_f32(s, (4,)) for s in initial_state])[None],
                                             np.stack([as_f32(c, (2,)) for c in controls])[None])[1][0]
        return [jac[:, t, :] for t in range(jac.shape[1])] if index is None else [jac[:, index, :]]

    def total_jacobian(self, trajectory: List[Tuple]) -> np.ndarray:
        """[K, 2 (n_windows - 1) + 2 H] (:242-268)."""
        return np.concatenate(self.jacobian_blocks(trajectory), axis=-1)

    def rationalize(self, trajectory: List[Tuple], **kwargs) -> np.ndarray:
        self.unnorm_weights = nullspace_weights(self.total_jacobian(trajectory))
        return self.weights

    def rationalize_trajectories(self, trajectories: Collection[List[Tuple]], **kwargs) -> np.ndarray:
        J = np.concatenate([self.total_jacobian(tr) for tr in trajectories], axis=-1)
        self.unnorm_weights = nullspace_weights(J)
        return self.weights
