"""Second-order inverse optimal control: mirror of LocalCIOC (interact_drive/reward_design/second_order_ioc.py:17-287
of the reference) -- Levine & Koltun (2012), "Continuous Inverse Optimal Control with Locally Optimal Examples", with
the augmented-Lagrangian treatment of the dummy feature theta_r of the paper's section 6.1.

The reference builds the trajectory's gradient vector g and the matrix H = d g / d u with nested TensorFlow tapes inside
every Adam step and differentiates the likelihood through both.  The reward is linear in the weights, r = w . phi, so
    g(w) = sum_k w_k g_k ,    H(w) = sum_k w_k H_k ,
with g_k, H_k independent of w.  Here they come from TWO batched launches over all sliding windows of the trajectory
(`ocd_feature_jacobian_batch`, `ocd_feature_hessian_batch`), once; the likelihood
    log L = 1/2 g^T A^-1 g + 1/2 sign log|det(-A)| - 1/2 mu theta_r^2 + lambda theta_r ,   A = H - theta_r I
(:149-163) and its derivatives in (w, theta_r) are then dense linear algebra on [2T x 2T] matrices on the host
(float64 torch autograd standing in for TensorFlow's), which is all that remains inside the Adam loop.

Row selection follows the reference exactly (:104-140): every window but the last contributes the gradient of its FIRST
control only (two rows), the last window all of its 2H rows -- or, with split_traj, disjoint windows contribute all
their rows.  A row's derivative is taken with respect to ALL controls of the trajectory, i.e. the window's Hessian rows
placed at the window's columns.

What differs from the reference, and why: as for the first-order classes (first_order_ioc.py) the reward features are
the kernels' own (the planning car must be a ThreeLaneTestCar); `segment_gradient` in the reference passes the weights
in the `other_controls` slot of today's `reward_func` (:59) -- the intended binding is used.  Parity for the class is
therefore UNPINNED against the reference; the Hessian operator underneath is checked against finite differences of the
CPU oracle's float64 gradient (tests/test_gpu_ioc.py).
"""
from __future__ import annotations

import logging
from typing import Collection, List, Optional, Tuple

import numpy as np
import torch

from ...runtime import as_f32
from .first_order_ioc import InverseLocallyOptimalControl

logger = logging.getLogger(__name__)


class LocalCIOC(InverseLocallyOptimalControl):
    """Maximum-likelihood weights of a Boltzmann-rational demonstrator under the Laplace approximation
    (reference :17-32)."""

    def __init__(self, *args, split_traj: bool = False, **kwargs):
        super().__init__(*args, **kwargs)
        self.theta_r = 0.01
        self.split_traj = bool(split_traj)

    # -- device part: per-feature gradient rows and Hessian rows of the whole trajectory ---------------------------
    def window_starts(self, n_steps: int) -> List[int]:
        H = self.car.planner.horizon
        if self.split_traj:
            return list(range(0, max(n_steps - H + 1, 1), H))
        return list(range(max(n_steps - H + 1, 1)))

    def trajectory_terms(self, trajectory: List[Tuple]) -> Tuple[np.ndarray, np.ndarray]:
        """-> (G [K, R], Hm [K, R, 2T]) in float64: per feature, the selected gradient rows of every window and
        their derivatives with respect to all 2T controls of the trajectory (:104-147).  Two launches."""
        planner = self.car.planner
        H, T, me = planner.horizon, len(trajectory), self.car.index
        if T < H:
            raise ValueError("trajectory is shorter than the planning horizon")
        starts = self.window_starts(T)
        states = np.stack([np.stack([as_f32(s, (4,)) for s in trajectory[i][0]]) for i in starts])
        controls = np.stack([np.stack([as_f32(trajectory[i + j][1][me], (2,)) for j in range(H)]) for i in starts])
        jac = planner.feature_jacobian_batch(states, controls)[1].astype(np.float64)      # [n, K, H, 2]
        hes = planner.feature_hessian_batch(states, controls).astype(np.float64)           # [n, K, 2H, 2H]
        K = jac.shape[1]
        g_rows, h_rows = [], []
        for w_i, i in enumerate(starts):
            last = self.split_traj or i >= T - H          # all rows; otherwise the first control only
            rows = range(2 * H) if last else range(2)
            for r in rows:
                g_rows.append(jac[w_i].reshape(K, 2 * H)[:, r])
                full = np.zeros((K, 2 * T))
                full[:, 2 * i:2 * i + 2 * H] = hes[w_i, :, r, :]
                h_rows.append(full)
        return np.stack(g_rows, axis=1), np.stack(h_rows, axis=1)

    # -- host part ---------------------------------------------------------------------------------------------------
    @staticmethod
    def augmented_loss(weights: torch.Tensor, theta_r: torch.Tensor, G: torch.Tensor, Hm: torch.Tensor, mu: float,
                       lm: float) -> Tuple[torch.Tensor, torch.Tensor]:
        """-(log-likelihood with augmented-Lagrangian terms), sign of det(-A)   (reference :149-165)."""
        g = (weights[:, None] * G).sum(0)[:, None]                      # [R, 1]
        A = (weights[:, None, None] * Hm).sum(0)                        # [R, 2T]  (R == 2T for a full trajectory)
        if A.shape[0] != A.shape[1]:
            raise ValueError("the selected gradient rows do not cover the trajectory's controls once each")
        A = A - theta_r * torch.eye(A.shape[0], dtype=A.dtype)
        sign, logabs = torch.linalg.slogdet(-A)
        log_ll = (0.5 * (g.T @ torch.linalg.solve(A, g)).squeeze() + 0.5 * sign.detach() * logabs
                  - 0.5 * mu * theta_r ** 2 + lm * theta_r)
        return -log_ll, sign.detach()

    def _normalised(self, unnorm: torch.Tensor) -> torch.Tensor:
        return self.weight_norm * unnorm / torch.sqrt(torch.clamp((unnorm * unnorm).sum(), min=1e-12))

    def compute_total_augmented_loss(self, weights, trajectory: List[Tuple], theta_r: float, mu: float,
                                     lm: float) -> Tuple[float, float]:
        G, Hm = (torch.as_tensor(a) for a in self.trajectory_terms(trajectory))
        loss, sign = self.augmented_loss(torch.as_tensor(np.asarray(weights, np.float64)),
                                         torch.tensor(float(theta_r), dtype=torch.float64), G, Hm, mu, lm)
        return float(loss), float(sign)

    def rationalize(self, trajectory: List[Tuple], initial_theta_r: float = 0.01, n_iter: int = 200,
                    initial_mu: float = 10.0, tol: float = 0.01, max_outer: int = 50) -> np.ndarray:
        """The reference's augmented-Lagrangian loop (:167-279): double theta_r until det(-A) > 0, Adam (lr 0.1) on
        (unnormalised weights, theta_r), then raise the multiplier / penalty until |theta_r| <= tol."""
        G, Hm = (torch.as_tensor(a) for a in self.trajectory_terms(trajectory))
        unnorm = torch.tensor(np.asarray(self.initial_weights, np.float64), requires_grad=True)
        theta = torch.tensor(float(initial_theta_r), dtype=torch.float64, requires_grad=True)
        mu, lm = float(initial_mu), 0.0
        with torch.no_grad():
            while self.augmented_loss(self._normalised(unnorm), theta, G, Hm, mu, lm)[1] < 0:
                theta.mul_(2.0)
                logger.info("doubling theta_r to %.3f", float(theta))
                if float(theta) > 1e12:
                    raise FloatingPointError("det(-H + theta_r I) stays negative")
        theta_val = float(theta.detach())
        opt = torch.optim.Adam([unnorm, theta], lr=0.1, betas=(0.9, 0.999), eps=1e-7)

        def train():
            for _ in range(n_iter):
                opt.zero_grad()
                loss, _ = self.augmented_loss(self._normalised(unnorm), theta, G, Hm, mu, lm)
                loss.backward()
                opt.step()

        train()
        outer = 0
        while abs(float(theta.detach())) > tol and outer < max_outer:
            lm = lm - mu * float(theta.detach())
            if abs(float(theta.detach())) - abs(theta_val) >= -5e-4:
                mu *= 10.0
                logger.info("theta_r did not decrease: penalty raised to %.2f", mu)
            theta_val = float(theta.detach())
            train()
            outer += 1
        with torch.no_grad():
            _, sign = self.augmented_loss(self._normalised(unnorm), theta, G, Hm, mu, lm)
        if sign < 0:
            logger.warning("Negative Hessian is ill-conditioned (|-H| < 0): results may be nonsensical; try a larger "
                           "tolerance")
        self.unnorm_weights = unnorm.detach().numpy().copy()
        self.theta_r = float(theta.detach())
        return self.weights

    def rationalize_trajectories(self, trajectories: Collection[List[Tuple]], **kwargs):
        raise NotImplementedError            # as in the reference (:539-543)
