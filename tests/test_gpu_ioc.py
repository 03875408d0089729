"""Feature Jacobian operator (ocd_feature_jacobian_batch) and the first-order IOC drop-ins built on it.

Row i of the Jacobian is the gradient of the horizon-summed feature i, i.e. the oracle's mpc_reward
gradient with weights e_i; J^T w must equal the reward gradient for any w (linearity, checked at full
size).  The IOC classes have no running reference counterpart (the reference's segment_loss mis-binds its
arguments and its tests use a car with other features), so they are checked by their own properties."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

import l4dc_mpc_ocd_b200 as ocd                                                    # noqa: E402
from l4dc_mpc_ocd_b200 import synthetic                                           # noqa: E402
from l4dc_mpc_ocd_b200.experiments.merging import ThreeLaneCarWorld, ThreeLaneTestCar   # noqa: E402
from l4dc_mpc_ocd_b200.interact_drive.car import FixedVelocityCar                 # noqa: E402
from l4dc_mpc_ocd_b200.interact_drive.reward_design import (                      # noqa: E402
    InverseLocallyOptimalControl, LinearInverseLocallyOptimalControl)
from l4dc_mpc_ocd_b200.interact_drive.reward_design.first_order_ioc import gradient_norm_loss, l2_normalize  # noqa: E402

MODES = [(ocd.MATH_FAST, 3e-4), (ocd.MATH_PRECISE, 2e-5)]


def _inputs(B, H, C, lanes, seed):
    lane_x = (-0.1, 0.0, 0.1) if lanes == 3 else (-0.05, 0.05)
    batch = synthetic.make_batch(B, C=C, lane_x=lane_x, seed=seed)
    rng = np.random.default_rng(seed + 1)
    world = batch["world"].copy()
    world[:, 0, 3] += rng.uniform(-0.3, 0.3, B).astype(np.float32)
    world[:, 0, 0] += rng.uniform(-0.06, 0.06, B).astype(np.float32)
    u = (rng.normal(size=(B, H, 2)) * np.array([1.0, 1.5])).astype(np.float32)
    return lane_x, batch, world, u


@pytest.mark.parametrize("mode,tol", MODES)
@pytest.mark.parametrize("H,C,lanes,other_mode", [(5, 2, 3, 0), (5, 3, 2, 1), (8, 4, 3, 0)])
def test_feature_jacobian_vs_oracle(engine, mode, tol, H, C, lanes, other_mode):
    B = 96
    lane_x, batch, world, u = _inputs(B, H, C, lanes, 31 + H + C)
    oc = synthetic.make_other_controls(B, C, H) if other_mode else None
    p = ocd.PlannerParams(H=H, C=C, lane_x=lane_x, num_lanes=lanes, other_mode=other_mode, math_mode=mode,
                          target_speed=1.0 if lanes == 3 else 1.2)
    phi, jac = engine.feature_jacobian(p, world, u, other_controls=oc)
    phi, jac = phi.cpu().numpy(), jac.cpu().numpy()
    assert phi.shape == (B, p.K) and jac.shape == (B, p.K, H, 2)
    op = O.OracleParams(H=H, C=C, lane_x=lane_x, num_lanes=lanes, other_mode=other_mode, target_speed=p.target_speed)
    eye = np.eye(p.K)
    worst_p = worst_j = 0.0
    for b in range(0, B, 3):
        ocb = None if oc is None else oc[b].astype(np.float64)
        for i in range(p.K):
            Rr, Gr = O.mpc_reward(op, world[b].astype(np.float64), u[b].astype(np.float64), eye[i], other_controls=ocb,
                                  dtype=np.float64)
            R32, G32 = O.mpc_reward(op, world[b], u[b], eye[i].astype(np.float32),
                                    other_controls=None if oc is None else oc[b])
            gs = max(1.0, np.abs(Gr).max())
            worst_p = max(worst_p, abs(phi[b, i] - Rr) / max(1.0, abs(Rr)) - 3 * abs(R32 - Rr) / max(1.0, abs(Rr)))
            worst_j = max(worst_j, np.max(np.abs(jac[b, i] - Gr)) / gs - 3 * np.max(np.abs(G32 - Gr)) / gs)
    assert worst_p <= tol and worst_j <= tol, (worst_p, worst_j)


def test_jacobian_is_linear_in_the_weights_full_size(engine):
    """J^T w == d(w . phi_sum)/du and w . phi_sum == reward for any w: 65 536 problems, precise math."""
    B, H = 65536, 5
    lane_x, batch, world, u = _inputs(B, H, 2, 3, 77)
    p = ocd.PlannerParams(math_mode=ocd.MATH_PRECISE)
    phi, jac = engine.feature_jacobian(p, world, u)
    R, G = engine.reward(p, world, u, batch["weights"], weight_idx=batch["weight_idx"])
    w = batch["weights"][batch["weight_idx"]]                                        # [B, K]
    wt = np.asarray(w, np.float64)
    Rj = np.einsum("bk,bk->b", wt, phi.cpu().numpy().astype(np.float64))
    Gj = np.einsum("bk,bkhc->bhc", wt, jac.cpu().numpy().astype(np.float64))
    R, G = R.cpu().numpy(), G.cpu().numpy()
    assert np.max(np.abs(Rj - R) / np.maximum(1.0, np.abs(R))) <= 2e-5
    # the min / max features switch branch with the weights' sign nowhere: rows are weight-independent
    assert np.max(np.abs(Gj - G) / np.maximum(1.0, np.abs(G).max(axis=(1, 2), keepdims=True))) <= 5e-5


def _drive(T=9):
    """finite_horizon-like world: the planning car and one constant-velocity car; T recorded steps."""
    world = ThreeLaneCarWorld()
    w_true = np.array([-5., 0., 0., 0., -6., -50., -50.])
    car = ThreeLaneTestCar(world, np.array([0.02, -0.9, 0.8, np.pi / 2], np.float32), horizon=5, weights=w_true,
                           planner_args=dict(n_iter=100, learning_rate=0.1))
    other = FixedVelocityCar(world, np.array([0.0, -0.6, 0.5, np.pi / 2], np.float32))
    world.add_cars([car, other])
    traj = []
    for _ in range(T):
        past, controls, _ = world.step()
        traj.append((past, controls))
    return world, car, traj, l2_normalize(w_true)


def test_ioc_drop_ins():
    world, car, traj, w_true = _drive()
    lin = LinearInverseLocallyOptimalControl(car, weight_norm=1.)
    J = lin.total_jacobian(traj)
    n = len(traj) - 5 + 1
    assert J.shape == (7, 2 * (n - 1) + 2 * 5)
    w_svd = lin.rationalize(traj)
    assert abs(np.linalg.norm(w_svd) - 1) < 1e-6
    blocks = lin.jacobian_blocks(traj)
    loss_svd = gradient_norm_loss(np.asarray(w_svd, np.float64), blocks)
    rng = np.random.default_rng(0)
    for _ in range(50):        # the SVD direction minimises || J^T w || over the unit sphere
        assert loss_svd <= gradient_norm_loss(l2_normalize(rng.normal(size=7)), blocks) + 1e-12
    # an MPC trajectory (100 SGD steps per solve: not converged) is closer to first-order optimal for the
    # weights that produced it than for a typical direction
    loss_true = lin.compute_total_loss(w_true, traj)
    assert loss_svd <= loss_true <= 0.5 * np.median([gradient_norm_loss(l2_normalize(rng.normal(size=7)), blocks)
                                          for _ in range(200)])
    iloc = InverseLocallyOptimalControl(car, weight_norm=1., initial_weights=-np.ones(7))
    l0 = iloc.compute_total_loss(iloc.weights, traj)
    w_adam = iloc.rationalize(traj, n_iter=300)
    assert iloc.compute_total_loss(w_adam, traj) < 0.1 * l0
    # segment_loss of one window == the block of compute_total_loss
    init, ctr = traj[0][0], [traj[j][1][car.index] for j in range(5)]
    assert abs(iloc.segment_loss(w_true, init, ctr, index=0) - np.sum((blocks[0].T @ w_true) ** 2)) < 1e-9
    with pytest.raises(ValueError):
        lin.total_jacobian(traj[:3])


# ---- second order: feature Hessians (ocd_feature_hessian_batch) and LocalCIOC --------------------------------------
@pytest.mark.parametrize("H,C,lanes,other_mode", [(5, 2, 3, 0), (5, 3, 2, 1), (7, 4, 3, 0)])
def test_feature_hessian_vs_oracle_differences(engine, H, C, lanes, other_mode):
    """hess[k] = d^2 (sum_t phi_k) / du du: compared with central differences of the oracle's float64 GRADIENT (weights
    e_k), an independent route to the same matrix.  Symmetric by construction of the kernel; the comparison covers both
    triangles."""
    from l4dc_mpc_ocd_b200.interact_drive.reward_design import LocalCIOC  # noqa: F401  (importable on a GPU box)
    B = 24
    lane_x, batch, world, u = _inputs(B, H, C, lanes, 91 + H + C)
    u = (0.5 * u).astype(np.float32)
    oc = 0.3 * synthetic.make_other_controls(B, C, H) if other_mode else None
    p = ocd.PlannerParams(H=H, C=C, lane_x=lane_x, num_lanes=lanes, other_mode=other_mode,
                          target_speed=1.0 if lanes == 3 else 1.2)
    hes = engine.feature_hessian(p, world, u, other_controls=oc).cpu().numpy()
    n = 2 * H
    assert hes.shape == (B, p.K, n, n)
    assert np.array_equal(hes, hes.transpose(0, 1, 3, 2))
    op = O.OracleParams(H=H, C=C, lane_x=lane_x, num_lanes=lanes, other_mode=other_mode, target_speed=p.target_speed)
    eye = np.eye(p.K)
    h = 1e-5
    worst, checked, skipped = 0.0, 0, 0
    for b in range(0, B, 2):
        ocb = None if oc is None else oc[b].astype(np.float64)
        w64, u64 = world[b].astype(np.float64), u[b].astype(np.float64)
        for k_ in range(p.K):
            ref = np.zeros((n, n))
            for i in range(n):
                e = np.zeros(n); e[i] = h
                gp = O.mpc_reward(op, w64, u64 + e.reshape(H, 2), eye[k_], other_controls=ocb, dtype=np.float64)[1]
                gm = O.mpc_reward(op, w64, u64 - e.reshape(H, 2), eye[k_], other_controls=ocb, dtype=np.float64)[1]
                ref[i] = (np.asarray(gp) - np.asarray(gm)).reshape(n) / (2 * h)
            # a kink of min / max / clip inside the differencing interval makes the difference quotient meaningless
            # (the one-sided quotients disagree): skip those matrices, they are rare
            g0 = np.asarray(O.mpc_reward(op, w64, u64, eye[k_], other_controls=ocb, dtype=np.float64)[1]).reshape(n)
            asym = np.abs(ref - ref.T).max()
            scale = max(1.0, np.abs(ref).max())
            if asym > 1e-4 * scale:
                skipped += 1
                continue
            worst = max(worst, np.abs(hes[b, k_] - ref).max() / scale)
            checked += 1
            assert np.all(np.isfinite(g0))
    assert checked >= 0.85 * (checked + skipped), (checked, skipped)
    assert worst <= 2e-3, worst


def test_hessian_of_quadratic_features_is_exact(engine):
    """Straight driving with zero steering: the lane features 10 (x - l)^2 do not depend on the accelerations at all
    and the whole matrix of the far-away collision feature is zero -- exact zeros, not small numbers."""
    H = 5
    world = np.array([[[0.02, -0.9, 0.8, np.pi / 2], [0.1, 5.0, 0.5, np.pi / 2]]], np.float32)
    u = np.zeros((1, H, 2), np.float32)
    u[0, :, 0] = 0.3
    hes = engine.feature_hessian(ocd.PlannerParams(), world, u).cpu().numpy()[0]
    acc = np.arange(0, 2 * H, 2)
    assert np.all(hes[5] == 0.0)                                   # collision: the other car is far outside the support
    # the steering block of the speed feature (v sin th - ts)^2 is not zero
    assert np.abs(hes[0][1::2, 1::2]).max() > 0.0
    assert np.all(np.isfinite(hes))
    assert np.abs(hes[1][np.ix_(acc, acc)]).max() < 1e-6            # lanes: heading exactly along y, acc moves y only


def test_local_cioc_runs_on_a_planned_trajectory():
    from l4dc_mpc_ocd_b200.interact_drive.reward_design import LocalCIOC
    world, car, traj, w_true = _drive(T=8)
    cioc = LocalCIOC(car, weight_norm=1., initial_weights=-np.ones(7))
    G, Hm = cioc.trajectory_terms(traj)
    T = len(traj)
    assert G.shape == (7, 2 * T) and Hm.shape == (7, 2 * T, 2 * T)
    # the rows are the first-order classes' columns, in the same order
    blocks = np.concatenate(cioc.jacobian_blocks(traj), axis=-1)
    np.testing.assert_allclose(G, blocks, rtol=0, atol=1e-12)
    l0, s0 = cioc.compute_total_augmented_loss(cioc.weights, traj, theta_r=100.0, mu=10.0, lm=0.0)
    assert np.isfinite(l0) and s0 > 0                               # a large theta_r makes -A positive definite
    w = cioc.rationalize(traj, n_iter=60, tol=0.05, max_outer=6)
    assert w.shape == (7,) and np.all(np.isfinite(w)) and abs(np.linalg.norm(w) - 1.0) < 1e-5
    split = LocalCIOC(car, weight_norm=1., split_traj=True)
    Gs, Hs = split.trajectory_terms(traj[:5])
    assert Gs.shape == (7, 10) and Hs.shape == (7, 10, 10)
