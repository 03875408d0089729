"""The drop-in layer (interact_drive / experiments mirror) on the GPU: the reference's own tests for
the path, re-run against the engine through the reference's class interface, plus MPC_ORD against the
golden episodes."""
import pickle

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

import l4dc_mpc_ocd_b200 as ocd                                                     # noqa: E402
from l4dc_mpc_ocd_b200.experiments.local_opt_scenario import local_opt_env          # noqa: E402
from l4dc_mpc_ocd_b200.experiments.merging import ThreeLaneTestCar                  # noqa: E402
from l4dc_mpc_ocd_b200.experiments.replanning_world import setup_world              # noqa: E402
from l4dc_mpc_ocd_b200.experiments import run_mpc_ord                               # noqa: E402
from l4dc_mpc_ocd_b200.interact_drive import math_utils, simulation_utils           # noqa: E402
from l4dc_mpc_ocd_b200.interact_drive.car import FixedVelocityCar                   # noqa: E402
from l4dc_mpc_ocd_b200.interact_drive.planner import CarPlanner, NaivePlanner       # noqa: E402
from l4dc_mpc_ocd_b200.interact_drive.reward_design.mpc_ord import MPC_ORD, finite_horizon_env  # noqa: E402
from l4dc_mpc_ocd_b200.interact_drive.world import ThreeLaneCarWorld                # noqa: E402

ENVS = {"finite_horizon": finite_horizon_env, "local_opt": local_opt_env, "replanning": setup_world}


@pytest.fixture(autouse=True)
def _need_gpu(engine):
    return engine


class TargetSpeedPlannerCar(ThreeLaneTestCar):
    """The helper car of the reference's planner tests (interact_drive/planner/tests/
    targetSpeedRewardMaximizerCar.py:45-53): reward -(v - target)^2, i.e. feature 0 alone."""

    def __init__(self, env, init_state, horizon, target_speed=1., friction=0.2, **kw):
        super().__init__(env, init_state, horizon=horizon, weights=[-1., 0, 0, 0, 0, 0, 0],
                         target_speed=target_speed, friction=friction, **kw)


# ---- interact_drive/tests/test_simulation_utils.py:113-158 ------------------------------------------
def test_next_car_state_known_answers():
    hp = np.pi / 2
    for mu, exp in ((0.0, [0, 1, 1, hp]), (1.0, [0, 0.5, 0, hp]), (0.5, [0, 0.75, 0.5, hp])):
        out = simulation_utils.next_car_state([0., 0., 1., hp], [0., 0.], 1.0, mu)
        np.testing.assert_allclose(out, exp, atol=1e-6)
    np.testing.assert_allclose(simulation_utils.next_car_state([0., 0., 1., 0.], [0., 0.], 1.0, 0.5),
                               [0.75, 0, 0.5, 0], atol=1e-6)
    out = simulation_utils.batched_next_car_state([[0., 0., 1., hp], [0., 0., 1., 0.]], [[0., 0.]] * 2, 1.0, 0.5)
    np.testing.assert_allclose(out, [[0, 0.75, 0.5, hp], [0.75, 0, 0.5, 0]], atol=1e-6)
    f = simulation_utils.get_dynamics_fn(0.5)
    np.testing.assert_allclose(f([0., 0., 1., 0.], [0., 0.], 1.0), [0.75, 0, 0.5, 0], atol=1e-6)
    x, y, v, th = simulation_utils.car_dynamics_step(0., 0., 1., hp, 0., 0., 1.0, 0.5)
    np.testing.assert_allclose([x, y, v, th], [0, 0.75, 0.5, hp], atol=1e-6)


def test_next_car_state_shape_errors():
    with pytest.raises(ValueError):
        simulation_utils.next_car_state([0., 0., 1.], [0., 0.], 0.1)
    with pytest.raises(ValueError):
        simulation_utils.next_car_state([0., 0., 1., 0.], [0.], 0.1)
    with pytest.raises(ValueError):
        simulation_utils.batched_next_car_state(np.zeros((3, 5)), np.zeros((3, 2)), 0.1)


# ---- interact_drive/math_utils.py doctests ----------------------------------------------------------
def test_math_utils_doctests():
    g = load_golden("primitives.json")
    assert math_utils._f(0.) == 0 and math_utils._f(1.) > 0
    assert abs(math_utils._f(1e10) - 1) < 1e-6
    t = math_utils.smooth_threshold(0., 1.)
    assert t(0.) == 1 and t(-1.) == 0 and abs(t(-0.5) - 0.5) < 1e-6
    b = math_utils.smooth_bump(-1., 1.)
    assert b(0.) == 1 and b(-1.) == 0 and b(1.) == 0 and b(0.5) > 0
    for c in g["f"]:
        assert abs(math_utils._f(c["x"], c["shape"]) - c["y"]) <= 1e-6
    for c in g["threshold"]:
        assert abs(math_utils.smooth_threshold(c["threshold"], c["width"])(c["z"]) - c["y"]) <= 2e-6
    for c in g["bump"]:
        assert abs(math_utils.smooth_bump(c["start"], c["end"])(c["z"]) - c["y"]) <= 2e-6


# ---- interact_drive/planner/tests/test_naivePlanner.py ----------------------------------------------
def test_car_planner_is_abstract():
    with pytest.raises(NotImplementedError):
        CarPlanner(None, None).generate_plan()


def test_zero_friction_correct_speed():        # reference :21-32
    world = ThreeLaneCarWorld()
    car = TargetSpeedPlannerCar(world, np.array([0., 0., 1., np.pi / 2], np.float32), 4, target_speed=1., friction=0.)
    world.add_car(car)
    planner = NaivePlanner(world, car, horizon=5, learning_rate=5.0, n_iter=100)
    plan = planner.generate_plan([car.state])
    assert len(plan) == 5
    np.testing.assert_allclose(np.stack(plan), [[0., 0.]] * 5, atol=1e-5)


def test_friction_correct_speed():             # reference :50-63
    world = ThreeLaneCarWorld()
    car = TargetSpeedPlannerCar(world, np.array([0., 0., 1., np.pi / 2], np.float32), 4, target_speed=1., friction=0.5)
    world.add_car(car)
    planner = NaivePlanner(world, car, horizon=3, learning_rate=5.0, n_iter=500)
    plan = planner.generate_plan([car.state])
    np.testing.assert_allclose(np.stack(plan), [[0.5, 0.]] * 3, atol=5e-5)
    kat = load_golden("planner_kats.json")["cases"][1]
    np.testing.assert_allclose(np.stack(plan), kat["plan"], atol=5e-5)


def test_use_lbfgs_reaches_the_known_answers():
    """generate_plan(use_lbfgs=True): the engine's L-BFGS on the same two known-answer problems."""
    world = ThreeLaneCarWorld()
    car = TargetSpeedPlannerCar(world, np.array([0., 0., 1., np.pi / 2], np.float32), 4, target_speed=1., friction=0.5)
    world.add_car(car)
    planner = NaivePlanner(world, car, horizon=3, learning_rate=5.0, n_iter=500)
    plan = planner.generate_plan([car.state], use_lbfgs=True)
    np.testing.assert_allclose(np.stack(plan)[:, 0], [0.5] * 3, atol=1e-4)
    assert float(planner.last_losses[planner.last_best]) <= 1e-9


def test_no_interaction():                     # reference :76-96 (other_controls path; plus a value check)
    world = ThreeLaneCarWorld()
    car = TargetSpeedPlannerCar(world, np.array([0., 0., 1., np.pi / 2], np.float32), 4, friction=0.)
    other = FixedVelocityCar(world, np.array([0.1, 0.5, 1., np.pi / 2], np.float32))
    world.add_cars([car, other])
    planner = NaivePlanner(world, car, horizon=5, learning_rate=5.0, n_iter=100)
    oc = [np.zeros((5, 2), np.float32), np.tile(np.array([0.2, 0.1], np.float32), (5, 1))]
    plan = planner.generate_plan(other_controls=oc)
    np.testing.assert_allclose(np.stack(plan), [[0., 0.]] * 5, atol=1e-5)
    r = planner.reward_func(world.state, plan, other_controls=oc)
    assert abs(float(r)) < 1e-8
    r2, g = planner.mpc_reward_and_grad(world.state, [np.array([1., 0.], np.float32)] * 5)
    assert r2 < 0 and g.shape == (5, 2) and g[0, 0] < 0


def test_planner_rejects_unsupported():
    world = ThreeLaneCarWorld()
    car = TargetSpeedPlannerCar(world, np.array([0., 0., 1., np.pi / 2], np.float32), 4)
    world.add_car(car)
    with pytest.raises(NotImplementedError):
        NaivePlanner(world, car, 17).generate_plan(use_lbfgs=True)       # L-BFGS kernel: horizon <= 16
    with pytest.raises(NotImplementedError):
        NaivePlanner(world, car, 5, leaf_evaluation=lambda s, u: 0)
    with pytest.raises(TypeError):
        NaivePlanner(world, FixedVelocityCar(world, [0, 0, 1, 0]), 5).generate_plan()


# ---- features and plans of the shipped scenarios ------------------------------------------------------
def test_features_through_car():
    g = load_golden("features.json")
    c = g["cases"][0]
    world = ThreeLaneCarWorld()
    st = np.asarray(c["state"], np.float32)
    car = ThreeLaneTestCar(world, st[0], horizon=5, weights=np.ones(7), target_speed=c["target_speed"])
    world.add_cars([car] + [FixedVelocityCar(world, s) for s in st[1:]])
    np.testing.assert_allclose(car.features(list(st), None), c["phi"], rtol=2e-5, atol=2e-6)
    w = np.arange(1, 8, dtype=np.float32)
    assert abs(car.reward_fn(list(st), None, weights=w) - float(np.dot(w, c["phi"]))) < 1e-4


@pytest.mark.parametrize("name", list(ENVS))
def test_scenario_first_plan_matches_golden(name):
    g = load_golden("plans.json")
    car, world, inits = ENVS[name](env_seeds=[1000000])
    for c in [c for c in g["cases"] if c["scenario"] == name]:
        car.weights = np.asarray(c["weights_in"])
        np.testing.assert_allclose(car.weights, c["weights_normalised"], atol=1e-7)
        car.init_state = np.asarray(c["init"], np.float32)
        world.reset()
        ctrl = car._get_next_control()
        np.testing.assert_allclose(np.stack(car.plan), c["plan"], atol=1e-3)
        np.testing.assert_allclose(ctrl, c["control"], atol=1e-3)


# ---- CarWorld.step: the reference's serial loop, object by object ---------------------------------------
@pytest.mark.parametrize("fname", ["episode_finite_horizon_true_full.json", "episode_replanning_true_full.json"])
def test_world_step_loop_matches_golden(fname):
    e = load_golden(fname)
    car, world, _ = ENVS[e["scenario"]](env_seeds=[1000000])
    true_w = np.asarray(e["true_weights"], np.float32)
    car.weights = np.asarray(e["weights_in"])
    car.init_state = np.asarray(e["init"], np.float32)
    for smp in e["samples"]:
        world.reset()
        if hasattr(world, "unlucky_car_idx"):
            assert world.unlucky_car_idx == smp["unlucky_car_idx"]
        total = 0.0
        for i in range(e["T"]):
            past, controls, state = world.step()
            total += float(car.reward_fn(past, controls[car.index], weights=true_w))
            np.testing.assert_allclose(controls[0], smp["controls"][i], atol=1e-3)
        assert abs(total - smp["return"]) <= 1e-3 * abs(smp["return"])
        np.testing.assert_allclose(np.stack([s for s, _ in car.past_traj]), smp["robot_states"], atol=1e-3)


# ---- MPC_ORD ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fname", ["episode_finite_horizon_true_full.json", "episode_finite_horizon_tuned_full.json",
                                   "episode_finite_horizon_true_h6.json", "episode_local_opt_true_extra_inits.json",
                                   "episode_local_opt_scaled_short.json", "episode_replanning_true_full.json",
                                   "episode_replanning_tuned_full.json"])
def test_mpc_ord_matches_golden(fname):
    e = load_golden(fname)
    kw = {}
    if e["variant"] == "extra_inits":
        kw["extra_inits"] = True
    if e["variant"] == "h6":
        kw["horizon"] = 6
    car, world, _ = ENVS[e["scenario"]](env_seeds=[1000000], **kw)
    init = np.asarray(e["init"])
    nsamp = len(e["samples"])
    m = MPC_ORD(world, car, [init], e["T"], num_samples=nsamp, verbose=False)
    np.testing.assert_allclose(m.designer_weights, e["true_weights"], atol=1e-7)
    want = sum(s["return"] for s in e["samples"])
    got = m.eval_weights_for_init(init, np.asarray(e["weights_in"], np.float64), False)
    assert abs(got - want) <= 1e-3 * abs(want), (got, want)
    np.testing.assert_allclose(car.weights, e["plan_weights"], atol=1e-7)
    # debug cars keep the trajectory of the last sample, like the reference's past_traj
    last = e["samples"][-1]
    np.testing.assert_allclose(np.stack([c for _, c in car.past_traj]), last["controls"], atol=1e-3)
    # eval_weights = minus the per-sample average, and it is appended to the history
    neg = m.eval_weights(np.asarray(e["weights_in"], np.float64))
    assert abs(-neg - want / nsamp) <= 1e-3 * abs(want / nsamp)
    assert len(m.history) == 1 and m.iter == 1
    assert abs(m.history[0][1] - want / nsamp) <= 1e-3 * abs(want / nsamp)


def test_eval_weights_batch_equals_serial():
    car, world, inits = finite_horizon_env(env_seeds=[1000000, 1000001, 1000002])
    m = MPC_ORD(world, car, inits, 15, verbose=False)
    rng = np.random.default_rng(0)
    cands = [m.designer_weights + 0.05 * rng.normal(size=7) for _ in range(9)]
    batch = m.eval_weights_batch(cands)
    assert m.kernel_launches == 1 and len(m.history) == 9
    serial = np.array([m.eval_weights(c) for c in cands])
    np.testing.assert_array_equal(batch, serial)            # same kernel, same problems: bit-identical


def test_cmaes_run_and_history_pickle(tmp_path):
    car, world, inits = finite_horizon_env(env_seeds=[1000000, 1000001])
    path = tmp_path / "hist.pkl"
    m = MPC_ORD(world, car, inits, 15, save_path=str(path), verbose=False)
    x = m.optimize_cmaes(seed=7, sigma0=0.05, maxfevals=27)
    assert m.done and len(x) == 7
    assert len(m.history) == 1 + 27 and m.kernel_launches == 1 + 3     # designer + 3 generations of 9
    best = max(m.history, key=lambda a: a[1])
    assert best[1] >= m.history[0][1]                                     # never worse than the designer weights
    ocd.install_as_reference()
    with open(path, "rb") as f:
        hist = pickle.load(f)
    assert hist.seed == 7 and len(hist) == 28 and hist[0][0].shape == (7,)


def test_run_mpc_ord_cli_vis(capsys):
    out = run_mpc_ord.main(["finite_horizon", "vis", "--n_inits", "3", "--seed", "1", "--quiet", "--no_save"])
    m = out[0]
    assert len(m.history) == 2 and len(m.init_car_states) == 3
    # true weights on 3 inits, then the tuned weights of run_mpc_ord.py:34-35
    assert all(np.isfinite(h[1]) for h in m.history)
    assert "return of the tuned weights" in capsys.readouterr().out


# ---- "next" rows: reward heat-map and generalisation evaluation -------------------------------------
def test_reward_heatmap_matches_oracle_features():
    import oracle as O
    from l4dc_mpc_ocd_b200.heatmap import reward_heatmap
    car, world, inits = finite_horizon_env(env_seeds=[1000000])
    world.reset()
    lo, hi = (-0.15, -1.4), (0.15, -0.4)
    img = reward_heatmap(car, lo, hi, size=(24, 16))
    assert img.shape == (16, 24)
    xs = np.linspace(lo[0] + 1e-6, hi[0] - 1e-6, 24)
    ys = np.linspace(lo[1] + 1e-6, hi[1] - 1e-6, 16)
    p = O.OracleParams()
    for (j, i) in ((0, 0), (5, 11), (9, 12), (15, 23), (10, 2)):
        st = np.stack([c.state for c in world.cars]).astype(np.float32)
        st[0, 0], st[0, 1] = xs[i], ys[j]
        ref = float(np.dot(car.weights, O.features(p, st)))
        assert abs(img[j, i] - ref) <= 2e-6 * max(1.0, abs(ref))
    # the collision bump of the other car (at x=0, y=-0.6) must show up as the minimum of the map
    j, i = np.unravel_index(np.argmin(img[:, 8:16]), img[:, 8:16].shape)
    assert abs(ys[j] - (-0.6)) < 0.1


def test_generalization_evaluation_is_one_launch():
    import oracle as O
    from l4dc_mpc_ocd_b200.experiments import generalization_data as gd
    env = run_mpc_ord.envs["local_opt"]
    weights = {(1, 2): env["tuned_weights"], (3, 2): np.array([-5, 0., 0., -10, 0, -50, -50])}
    res = gd.evaluate_on_test_inits("local_opt", weights, n_test=4)
    assert sorted(res) == [0, 1, 2, 3] and set(res[0]) == set(weights)
    # one of them, the serial way
    car, world, test_inits = gd.make_test_env("local_opt", 4)
    m = MPC_ORD(world, car, [], env["eval_horizon"], num_samples=env["num_eval_samples"], verbose=False)
    one = m.eval_weights_for_init(test_inits[2], np.asarray(weights[(1, 2)], np.float64), False)
    assert abs(one - res[2][(1, 2)]) <= 1e-6 * abs(one)
    # ... and every (weights, held-out state) return against the ORACLE's serial episodes (1e-3 relative, BASELINE)
    spec = O.scenario_params("local_opt")
    T, ns = env["eval_horizon"], env["num_eval_samples"]
    w_true = np.asarray(m.designer_weights, np.float32)
    for key, wv in weights.items():
        wp = MPC_ORD._planning_weights(np.asarray(wv, np.float64))
        ri = np.repeat(np.asarray(test_inits[:4], np.float32), ns, axis=0)
        ref = O.episode_batch(spec.params, spec.scenario, ri, np.tile(wp, (ri.shape[0], 1)), w_true, T).reshape(4, ns).sum(1)
        got = np.array([res[i][key] for i in range(4)])
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-6)) <= 1e-3, (key, got, ref)
    hist = [(np.ones(7), -3.0), (np.arange(7.0), -1.0), (np.zeros(7), -2.0)]
    np.testing.assert_array_equal(gd.best_weights(hist), np.arange(7.0))
    np.testing.assert_array_equal(gd.best_weights(hist, num_evals=1), np.ones(7))
