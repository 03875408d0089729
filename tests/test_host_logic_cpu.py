"""Host-side logic that needs no GPU: scenario constructors, the world compiler, weight
normalisation, CMA-ES, sharding arithmetic, the synthetic generator and bench.py's reference arm."""
import json
import pickle
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import l4dc_mpc_ocd_b200 as ocd
from l4dc_mpc_ocd_b200 import cmaes, parallel, synthetic
from l4dc_mpc_ocd_b200.batched import compile_world, unlucky_sequence
from l4dc_mpc_ocd_b200.experiments.local_opt_scenario import local_opt_env
from l4dc_mpc_ocd_b200.experiments.replanning_world import setup_world, og_weights
from l4dc_mpc_ocd_b200.experiments import run_mpc_ord
from l4dc_mpc_ocd_b200.interact_drive.car import FixedPlanCar, FixedVelocityCar, PlannerCar
from l4dc_mpc_ocd_b200.interact_drive.reward_design.mpc_ord import MPC_ORD, finite_horizon_env, list2
from l4dc_mpc_ocd_b200.interact_drive.world import StraightLane, ThreeLaneCarWorld, TwoLaneCarWorld
from conftest import load_golden

ROOT = Path(__file__).resolve().parent.parent


def test_lane_geometry():
    assert ThreeLaneCarWorld().lane_medians() == pytest.approx((-0.1, 0.0, 0.1))
    assert TwoLaneCarWorld().lane_medians() == pytest.approx((-0.05, 0.05))
    lane = StraightLane((0.0, -5.), (0.0, 10.), 0.1)
    assert lane.dist2median((0.03, 7.0)) == pytest.approx(0.03 ** 2)
    assert lane.shifted(-1).dist2median((0.1, 0.0)) == pytest.approx(0.0)
    with pytest.raises(ValueError):
        StraightLane((0, 0), (1, 1), 0.1).median_x


def test_scenario_constructors_and_compiler():
    car, world, inits = finite_horizon_env(env_seeds=[1000000, 1000001])
    assert len(inits) == 2 and inits[0][3] == pytest.approx(np.pi / 2)
    assert -0.1 <= inits[0][0] <= 0.1 and -0.95 <= inits[0][1] <= -0.85 and 0.7 <= inits[0][2] <= 0.9
    np.testing.assert_allclose(car.weights, np.array([-5, 0, 0, 0, -6, -50, -50]) / np.linalg.norm([-5, 0, 0, 0, -6, -50, -50]),
                               atol=1e-7)
    prog = compile_world(world, car)
    assert (prog.params.H, prog.params.C, prog.params.n_iter, prog.params.other_mode) == (5, 2, 100, 0)
    assert prog.scenario.kind == [0] and prog.scenario.friction == [0.0] and not prog.replanning
    car6, world6, _ = finite_horizon_env(horizon=6, extra_inits=True)
    p6 = compile_world(world6, car6).params
    assert (p6.H, p6.n_iter, p6.extra_inits, p6.S) == (6, 200, True, 6)

    car, world, inits = local_opt_env(env_seeds=[5])
    assert -0.12 <= inits[0][0] <= -0.08 and 0.9 <= inits[0][2] <= 1.1
    assert compile_world(world, car).scenario.init_state[0][1] == pytest.approx(-0.9)

    car, world, inits = setup_world(env_seeds=[5])
    prog = compile_world(world, car)
    assert prog.replanning and prog.scenario.critical_t == 4 and prog.scenario.kind == [1, 1]
    assert prog.scenario.friction == [0.2, 0.2]               # FixedPlanCar keeps the Car default
    assert (prog.params.C, prog.params.L, prog.params.num_lanes, prog.params.other_mode) == (3, 2, 2, 1)
    assert prog.params.target_speed == pytest.approx(1.2)
    np.testing.assert_allclose(prog.scenario.plan[0][1], [0.7, 2.7])
    np.testing.assert_allclose(prog.scenario.plan[1][3], [0.0, 2.7])
    np.testing.assert_allclose(car.weights, og_weights, atol=1e-7)
    # ctor reset already toggled once (reference replanning_world.py:93): the next resets give 1, 2, 1
    assert world.unlucky_car_idx == 2
    assert unlucky_sequence(world, 3) == [1, 2, 1] and world.unlucky_car_idx == 1


def test_known_other_plans_replay_from_index_zero():
    car, world, _ = setup_world(env_seeds=[5])
    world.cars[1].t = 3            # the other car's own clock must not matter (quirk Q4)
    plans = car.known_other_plans()
    assert len(plans) == 3 and plans[0].shape == (5, 2) and not plans[0].any()
    np.testing.assert_allclose(plans[1], [[0, 0], [0.7, 2.7], [0, 0], [0, -2.7], [0, 0]])
    np.testing.assert_allclose(plans[2], [[0, 0], [0.7, -2.7], [0, 0], [0, 2.7], [0, 0]])


def test_compile_world_rejects_what_the_kernel_cannot_run():
    world = ThreeLaneCarWorld()
    a = PlannerCar(world, [0, 0, 1, 0], horizon=5)
    world.add_car(a)
    with pytest.raises(TypeError):
        compile_world(world, a)
    car, world, _ = finite_horizon_env()
    world.add_car(PlannerCar(world, [0, 0, 1, 0], horizon=5))
    with pytest.raises(TypeError):
        compile_world(world, car)
    car, world, _ = finite_horizon_env()
    car.planner_args["leaf_evaluation"] = lambda s, u: 0
    with pytest.raises(TypeError):
        compile_world(world, car)


def test_triple_normalisation_matches_reference():
    e = load_golden("episode_local_opt_scaled_short.json")
    w = MPC_ORD._planning_weights(np.asarray(e["weights_in"]))
    np.testing.assert_allclose(w, np.asarray(e["plan_weights"], np.float32), atol=1e-7)
    np.testing.assert_allclose(MPC_ORD._planning_weights(np.asarray(e["weights_in"])[None]), w)


def test_linear_reward_car_weight_setter():
    car, world, _ = finite_horizon_env()
    car.weights = [0, 3, 0, 4, 0, 0, 0]
    np.testing.assert_allclose(car.weights, [0, 0.6, 0, 0.8, 0, 0, 0], atol=1e-7)
    assert car.weights.dtype == np.float32


def test_fixed_plan_car_schedule_without_dynamics():
    world = TwoLaneCarWorld()
    c = FixedPlanCar(world, [0, 0, 1, 0], plan=[[1, 0], [2, 0]], default_control=[9, 9])
    c.reset()
    np.testing.assert_allclose(c.control, [1, 0])
    assert c.control_already_determined_for_current_step and c.friction == 0.2
    v = FixedVelocityCar(world, [0, 0, 1, 0])
    assert v.friction == 0.0 and not v.control.any()


def test_cmaes_minimises():
    f = lambda x: float(np.sum((np.asarray(x) - np.arange(7) / 10.0) ** 2))
    x, es = cmaes.fmin2(f, [0.0] * 7, 0.3, {"seed": 3, "maxfevals": 3000})
    assert es.lam == 9 and es.mu == 4                      # pycma defaults for N = 7 (4 + floor(3 ln 7))
    assert f(x) < 1e-8
    calls = []
    xb, esb = cmaes.fmin2(None, [0.0] * 7, 0.3, {"seed": 3, "maxfevals": 3000},
                          batch_objective=lambda P: (calls.append(len(P)), [f(p) for p in P])[1])
    np.testing.assert_array_equal(x, xb)                    # same seed, same path, one call per generation
    assert set(calls) == {9}
    ros = lambda x: float(sum(100 * (x[i + 1] - x[i] ** 2) ** 2 + (1 - x[i]) ** 2 for i in range(len(x) - 1)))
    xr, _ = cmaes.fmin2(ros, [0.0] * 4, 0.5, {"seed": 1, "maxfevals": 20000})
    assert ros(xr) < 1e-6


def test_history_pickles_under_the_reference_path(tmp_path):
    h = list2()
    h.seed = 11
    h.append((np.ones(3), -1.5))
    class Holder:
        save_path, history, verbose = str(tmp_path / "h.pkl"), h, False
    MPC_ORD.save_history(Holder)
    raw = (tmp_path / "h.pkl").read_bytes()
    assert b"interact_drive.reward_design.mpc_ord" in raw and b"l4dc" not in raw
    back = pickle.loads(raw)
    assert back.seed == 11 and back[0][1] == -1.5


def test_sharding_arithmetic():
    for B in (0, 1, 7, 45, 90, 1000):
        for ws in (1, 2, 4, 8):
            seen = []
            for r in range(ws):
                lo, hi, per = parallel.shard_bounds(B, r, ws)
                idx = parallel.shard_indices(B, r, ws)
                assert len(idx) == per and (B == 0 or idx.max() <= B - 1)
                seen += list(range(lo, hi))
            assert seen == list(range(B))
    assert parallel.world() == (0, 1)


def test_synthetic_batch_and_flop_model():
    b = synthetic.make_batch(1000, C=4)
    assert b["world"].shape == (1000, 4, 4) and b["weights"].shape == (200, 7) and b["weight_idx"].max() == 199
    np.testing.assert_allclose(np.linalg.norm(b["weights"], axis=1), 1.0, atol=1e-6)
    assert np.all(np.abs(b["world"][:, 0, 0]) <= 0.1) and np.all(b["world"][:, :, 3] == np.float32(np.pi / 2))
    # 300 030 FLOP per finite_horizon solve, 337 740 per replanning solve
    assert synthetic.flops_per_solve(5, 2, 3) == 300030
    assert synthetic.flops_per_solve(5, 3, 2) == 337740


def test_cli_parsing_is_the_reference_one():
    with pytest.raises(SystemExit):
        run_mpc_ord.main(["nowhere", "cmaes"])
    with pytest.raises(AssertionError):
        run_mpc_ord.main(["finite_horizon", "cmaes", "--seed", "0"])
    assert set(run_mpc_ord.envs) == {"local_opt", "finite_horizon", "replanning"}
    assert run_mpc_ord.envs["replanning"]["num_eval_samples"] == 2 and run_mpc_ord.envs["replanning"]["eval_horizon"] == 20
    assert run_mpc_ord.fmt(np.array([1.0, 2.0])) == "[1. 2.]"


def test_install_as_reference_aliases():
    ocd.install_as_reference()
    import interact_drive.planner.naive_planner as npl
    import experiments.merging as mg
    from interact_drive.car import LinearRewardCar
    assert npl.NaivePlanner is ocd.interact_drive.planner.naive_planner.NaivePlanner
    assert issubclass(mg.ThreeLaneTestCar, LinearRewardCar)


def test_bench_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, check=True, timeout=600)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mpc_solves_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
    assert "workload" in line["config"]


def test_torch_custom_ops_register_and_refuse_cpu():
    import torch
    from l4dc_mpc_ocd_b200 import torch_ops as T
    p = ocd.PlannerParams(H=6, C=3, lane_x=(-0.05, 0.05), extra_inits=True)
    assert T.unpack_params(T.pack_params(p)) == p
    sc = ocd.Scenario(init_state=[[0, -0.7, 0.8, 1.57]] * 2, kind=[1, 0], friction=[0.2, 0.0],
                      control=[[0, 0], [0.1, 0.2]], plan=[[[0, 0], [0.7, 2.7]], []], critical_t=4)
    back = T.unpack_scenario(T.pack_scenario(sc))
    assert back.kind == [1, 0] and back.plan[0][1] == [0.7, 2.7] and back.plan[1] == [] and back.critical_t == 4
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():          # shape inference needs no device
        out = torch.ops.ocd_b200.solve(torch.empty((3, 4, 100)), torch.empty((6, 1)), None, None, T.pack_params(p))
        assert [tuple(o.shape) for o in out] == [(6, 2, 100), (6, 100), (100,)]
    with pytest.raises(RuntimeError):    # no CPU kernel behind the op
        torch.ops.ocd_b200.solve(torch.zeros((3, 4, 2)), torch.zeros((6, 1)), None, None, T.pack_params(p))


def test_first_order_ioc_host_math():
    """The optimisers of the IOC drop-ins are host computations on the Jacobian the engine returns:
    the SVD direction and Keras-style Adam on the normalised weights both recover a planted null vector."""
    from l4dc_mpc_ocd_b200.interact_drive.reward_design import first_order_ioc as F
    rng = np.random.default_rng(0)
    w = F.l2_normalize(rng.normal(size=7))
    A = rng.normal(size=(7, 12))
    A -= np.outer(w, w @ A)                       # every column orthogonal to w  =>  J^T w = 0
    assert abs(abs(F.nullspace_weights(A) @ w) - 1) < 1e-9
    theta = F.adam_on_sphere([A[:, :2], A[:, 2:]], np.ones(7), 1.0, 300)
    assert abs(abs(F.l2_normalize(theta) @ w) - 1) < 1e-6
    assert F.gradient_norm_loss(w, [A]) < 1e-20 < F.gradient_norm_loss(F.l2_normalize(np.ones(7)), [A])
    assert abs(np.linalg.norm(F.l2_normalize(np.zeros(3) + 1e-30))) <= 1.0


def test_kernel_form_selection(monkeypatch):
    """ocd_kernel_form is host logic: which kernel form a batch would run ("Form selection")."""
    import l4dc_mpc_ocd_b200 as ocd
    monkeypatch.delenv("OCD_KERNEL_FORM", raising=False)
    fh = ocd.PlannerParams()                                            # H=5, one other car: the bench shape
    assert [ocd.kernel_form(fh, B) for B in (1, 45, 1100, 1200, 20000, 30000, 1 << 20)] == \
        ["time-parallel", "time-parallel", "time-parallel", "latency", "latency", "wide", "wide"]
    assert [ocd.kernel_form(fh, B, episode=True) for B in (45, 5000, 60000, 100000)] == \
        ["time-parallel", "latency", "wide", "throughput"]
    rp = ocd.PlannerParams(C=3, lane_x=(-0.05, 0.05), num_lanes=2)      # replanning: two other cars
    assert [ocd.kernel_form(rp, B) for B in (180, 40000, 1 << 20)] == ["time-parallel", "latency", "wide"]
    assert ocd.kernel_form(rp, 100000, episode=True) == "throughput"
    six = ocd.PlannerParams(C=6)                                        # step-fenced wide form: big batches only
    assert [ocd.kernel_form(six, B) for B in (500, 30000, 65536, 262144)] == \
        ["time-parallel", "latency", "throughput", "wide"]
    # H = 15: the 16-lane time-parallel form up to ~1 800 warps (two problems x three starts per block)
    assert [ocd.kernel_form(ocd.PlannerParams(H=15), B) for B in (100, 1000, 1400, 40000, 1 << 20)] == \
        ["time-parallel", "time-parallel", "latency", "latency", "wide"]
    assert ocd.kernel_form(ocd.PlannerParams(H=50), 100) == "latency"      # no time-parallel form above 16 steps
    assert ocd.kernel_form(ocd.PlannerParams(H=50, C=6), 65536) == "wide"
    assert ocd.kernel_form(ocd.PlannerParams(H=15), 1000, episode=True) == "throughput"     # segmented episodes: one form
    assert ocd.kernel_form(ocd.PlannerParams(math_mode=ocd.MATH_PRECISE), 100) == "throughput"
    assert ocd.kernel_form(ocd.PlannerParams(optimizer=ocd.OPT_LBFGS), 100) == "throughput"
    for form, want in (("throughput", "throughput"), ("latency", "latency"), ("wide", "wide"), ("tp", "time-parallel")):
        monkeypatch.setenv("OCD_KERNEL_FORM", form)
        assert ocd.kernel_form(fh, 4096) == want
    monkeypatch.setenv("OCD_KERNEL_FORM", "tp")
    assert ocd.kernel_form(ocd.PlannerParams(H=15), 4096) == "time-parallel"   # 16 lanes per start up to H = 16
    assert ocd.kernel_form(ocd.PlannerParams(H=50), 4096) == "latency"   # no time-parallel form above H = 16: falls back
    with pytest.raises(ValueError):
        ocd.kernel_form(ocd.PlannerParams(H=65), 10)


# ---- CMA-ES: strategy parameters against the closed forms of Hansen's tutorial, trajectories against fixtures ---------
# (N, lambda, mu, mu_eff, c_sigma, d_sigma, c_c, c_1, c_mu, chi_N) from "The CMA Evolution Strategy: A Tutorial"
# (Hansen 2016), Table 1, evaluated independently of cmaes.py for the two weight dimensions of the scenarios
HANSEN_TABLE1 = [
    (6, 9, 4, 2.840610429717054, 0.34973966316715493, 1.349739663167155, 0.40864968827481546, 0.03563118207139872,
     0.03568631199675212, 2.3506677359645445),
    (7, 9, 4, 2.840610429717054, 0.3261732698019042, 1.3261732698019042, 0.3730062293365141, 0.027882099260254246,
     0.028450351990791156, 2.553831379703503),
]


@pytest.mark.parametrize("N,lam,mu,mueff,cs,ds,cc,c1,cmu,chi", HANSEN_TABLE1)
def test_cmaes_strategy_parameters_are_the_tutorial_defaults(N, lam, mu, mueff, cs, ds, cc, c1, cmu, chi):
    es = cmaes.CMAES([0.0] * N, 0.05, seed=1)
    assert (es.lam, es.mu) == (lam, mu)
    np.testing.assert_allclose(es.weights, [0.493738377484, 0.281096832481, 0.156709502558, 0.068455287477], atol=1e-12)
    got = (es.mueff, es.cs, es.damps, es.cc, es.c1, es.cmu, es.chiN)
    np.testing.assert_allclose(got, (mueff, cs, ds, cc, c1, cmu, chi), rtol=1e-13)
    assert es.weights.sum() == pytest.approx(1.0) and abs(es.c1 + es.cmu) < 1       # a convex covariance update


def test_cmaes_trajectories_match_the_fixtures():
    """Fixed-seed runs on sphere / rotated ellipsoid (tests/golden/make_cmaes_golden.py): sampling order, update and
    step-size path are pinned against accidental change (regression fixtures of the restatement, not pycma output)."""
    sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import make_cmaes_golden as G
    for tr in load_golden("cmaes_trajectories.json"):
        es = cmaes.CMAES(list(np.linspace(-0.5, 0.5, tr["N"])), tr["sigma0"], seed=tr["seed"])
        for row in tr["generations"]:
            pop = es.ask()
            np.testing.assert_allclose(pop[0], row["first_candidate"], rtol=1e-12, atol=1e-15)
            fit = [G.OBJ[tr["objective"]](x) for x in pop]
            es.tell(fit)
            assert min(fit) == pytest.approx(row["best_f"], rel=1e-10)
            assert es.sigma == pytest.approx(row["sigma"], rel=1e-10)
            np.testing.assert_allclose(es.mean, row["mean"], rtol=1e-9, atol=1e-12)
        assert tr["generations"][-1]["best_f"] < tr["generations"][0]["best_f"]


def test_cmaes_lockstep_runs_equal_serial_runs():
    """R independent runs advanced in lock step (one batched evaluation per generation for all of them) take exactly
    the steps they take alone: same candidates, same stopping generation, same result -- bit for bit."""
    def f(x):
        return float(np.sum((np.asarray(x) - 0.3) ** 2 * np.arange(1, len(x) + 1)))
    x0s = [[0.0] * 6, [0.1] * 6, [-0.2] * 6, [0.5] * 6]
    opts = [dict(seed=5, maxiter=12), dict(seed=6, maxiter=7), dict(seed=7, maxfevals=40), dict(seed=5, maxiter=12)]
    serial = [cmaes.fmin2(f, x0, 0.05, o) for x0, o in zip(x0s, opts)]
    calls = []

    def multi(pops):
        calls.append([len(p) for p in pops])
        return [np.array([f(x) for x in p]) for p in pops]

    lock = cmaes.fmin2_lockstep(multi, x0s, 0.05, opts)
    for (xs, es_s), (xl, es_l) in zip(serial, lock):
        assert np.array_equal(xs, xl) and es_s.countiter == es_l.countiter and es_s.counteval == es_l.counteval
        assert es_s.sigma == es_l.sigma and np.array_equal(es_s.mean, es_l.mean)
    assert len(calls) == 12 and calls[0] == [9, 9, 9, 9] and calls[-1] == [9, 0, 0, 9]     # runs drop out as they stop


def test_torch_autograd_restatement_agrees_with_the_c_oracle():
    """Two independent derivations of the planner's gradient: the C oracle's closed-form adjoint and torch autograd
    on an op-for-op restatement of mpc_reward (oracle/torch_serial.py, the `cpu_ref_serial` baseline).  Same plans."""
    import oracle as O
    from oracle import torch_serial as TS
    b = synthetic.make_batch(3, seed=5)
    w = b["weights"][b["weight_idx"]]
    ref = O.generate_plan_batch(O.OracleParams(n_iter=12), b["world"], w)
    prob = TS.Problem(n_iter=12)
    for i in range(3):
        plan, losses, best = prob.generate_plan(b["world"][i], w[i])
        assert best == ref["best"][i]
        np.testing.assert_allclose(plan, ref["plan"][i], atol=2e-6)
        np.testing.assert_allclose(losses, ref["losses"][i], rtol=2e-6, atol=2e-6)
    # the replanning shape: two lanes, two other cars with known controls, target speed 1.2
    b = synthetic.make_batch(2, C=3, lane_x=(-0.05, 0.05), seed=8)
    w = b["weights"][b["weight_idx"]]
    oc = 0.3 * synthetic.make_other_controls(2, 3, 5)
    op = O.OracleParams(C=3, lane_x=(-0.05, 0.05), num_lanes=2, other_mode=1, target_speed=1.2, n_iter=8)
    ref = O.generate_plan_batch(op, b["world"], w, other_controls=oc)
    prob = TS.Problem(lane_x=(-0.05, 0.05), num_lanes=2, target_speed=1.2, n_iter=8)
    for i in range(2):
        plan, losses, best = prob.generate_plan(b["world"][i], w[i], other_controls=oc[i])
        assert best == ref["best"][i]
        np.testing.assert_allclose(plan, ref["plan"][i], atol=2e-6)


def test_bind_rank_cpus_partitions_the_affinity_mask():
    import os
    if not hasattr(os, "sched_getaffinity"):
        pytest.skip("no affinity support")
    before = sorted(os.sched_getaffinity(0))
    try:
        if len(before) >= 2:
            n = parallel.bind_rank_cpus(1, 2)
            mine = sorted(os.sched_getaffinity(0))
            assert n == len(before) // 2 and mine == before[len(before) // 2:2 * (len(before) // 2)]
        os.sched_setaffinity(0, before)
        assert parallel.bind_rank_cpus(0, 1) == len(before) and sorted(os.sched_getaffinity(0)) == before
        assert parallel.bind_rank_cpus(0, len(before) + 1) == len(before)          # more ranks than cores: left alone
    finally:
        os.sched_setaffinity(0, before)


def test_reference_arm_never_maps_the_cuda_library():
    """bench.py --impl reference times the CPU oracle only: the process must not even load libocd_b200.so."""
    code = ("import sys; sys.argv=['bench.py']; sys.path.insert(0, %r); import bench; bench.cpu_arm(64, 1); "
            "maps = open('/proc/self/maps').read(); print('MAPPED' if 'libocd_b200' in maps else 'CLEAN', "
            "'ORACLE' if 'libocd_oracle' in maps else 'NOORACLE')" % str(ROOT))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.split()[-2:] == ["CLEAN", "ORACLE"], out.stdout


def test_plan_rank_cpus_follows_the_gpu_numa_nodes():
    """8 ranks on a two-socket host: GPUs 0-3 hang off node 0 (CPUs 0-31), GPUs 4-7 off node 1 (CPUs 32-63) -- every
    rank gets 8 cores of its own GPU's node, disjoint from every other rank's; without topology (or with a node that
    is the whole machine) the mask is cut contiguously by rank."""
    from l4dc_mpc_ocd_b200 import parallel
    allowed = list(range(64))
    aff = [list(range(32))] * 4 + [list(range(32, 64))] * 4
    plan = parallel.plan_rank_cpus(allowed, 8, aff)
    assert [len(p) for p in plan] == [8] * 8
    assert sorted(c for p in plan for c in p) == allowed
    assert all(set(plan[r]) <= set(aff[r]) for r in range(8))
    # interleaved GPU numbering (rank r's GPU on node r % 2) still lands every rank on its own node
    aff2 = [list(range(32)) if r % 2 == 0 else list(range(32, 64)) for r in range(8)]
    plan2 = parallel.plan_rank_cpus(allowed, 8, aff2)
    assert all(set(plan2[r]) <= set(aff2[r]) for r in range(8))
    assert sorted(c for p in plan2 for c in p) == allowed
    # restricted mask: only what the process may use is handed out
    plan3 = parallel.plan_rank_cpus(list(range(16, 48)), 2, [list(range(32)), list(range(32, 64))])
    assert plan3 == [list(range(16, 32)), list(range(32, 48))]
    # no topology, or one node = all CPUs: contiguous by rank
    assert parallel.plan_rank_cpus(allowed, 4, None) == [list(range(16 * r, 16 * r + 16)) for r in range(4)]
    assert parallel.plan_rank_cpus(allowed, 4, [allowed] * 4) == [list(range(16 * r, 16 * r + 16)) for r in range(4)]
    # a rank with unknown topology sends everyone to the contiguous rule (no core handed out twice)
    plan4 = parallel.plan_rank_cpus(allowed, 4, [list(range(32)), None, list(range(32, 64)), list(range(32, 64))])
    assert sorted(c for p in plan4 for c in p) == allowed


def test_cioc_augmented_loss_matches_a_direct_evaluation():
    """LocalCIOC's host part (second_order_ioc.py:149-165 of the reference): the loss assembled from per-feature gradient
    rows and Hessian rows equals a direct numpy evaluation, and its autograd gradient in (weights, theta_r) equals
    central differences."""
    import torch
    from l4dc_mpc_ocd_b200.interact_drive.reward_design.second_order_ioc import LocalCIOC
    rng = np.random.default_rng(3)
    K, n = 4, 6
    G = rng.normal(size=(K, n))
    Hm = rng.normal(size=(K, n, n))
    w = rng.normal(size=K)
    th, mu, lm = 0.7, 10.0, 0.3

    def direct(w, th):
        g = (w[:, None] * G).sum(0)
        A = (w[:, None, None] * Hm).sum(0) - th * np.eye(n)
        s, la = np.linalg.slogdet(-A)
        return -(0.5 * g @ np.linalg.solve(A, g) + 0.5 * s * la - 0.5 * mu * th ** 2 + lm * th), s

    wt = torch.tensor(w, requires_grad=True)
    tt = torch.tensor(th, dtype=torch.float64, requires_grad=True)
    loss, sign = LocalCIOC.augmented_loss(wt, tt, torch.as_tensor(G), torch.as_tensor(Hm), mu, lm)
    ref, s = direct(w, th)
    assert abs(float(loss) - ref) < 1e-10 and float(sign) == s
    loss.backward()
    h = 1e-6
    for i in range(K):
        e = np.zeros(K); e[i] = h
        fd = (direct(w + e, th)[0] - direct(w - e, th)[0]) / (2 * h)
        assert abs(float(wt.grad[i]) - fd) < 1e-5 * max(1.0, abs(fd))
    fd = (direct(w, th + h)[0] - direct(w, th - h)[0]) / (2 * h)
    assert abs(float(tt.grad) - fd) < 1e-5 * max(1.0, abs(fd))


def test_packaging_metadata_lists_every_subpackage():
    """pyproject.toml maps the importable name onto `l4dc-mpc-ocd_b200/`; its package list must name every directory
    with an __init__.py there, or an installed copy would miss modules the in-tree alias finds."""
    import tomllib
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    meta = tomllib.loads((root / "pyproject.toml").read_text())["tool"]["setuptools"]
    src = root / meta["package-dir"]["l4dc_mpc_ocd_b200"]
    found = {"l4dc_mpc_ocd_b200" + "".join("." + p for p in d.parent.relative_to(src).parts)
             for d in src.rglob("__init__.py") if "csrc" not in d.parts}
    assert set(meta["packages"]) == found


def test_row_wise_bookkeeping_equals_the_scalar_path():
    """The lock-step evaluation normalises all candidates in one call and sums all runs' returns in one reduction;
    both must give, bit for bit, what the per-candidate / per-run code of a serial evaluation gives (np.vecdot runs the
    same BLAS dot np.linalg.norm takes per vector; a reduction over the trailing axes is the per-row one)."""
    from l4dc_mpc_ocd_b200.interact_drive.reward_design import mpc_ord as M
    rng = np.random.default_rng(11)
    for K in (6, 7):
        W = rng.normal(size=(500, K)) * np.exp(rng.normal(size=(500, 1)) * 4)
        unit = M._unit_rows(W)
        plan = M._planning_weight_rows(W)
        for w, u, p in zip(W, unit, plan):
            assert np.array_equal(u, w / np.linalg.norm(w))
            assert np.array_equal(p.view(np.int32), M.MPC_ORD._planning_weights(w).view(np.int32))
            x = w
            for _ in range(3):
                x = x / np.linalg.norm(x)                       # the reference's three normalisations, literally
            assert np.array_equal(p, x.astype(np.float32))
    for ni, ns in ((5, 1), (10, 1), (5, 2), (32, 1)):
        ret = (rng.normal(size=(7, 9, ni, ns)) * 30).astype(np.float32)
        one = ret.reshape(-1, ni * ns).sum(axis=1, dtype=np.float64).reshape(7, 9)
        for r in range(7):
            assert np.array_equal(one[r], ret[r].sum(axis=(1, 2), dtype=np.float64))


def test_cmaes_batch_reproduces_the_single_run_class_bit_for_bit():
    """`CMAESBatch` (stacked arrays, what the lock-step optimisation drives) against `CMAES` (plain 2-D numpy, what
    `optimize_cmaes` drives): populations, means, step sizes and covariance matrices stay bit-identical over long runs,
    with the runs of the batch asked / told in changing subsets."""
    def ell(x):
        n = len(x)
        return float(sum((10.0 ** (3.0 * i / (n - 1))) * x[i] ** 2 for i in range(n)))
    for N in (6, 7):
        x0 = list(np.linspace(-0.5, 0.5, N))
        seeds = [1, 2, 3, 12345, 77]
        singles = [cmaes.CMAES(x0, 0.05, seed=s) for s in seeds]
        batch = cmaes.CMAESBatch([x0] * len(seeds), 0.05, seeds)
        for g in range(120):
            idx = np.array([r for r in range(len(seeds)) if (g + r) % 4 != 0 or g < 10])      # a changing subset
            pops = batch.ask(idx)
            fits = []
            for r, pb in zip(idx, pops):
                ps = singles[r].ask()
                assert np.array_equal(ps, pb)
                fit = [ell(x) for x in ps]
                singles[r].tell(fit)
                fits.append(fit)
            batch.tell(np.array(fits), idx)
            for r in idx:
                s = singles[r]
                assert s.sigma == batch.sigma[r] and np.array_equal(s.mean, batch.mean[r])
                assert np.array_equal(s.C, batch.C[r]) and np.array_equal(s.invsqrtC, batch.invsqrtC[r])
            assert batch.stop(idx, maxiter=100) == [singles[r].stop(maxiter=100) for r in idx]


class _OracleHostContext:
    """Stands in for the engine's host context (`ocd_episode_batch_host`) on CPU: same SoA call, oracle episodes."""

    def __init__(self, O, spec):
        self.O, self.spec = O, spec

    def episodes_soa(self, p, sc, robot_init, plan_weights, true_weights, T, weight_idx=None, unlucky_idx=None,
                     final_world=False, **_):
        ri = np.ascontiguousarray(np.asarray(robot_init, np.float32).T)
        W = np.asarray(plan_weights, np.float32).T[np.asarray(weight_idx)]
        ret = self.O.episode_batch(self.spec.params, self.spec.scenario, ri, W, np.asarray(true_weights, np.float32), T,
                                   nthreads=2)
        fw = np.zeros((p.C, 4, ri.shape[0]), np.float32)
        fw[0] = ri.T
        return (ret, fw) if final_world else ret


def test_lockstep_host_bookkeeping_equals_serial_runs(monkeypatch):
    """The host side of `optimize_cmaes_lockstep` (stacked CMA-ES state, one normalisation for all candidates, totals in
    one reduction, object state written at the end) against the same runs done one after the other with
    `optimize_cmaes` -- histories, iteration counters, results and the state left in the car, bit for bit.  Episodes come
    from the CPU oracle behind the host-context interface, so this guards the Python bookkeeping without a GPU."""
    import oracle as O
    import l4dc_mpc_ocd_b200.runtime as RT
    from l4dc_mpc_ocd_b200.interact_drive.reward_design import mpc_ord as M
    spec = O.scenario_params("finite_horizon")
    monkeypatch.setattr(RT, "get_host_context", lambda device=None: _OracleHostContext(O, spec))
    monkeypatch.setattr(M, "get_engine", lambda device=None: None)      # asked for, never used on these paths
    _, _, inits = M.finite_horizon_env(env_seeds=[1000000 + i for i in range(6)], debug=False)

    def fresh():
        runs = []
        for g in (inits[0:2], inits[2:4], inits[4:6]):
            car, world, _ = M.finite_horizon_env(debug=False)
            runs.append(M.MPC_ORD(world, car, g, 3, verbose=False))
        return runs

    seeds = [5, 6, 7]
    serial = fresh()
    xs = [r.optimize_cmaes(seed=s, sigma0=0.05, maxfevals=27) for r, s in zip(serial, seeds)]
    lock = fresh()
    xl = M.optimize_cmaes_lockstep(lock, seeds, sigma0=0.05, maxfevals=27)
    for a, b, xa, xb in zip(serial, lock, xs, xl):
        assert np.array_equal(xa, xb) and a.iter == b.iter == 28 and b.done
        assert len(a.history) == len(b.history) == 28
        for (wa, va), (wb, vb) in zip(a.history, b.history):
            assert np.array_equal(wa, wb) and va == vb
        assert np.array_equal(a.car.weights, b.car.weights) and np.array_equal(a.car.init_state, b.car.init_state)
        assert np.array_equal(a.world.cars[0].state, b.world.cars[0].state)
    assert lock[0].kernel_launches == 1 + 3 and all(r.kernel_launches == 1 + 3 for r in serial)
