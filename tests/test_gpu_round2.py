"""Round-2 parity additions (all through the C ABI): whole-batch checks that do not depend on the
conditioning filter of tests/test_gpu_parity.py, the exact-size CMA-ES generations of BASELINE configs[2] and
configs[3] against the oracle, the compile-time H = 15 / 50 kernels against the runtime-horizon kernels, the
lock-step multi-run CMA-ES against serial runs, and the new host entry point."""
import numpy as np
import pytest
import torch

import oracle as O
from conftest import load_golden  # noqa: F401

pytestmark = pytest.mark.gpu

import l4dc_mpc_ocd_b200 as ocd            # noqa: E402
from l4dc_mpc_ocd_b200 import synthetic    # noqa: E402
from l4dc_mpc_ocd_b200.experiments import run_mpc_ord    # noqa: E402
from l4dc_mpc_ocd_b200.interact_drive.reward_design.mpc_ord import (    # noqa: E402
    MPC_ORD, eval_weights_lockstep, optimize_cmaes_lockstep)

from test_gpu_parity import MODES, OBJ_TOL, U_TOL, _pp    # noqa: E402


# ---- whole batch, no conditioning filter ---------------------------------------------------------------------------
@pytest.mark.parametrize("mode,_n", MODES)
@pytest.mark.parametrize("H,C,lanes,other_mode,extra,lr", [
    (5, 2, 3, 0, False, 0.1), (5, 3, 2, 1, False, 0.1), (6, 2, 3, 0, True, 0.1), (5, 6, 3, 0, False, 0.1),
    (15, 2, 3, 0, False, 0.03), (15, 4, 3, 0, False, 0.03), (50, 2, 3, 0, False, 0.0003), (50, 4, 3, 0, False, 0.0003),
    (12, 3, 3, 1, False, 0.03)])
def test_two_iterations_every_problem_every_start(engine, mode, _n, H, C, lanes, other_mode, extra, lr):
    """Two SGD iterations amplify nothing, so there is no ill-conditioned subset to exclude: EVERY problem and EVERY
    start of a random batch must match the f32 oracle -- controls after the updates (all_plans), final losses and
    winners.  This pins rollout, feature gradient, adjoint, update and loss on the whole batch; the fixed-budget tests
    then add the iteration count on the problems where the reference is reproducible."""
    B = 2048
    lane_x = (-0.1, 0.0, 0.1) if lanes == 3 else (-0.05, 0.05)
    batch = synthetic.make_batch(B, C=C, lane_x=lane_x, seed=100 + H + C)
    oc = 0.3 * synthetic.make_other_controls(B, C, H) if other_mode else None
    op = O.OracleParams(H=H, C=C, lane_x=lane_x, n_iter=2, num_lanes=lanes, other_mode=other_mode, extra_inits=extra,
                        target_speed=1.0 if lanes == 3 else 1.2, lr=lr)
    w_full = batch["weights"][batch["weight_idx"]]
    one = [O.generate_plan(op, batch["world"][b], w_full[b], other_controls=None if oc is None else oc[b], all_plans=True)
           for b in range(B)]
    ref = {k: np.stack([np.asarray(o[k]) for o in one]) for k in ("plan", "losses", "best", "all_plans")}
    res = engine.solve(_pp(op, mode), batch["world"], batch["weights"], weight_idx=batch["weight_idx"],
                       other_controls=oc, all_plans=True)
    res = {k: v.cpu().numpy() for k, v in res.items()}
    # two updates of size lr * gradient: the controls carry the gradient's error scaled by lr
    tol_u = (2e-5 if mode == ocd.MATH_PRECISE else 2e-4) * max(1.0, lr / 0.1)
    assert np.max(np.abs(res["all_plans"] - ref["all_plans"])) <= tol_u
    # the loss the engine reports for ITS controls is the objective of those controls (f64 oracle at the engine's
    # controls: a difference in the controls, allowed above, times a large gradient is not a loss error)
    w64 = w_full.astype(np.float64)
    for s_ in range(op.S):
        R = np.array([O.mpc_reward(op, batch["world"][b].astype(np.float64), res["all_plans"][b, s_].astype(np.float64), w64[b],
                                   other_controls=None if oc is None else oc[b].astype(np.float64), dtype=np.float64,
                                   grad=False) for b in range(0, B, 4)])
        got = res["losses"][::4, s_]
        assert np.max(np.abs(got + R) / np.maximum(1.0, np.abs(R))) <= OBJ_TOL[mode]
    # and it is close to the oracle's own loss: within the objective tolerance plus what the control tolerance can move it
    rel = np.abs(res["losses"] - ref["losses"]) / np.maximum(1.0, np.abs(ref["losses"]))
    assert rel.max() <= OBJ_TOL[mode] + 50.0 * tol_u
    # winners: identical wherever the oracle's two best losses are further apart than that
    srt = np.sort(ref["losses"], axis=1)
    clear = (srt[:, 1] - srt[:, 0]) > 4 * (OBJ_TOL[mode] + 50.0 * tol_u) * np.maximum(1.0, np.abs(srt[:, 0]))
    assert clear.mean() > 0.25           # non-vacuity only: after two iterations many starts (extra_inits: pairs of them) are still near ties
    assert np.array_equal(res["best"][clear], ref["best"][clear])


@pytest.mark.parametrize("H,C,lr", [(5, 2, 0.1), (15, 2, 0.03), (50, 2, 0.0003), (5, 5, 0.1)])
def test_whole_batch_self_consistency(engine, H, C, lr):
    """Full budget (100 iterations), every problem, ill-conditioned or not: whatever plan the engine returns, the
    loss it reports for it is the loss of THAT plan (recomputed by the f64 oracle), the winner is the first minimum of
    the reported losses, and no output is left unwritten."""
    B = 4096
    batch = synthetic.make_batch(B, C=C, seed=500 + H)
    p = ocd.PlannerParams(H=H, C=C, lr=lr)
    res = engine.solve(p, batch["world"], batch["weights"], weight_idx=batch["weight_idx"], all_plans=True)
    res = {k: v.cpu().numpy() for k, v in res.items()}
    assert np.array_equal(res["best"], np.argmin(res["losses"], axis=1))
    assert np.array_equal(res["plan"], res["all_plans"][np.arange(B), res["best"]])
    op = O.OracleParams(H=H, C=C, lr=lr)
    w_full = batch["weights"][batch["weight_idx"]].astype(np.float64)
    sel = np.arange(0, B, 8)
    for s in range(3):
        R = np.array([O.mpc_reward(op, batch["world"][b].astype(np.float64), res["all_plans"][b, s].astype(np.float64),
                                   w_full[b], dtype=np.float64, grad=False) for b in sel])
        got = res["losses"][sel, s]
        fin = np.isfinite(R)
        assert np.array_equal(fin, np.isfinite(got))
        assert np.max(np.abs(got[fin] + R[fin]) / np.maximum(1.0, np.abs(R[fin]))) <= 1e-4


# ---- compile-time H = 15 / 50 kernels against the runtime-horizon kernels ---------------------------------------------
@pytest.mark.parametrize("H,C,lr,n_iter", [(15, 2, 0.03, 100), (15, 5, 0.03, 40), (50, 2, 0.0003, 100), (50, 6, 0.0003, 30)])
def test_compile_time_horizons_match_the_runtime_kernels(engine, monkeypatch, H, C, lr, n_iter):
    """H = 15 runs the Q kernels (saved step data through shared memory), H = 50 the segmented adjoint with a constant
    segment count -- both with the register-resident kernels' folded feature arithmetic; OCD_RUNTIME_H=1 sends the same
    problems through the runtime-horizon segmented kernels every other horizon uses.  Different rounding (the folded
    lane term, one reciprocal per bump), same plans: on the problems where the reference is reproducible at all."""
    B = 3000
    batch = synthetic.make_batch(B, C=C, seed=77)
    p = ocd.PlannerParams(H=H, C=C, lr=lr, n_iter=n_iter)
    new = engine.solve(p, batch["world"], batch["weights"], weight_idx=batch["weight_idx"])
    monkeypatch.setenv("OCD_RUNTIME_H", "1")
    old = engine.solve(p, batch["world"], batch["weights"], weight_idx=batch["weight_idx"])
    monkeypatch.delenv("OCD_RUNTIME_H")
    du = (new["plan"] - old["plan"]).abs().amax(dim=(1, 2)).cpu().numpy()
    same_best = (new["best"] == old["best"]).cpu().numpy()
    # ill-conditioned problems by the oracle's own probe: f32 vs f64 on a sample
    sel = np.arange(0, B, 10)
    op = O.OracleParams(H=H, C=C, lr=lr, n_iter=n_iter)
    w = batch["weights"][batch["weight_idx"]][sel]
    r32 = O.generate_plan_batch(op, batch["world"][sel], w)
    r64 = O.generate_plan_batch(op, batch["world"][sel].astype(np.float64), w.astype(np.float64), dtype=np.float64)
    well = (np.abs(r32["plan"] - r64["plan"]).reshape(len(sel), -1).max(1) < 1e-5) & (r32["best"] == r64["best"])
    assert well.mean() > 0.5
    # (the f32-vs-f64 probe alone misses a few problems that fast-math-sized perturbations still tip over a kink)
    assert np.mean(du[sel][well] <= 1e-3) >= 0.98 and np.mean(same_best[sel][well]) >= 0.98
    assert np.mean(du <= 1e-3) >= 0.9 * well.mean()
    # and both agree with the oracle there
    for res in (new, old):
        plan = res["plan"].cpu().numpy()[sel][well]
        assert np.mean(np.abs(plan - r32["plan"][well]).reshape(plan.shape[0], -1).max(1) <= 1e-3) >= 0.98


def test_compile_time_horizon_forms_are_bit_identical(engine, monkeypatch):
    """Throughput, latency and wide form of the H = 15 and H = 50 kernels run the same arithmetic."""
    for H, lr in ((15, 0.03), (50, 0.0003)):
        batch = synthetic.make_batch(2000, C=2, seed=31)
        p = ocd.PlannerParams(H=H, C=2, lr=lr, n_iter=30)
        outs = {}
        for form in ("throughput", "latency", "wide"):
            monkeypatch.setenv("OCD_KERNEL_FORM", form)
            outs[form] = engine.solve(p, batch["world"], batch["weights"], weight_idx=batch["weight_idx"], all_plans=True)
        monkeypatch.delenv("OCD_KERNEL_FORM")
        for form in ("latency", "wide"):
            for k in ("all_plans", "losses", "best"):
                assert torch.equal(outs["throughput"][k], outs[form][k]), (H, form, k)


# ---- BASELINE configs[2] / configs[3] at their exact sizes, through the drop-in ------------------------------------------
@pytest.mark.parametrize("scenario,n_inits", [("local_opt", 10), ("replanning", 5), ("finite_horizon", 5)])
def test_exact_size_generation_vs_oracle(scenario, n_inits):
    """One CMA-ES generation exactly as `run_mpc_ord.py <scenario> cmaes --n_inits n` evaluates it: popsize 9 candidates
    x n_inits initial states x num_eval_samples episodes of eval_horizon control steps (local_opt: 90 episodes x 15
    steps; replanning: 90 episodes x 20 steps, both vanishing cars), through MPC_ORD.eval_weights_batch -- one launch --
    against the oracle's serial episodes.  Per-candidate returns within BASELINE's 1e-3 relative."""
    env = run_mpc_ord.envs[scenario]
    car, world, inits = env["make_env"](env_seeds=[(1000000 + i) for i in range(n_inits)], debug=False)
    m = MPC_ORD(world, car, inits, env["eval_horizon"], num_samples=env["num_eval_samples"], verbose=False)
    rng = np.random.default_rng(3)
    cands = [m.designer_weights + 0.05 * rng.normal(size=m.weight_dim) for _ in range(9)]
    neg = m.eval_weights_batch(cands)
    assert m.kernel_launches == 1 and len(m.history) == 9
    spec = O.scenario_params(scenario)
    ns, T = env["num_eval_samples"], env["eval_horizon"]
    assert (spec.num_samples, spec.eval_horizon) == (ns, T)
    W = np.stack([MPC_ORD._planning_weights(c) for c in cands])
    ri = np.repeat(np.tile(np.asarray(inits, np.float32), (9, 1)), ns, axis=0)
    widx = np.repeat(np.arange(9), n_inits * ns)
    # ReplanningCarWorld toggles the vanishing car on every reset; the constructor's reset leaves it at 2, so the
    # first evaluated sample has unlucky_car_idx = 1 (replanning_world.py:19-27, quirk Q5)
    ul = (1 + (np.arange(ri.shape[0]) % 2)).astype(np.int32) if scenario == "replanning" else None
    ref = O.episode_batch(spec.params, spec.scenario, ri, W[widx], m.designer_weights.astype(np.float32), T, unlucky_idx=ul)
    want = ref.reshape(9, -1).sum(1) / ns
    assert np.max(np.abs(-neg - want) / np.abs(want)) <= 1e-3, (-neg, want)


def test_replanning_toggle_runs_over_candidates():
    """With an odd number of resets per candidate the reference's serial loop hands consecutive candidates opposite
    vanishing cars; the batch reproduces that sequence (no per-candidate tiling)."""
    env = run_mpc_ord.envs["replanning"]
    car, world, inits = env["make_env"](env_seeds=[1000000, 1000001, 1000002], debug=False)
    m = MPC_ORD(world, car, inits, 6, num_samples=1, verbose=False)
    b = m._episode_batch([m.designer_weights, m.designer_weights])
    assert b["unlucky"].tolist() == [1, 2, 1, 2, 1, 2] and world.unlucky_car_idx == 2
    # serial evaluation of the same two candidates gives the same returns as the batch
    car2, world2, _ = env["make_env"](env_seeds=[1000000, 1000001, 1000002], debug=False)
    m2 = MPC_ORD(world2, car2, inits, 6, num_samples=1, verbose=False)
    serial = np.array([m2.eval_weights(m2.designer_weights), m2.eval_weights(m2.designer_weights)])
    car3, world3, _ = env["make_env"](env_seeds=[1000000, 1000001, 1000002], debug=False)
    m3 = MPC_ORD(world3, car3, inits, 6, num_samples=1, verbose=False)
    batch = m3.eval_weights_batch([m3.designer_weights, m3.designer_weights])
    np.testing.assert_array_equal(batch, serial)
    assert serial[0] != serial[1]          # the two candidates really saw different worlds


# ---- lock-step multi-run CMA-ES ----------------------------------------------------------------------------------------
def _fresh_runs(scenario, groups, T):
    env = run_mpc_ord.envs[scenario]
    runs = []
    for g in groups:
        car, world, _ = env["make_env"](debug=False)
        runs.append(MPC_ORD(world, car, g, T, num_samples=env["num_eval_samples"], verbose=False))
    return runs


@pytest.mark.parametrize("scenario", ["finite_horizon", "replanning"])
def test_lockstep_cmaes_equals_serial_runs(scenario):
    """R independent CMA-ES runs (the reference: one worker process each, run_mpc_ord.py:83-90) advanced in lock step,
    ONE episode launch per generation for all of them, produce bit for bit the histories of the same runs done one
    after the other -- including runs that stop early."""
    env = run_mpc_ord.envs[scenario]
    _, _, inits = env["make_env"](env_seeds=[1000000 + i for i in range(5)], debug=False)
    groups = [inits[0:2], inits[2:3], inits[3:5]]
    seeds, budgets = [5, 6, 5], [27, 9, 18]
    T = 6
    serial = _fresh_runs(scenario, groups, T)
    for r, seed, mf in zip(serial, seeds, budgets):
        r.optimize_cmaes(seed=seed, sigma0=0.05, maxfevals=mf)
    lock = _fresh_runs(scenario, groups, T)
    from l4dc_mpc_ocd_b200 import cmaes
    # per-run budgets: the lock-step driver takes one stop dict for all runs, so drive fmin2_lockstep directly
    for r, seed in zip(lock, seeds):
        r.history.seed = seed
    eval_weights_lockstep(lock, [[r.designer_weights] for r in lock])
    cmaes.fmin2_lockstep(lambda pops: eval_weights_lockstep(lock, pops), [list(r.designer_weights) for r in lock], 0.05,
                         [dict(seed=s, maxfevals=mf) for s, mf in zip(seeds, budgets)])
    assert lock[0].kernel_launches == 1 + 3            # designer weights + the longest run's three generations
    for a, b in zip(serial, lock):
        assert len(a.history) == len(b.history)
        for (wa, va), (wb, vb) in zip(a.history, b.history):
            np.testing.assert_array_equal(wa, wb)
            assert va == vb
    # the public wrapper
    again = _fresh_runs(scenario, groups, T)
    xs = optimize_cmaes_lockstep(again, [5, 6, 5], sigma0=0.05, maxfevals=18)
    assert len(xs) == 3 and all(r.done for r in again) and len(again[0].history) == 1 + 18


def test_lockstep_rejects_mismatched_runs():
    a = _fresh_runs("finite_horizon", [[np.array([0.0, -0.9, 0.8, np.pi / 2])]], 5)[0]
    b = _fresh_runs("local_opt", [[np.array([-0.1, -0.9, 1.0, np.pi / 2])]], 5)[0]
    with pytest.raises(ValueError):
        eval_weights_lockstep([a, b], [[a.designer_weights], [b.designer_weights]])


# ---- host entry points -----------------------------------------------------------------------------------------------
def test_solve_first_host_returns_the_first_control(engine):
    B = 70000                                           # several pipeline chunks
    batch = synthetic.make_batch(B, seed=21)
    p = ocd.PlannerParams(n_iter=20)
    ctx = ocd.HostContext(0)
    world = np.ascontiguousarray(batch["world"].transpose(1, 2, 0))
    w = np.ascontiguousarray(batch["weights"].T)
    full = ctx.solve_soa(p, world, w, weight_idx=batch["weight_idx"])
    first = ctx.solve_first_soa(p, world, w, weight_idx=batch["weight_idx"])
    np.testing.assert_array_equal(first["first"], full["plan"][0])
    np.testing.assert_array_equal(first["losses"], full["losses"])
    np.testing.assert_array_equal(first["best"], full["best"])
    only = ctx.solve_first_soa(p, world, w, weight_idx=batch["weight_idx"], losses=False, best=False)
    assert set(only) == {"first"}
    np.testing.assert_array_equal(only["first"], full["plan"][0])
    # pinned outputs take the in-place path
    pin = dict(first=ocd.HostContext.pinned_empty((2, B)), best=ocd.HostContext.pinned_empty((B,), np.int32))
    got = ctx.solve_first_soa(p, world, w, weight_idx=batch["weight_idx"], out=pin)
    np.testing.assert_array_equal(got["first"], full["plan"][0])
    np.testing.assert_array_equal(got["best"], full["best"])
    ctx.close()


def test_weight_idx_out_of_range(engine):
    batch = synthetic.make_batch(64, seed=2)
    p = ocd.PlannerParams(n_iter=3)
    bad = batch["weight_idx"].copy()
    bad[5] = batch["weights"].shape[0]
    with pytest.raises(ValueError):
        engine.solve(p, batch["world"], batch["weights"], weight_idx=bad)
    ctx = ocd.HostContext(0)
    with pytest.raises(ValueError):
        ctx.solve_soa(p, np.ascontiguousarray(batch["world"].transpose(1, 2, 0)), np.ascontiguousarray(batch["weights"].T),
                      weight_idx=bad)
    bad[5] = -1
    with pytest.raises(ValueError):
        ctx.episodes_soa(p, ocd.Scenario(init_state=[[0.0, -0.6, 0.5, np.pi / 2]], kind=[0], friction=[0.0],
                                         control=[[0.0, 0.0]]),
                         np.ascontiguousarray(batch["world"][:, 0].T), np.ascontiguousarray(batch["weights"].T),
                         batch["weights"][0], 2, weight_idx=bad)
    ctx.close()
    # a device-resident index the launcher cannot inspect is clamped by the kernels: no fault, the last column is used
    dev_idx = torch.as_tensor(bad, device=engine.device)
    dev_idx[5] = 10 ** 6
    ws = torch.as_tensor(batch["world"], device=engine.device).permute(1, 2, 0).contiguous()
    wt = torch.as_tensor(batch["weights"], device=engine.device).t().contiguous()
    out = engine.solve_soa(p, ws, wt, wt.shape[1], dev_idx)
    good = batch["weight_idx"].copy()
    good[5] = batch["weights"].shape[0] - 1
    ref = engine.solve_soa(p, ws, wt, wt.shape[1], torch.as_tensor(good, device=engine.device))
    assert torch.equal(out["plan"], ref["plan"])


def test_soa_entry_points_refuse_wrong_tensors(engine):
    batch = synthetic.make_batch(32, seed=2)
    p = ocd.PlannerParams(n_iter=2)
    ws = torch.as_tensor(batch["world"], device=engine.device).permute(1, 2, 0).contiguous()
    wt = torch.as_tensor(batch["weights"], device=engine.device).t().contiguous()
    idx = torch.as_tensor(batch["weight_idx"], device=engine.device)
    engine.solve_soa(p, ws, wt, wt.shape[1], idx)
    for bad in (dict(world=ws.double()), dict(world=ws.cpu()), dict(weights=wt[:-1]), dict(idx=idx.long()),
                dict(world=ws.permute(1, 0, 2))):
        with pytest.raises(ValueError):
            engine.solve_soa(p, bad.get("world", ws), bad.get("weights", wt), wt.shape[1], bad.get("idx", idx))


def test_registered_host_arrays_take_the_in_place_path():
    """ocd_host_register page-locks a caller's ordinary arrays in place: same results as the staged path, and the
    registration can be undone; registering twice is an error, not a crash."""
    B = 20000
    p = ocd.PlannerParams()
    b = synthetic.make_batch(B, seed=8)
    world = np.ascontiguousarray(b["world"].transpose(1, 2, 0))
    w = np.ascontiguousarray(b["weights"].T)
    idx = b["weight_idx"]
    ctx = ocd.HostContext(0)
    ref = ctx.solve_soa(p, world, w, weight_idx=idx)
    ref = {k: np.array(v) for k, v in ref.items()}
    out = dict(plan=np.zeros((p.H, 2, B), np.float32), losses=np.zeros((p.S, B), np.float32), best=np.zeros((B,), np.int32))
    arrays = [world, w, idx] + list(out.values())
    for a in arrays:
        ocd.HostContext.register(a)
    with pytest.raises(ocd.OcdCudaError):
        ocd.HostContext.register(world)
    got = ctx.solve_soa(p, world, w, weight_idx=idx, out=out)
    for k in ref:
        assert np.array_equal(ref[k], got[k])
    for a in arrays:
        ocd.HostContext.unregister(a)
    got = ctx.solve_soa(p, world, w, weight_idx=idx, out=out)          # pageable again: staged path, same answer
    for k in ref:
        assert np.array_equal(ref[k], got[k])
    ctx.close()
