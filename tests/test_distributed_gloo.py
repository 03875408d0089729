"""The N>1 path on CPU: two gloo ranks shard a batch of episodes, each evaluates its shard (here
with the CPU oracle standing in for the GPU kernel -- the plumbing under test is the sharding and
the all-gather of returns), and every rank must end up with the full, correctly ordered vector."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle as O
    from l4dc_mpc_ocd_b200 import parallel
    spec = O.scenario_params("finite_horizon")
    w = (spec.designer_weights / np.linalg.norm(spec.designer_weights)).astype(np.float32)
    rng = np.random.default_rng(5)
    ri = np.tile(spec.example_init, (B, 1)).astype(np.float32)
    ri[:, 0] += rng.uniform(-0.02, 0.02, B).astype(np.float32)
    assert parallel.world() == (rank, ws)

    def evaluate(idx):
        return torch.from_numpy(O.episode_batch(spec.params, spec.scenario, ri[idx], w, w, 4, nthreads=1))

    full = parallel.sharded_returns(evaluate, B).numpy()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full)
    if rank == 0:
        np.save(os.path.join(out_dir, "ref.npy"), O.episode_batch(spec.params, spec.scenario, ri, w, w, 4, nthreads=1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_returns(tmp_path):
    B, ws = 11, 2            # not divisible: the last rank pads
    mp.spawn(_worker, args=(ws, _free_port(), B, str(tmp_path)), nprocs=ws, join=True)
    ref = np.load(tmp_path / "ref.npy")
    for r in range(ws):
        got = np.load(tmp_path / f"r{r}.npy")
        assert got.shape == (B,)
        np.testing.assert_array_equal(got, ref)


class _OracleEngine:
    """Stands in for the GPU engine on CPU ranks: the oracle evaluates the episodes; the final world of episode b is
    tagged with b so that the test can see which episode's state each rank is left with."""

    def __init__(self, O, spec, offset):
        self.O, self.spec, self.offset = O, spec, offset

    def episodes(self, p, sc, robot_init, plan_weights, true_weights, T, weight_idx=None, unlucky_idx=None,
                 final_world=False, **_):
        ri = np.asarray(robot_init, np.float32)
        W = np.asarray(plan_weights, np.float32)[np.asarray(weight_idx)]
        ret = self.O.episode_batch(self.spec.params, self.spec.scenario, ri, W, np.asarray(true_weights, np.float32), T,
                                   nthreads=1)
        fw = np.zeros((ri.shape[0], p.C, 4), np.float32)
        fw[:, 0, :] = ri                     # tag: the episode's own initial state
        return dict(returns=torch.from_numpy(ret), final_world=torch.from_numpy(fw))


def _lockstep_worker(rank, ws, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import pickle
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle as O
    from l4dc_mpc_ocd_b200.interact_drive.reward_design import mpc_ord as M
    spec = O.scenario_params("finite_horizon")
    M.get_engine = lambda device=None: _OracleEngine(O, spec, 0)
    _, _, inits = M.finite_horizon_env(env_seeds=[1000000 + i for i in range(5)], debug=False)
    runs = []
    for g in (inits[0:2], inits[2:3], inits[3:5]):
        car, world, _ = M.finite_horizon_env(debug=False)
        runs.append(M.MPC_ORD(world, car, g, 3, verbose=False))
    M.optimize_cmaes_lockstep(runs, [5, 6, 7], sigma0=0.05, maxfevals=18)
    state = [([(w.tolist(), float(v)) for w, v in r.history], np.asarray(r.world.cars[0].state).tolist()) for r in runs]
    with open(os.path.join(out_dir, f"lock{rank}.pkl"), "wb") as f:
        pickle.dump(state, f)
    dist.barrier()
    dist.destroy_process_group()


class _OracleHostContext:
    """Stands in for the engine's host context (`ocd_episode_batch_host`) on CPU: same SoA call, oracle episodes."""

    def __init__(self, O, spec):
        self.O, self.spec = O, spec

    def episodes_soa(self, p, sc, robot_init, plan_weights, true_weights, T, weight_idx=None, unlucky_idx=None,
                     final_world=False, **_):
        ri = np.ascontiguousarray(np.asarray(robot_init, np.float32).T)
        W = np.asarray(plan_weights, np.float32).T[np.asarray(weight_idx)]
        ret = self.O.episode_batch(self.spec.params, self.spec.scenario, ri, W, np.asarray(true_weights, np.float32), T,
                                   nthreads=1)
        fw = np.zeros((p.C, 4, ri.shape[0]), np.float32)
        fw[0] = ri.T                          # tag: the episode's own initial state
        return ret, fw


def _sharded_runs_worker(rank, ws, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    import pickle
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import oracle as O
    import l4dc_mpc_ocd_b200.runtime as RT
    from l4dc_mpc_ocd_b200.interact_drive.reward_design import mpc_ord as M
    spec = O.scenario_params("finite_horizon")
    M.get_engine = lambda device=None: _OracleEngine(O, spec, 0)
    RT.get_host_context = lambda device=None: _OracleHostContext(O, spec)
    _, _, inits = M.finite_horizon_env(env_seeds=[1000000 + i for i in range(5)], debug=False)

    def fresh():
        runs = []
        for g in (inits[0:2], inits[2:3], inits[3:5]):
            car, world, _ = M.finite_horizon_env(debug=False)
            runs.append(M.MPC_ORD(world, car, g, 3, verbose=False))
        return runs

    def dump(runs, xs):
        return [([(w.tolist(), float(v)) for w, v in r.history], np.asarray(r.world.cars[0].state).tolist(),
                 np.asarray(r.car.weights).tolist(), r.iter, r.done, np.asarray(x).tolist()) for r, x in zip(runs, xs)]

    by_runs = fresh()
    xs = M.optimize_cmaes_lockstep(by_runs, [5, 6, 7], sigma0=0.05, shard_runs=True, maxfevals=18)
    by_episodes = fresh()
    xe = M.optimize_cmaes_lockstep(by_episodes, [5, 6, 7], sigma0=0.05, maxfevals=18)
    with open(os.path.join(out_dir, f"shard{rank}.pkl"), "wb") as f:
        pickle.dump((dump(by_runs, xs), dump(by_episodes, xe)), f)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_lockstep_with_sharded_runs(tmp_path):
    """shard_runs=True: rank k optimises runs k, k + 2, ... alone (no collective during the optimisation) and the ranks
    exchange histories, results and object state at the end -- both ranks finish with every run's books, and they are the
    books the episode-sharded lock step (every rank keeps all runs) produces."""
    import pickle
    ws = 2
    mp.spawn(_sharded_runs_worker, args=(ws, _free_port(), str(tmp_path)), nprocs=ws, join=True)
    a = pickle.load(open(tmp_path / "shard0.pkl", "rb"))
    b = pickle.load(open(tmp_path / "shard1.pkl", "rb"))
    assert a == b
    by_runs, by_episodes = a
    assert by_runs == by_episodes
    assert [len(h) for h, *_ in by_runs] == [19, 19, 19] and all(r[4] for r in by_runs)


def test_two_rank_lockstep_cmaes_is_rank_consistent(tmp_path):
    """Three independent CMA-ES runs in lock step over two ranks: the episodes of every generation are sharded, returns
    AND final worlds all-gathered, so both ranks hold the same histories and leave the same state in the Python
    objects (the state of each run's LAST episode, wherever it was computed)."""
    import pickle
    ws = 2
    mp.spawn(_lockstep_worker, args=(ws, _free_port(), str(tmp_path)), nprocs=ws, join=True)
    a = pickle.load(open(tmp_path / "lock0.pkl", "rb"))
    b = pickle.load(open(tmp_path / "lock1.pkl", "rb"))
    assert a == b
    assert [len(h) for h, _ in a] == [19, 19, 19]
    for (hist, state), n_inits in zip(a, (2, 1, 2)):
        assert all(np.isfinite(v) for _, v in hist)
