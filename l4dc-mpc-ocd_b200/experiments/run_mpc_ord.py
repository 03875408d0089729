"""Cost-design driver: mirror of experiments/run_mpc_ord.py:19-127 of the reference (same CLI).

    python -m l4dc_mpc_ocd_b200.experiments.run_mpc_ord finite_horizon cmaes --n_inits 5

Differences that follow from the engine: no multiprocessing.Pool (one process drives one GPU and a
whole CMA-ES generation is one kernel launch; the per-init CMA-ES runs of `--one_by_one`, one worker
process each in the reference, advance in lock step and share one launch per generation), `vis` prints the returns of the true and the tuned weights instead of rendering,
and `--max_evals` bounds the CMA-ES run (the reference runs until pycma's own termination).  Under
torchrun (WORLD_SIZE > 1) the episodes of every generation are sharded over the ranks and the
per-episode returns all-gathered (see l4dc_mpc_ocd_b200.parallel)."""
from __future__ import annotations

from argparse import ArgumentParser

import numpy as np

from ..interact_drive.reward_design.mpc_ord import MPC_ORD, finite_horizon_env, optimize_cmaes_lockstep
from .local_opt_scenario import local_opt_env
from .replanning_world import setup_world as replanning_env


def fmt(arr):
    s = str(arr).replace("\n", ' ').replace('\t', " ")
    while '  ' in s:
        s = s.replace('  ', ' ')
    return s


envs = {
    'local_opt': {
        'make_env': local_opt_env, 'eval_horizon': 15,
        'init_offset_range': [[0., -0.1, 0., 0.], [0., 0.1, 0., 0.]], 'num_eval_samples': 1,
        'tuned_weights': np.array([-0.09686739, 0.25720383, -0.58355971, -0.23075428, -0.41237239, -0.4758984,
                                   -0.36625558]),
    },
    'finite_horizon': {
        'make_env': finite_horizon_env, 'eval_horizon': 15,
        'init_offset_range': [[-0.1, 0., 0., 0.], [0.1, 0., 0., 0.]], 'num_eval_samples': 1,
        'tuned_weights': np.array([-0.21963165, -0.01184596, 0.34379187, -0.04687411, -0.06364365, -0.54138792,
                                   -0.7308079]),
    },
    'replanning': {
        'make_env': replanning_env, 'eval_horizon': 20,
        'init_offset_range': [[-0.05, 0., 0., 0.], [0.05, 0., 0., 0.]],
        'num_eval_samples': 2,    # both outcomes are sampled to get the expected return
        'tuned_weights': np.array([-0.55899817, -0.4436692, -0.3724511, -0.19964276, -0.5438697, 0.12770043]),
    },
}


def make_ord(env, init_states, args, optimization_seed, verbose=True):
    """The MPC_ORD of one optimisation over `init_states` (first half of the reference's run_opt, :92-104)."""
    if verbose:
        print("OPTIMIZING REWARD FROM INIT STATES", init_states)
    car, world, _ = env['make_env'](debug=True)
    tag = args.n_inits if not args.one_by_one else fmt(init_states[0])
    save_path = (f'{args.optimizer}_{args.scenario}__designer_weights_{fmt(car.weights)}__{tag}_init_seed_{args.seed}'
                 f'_opt_seed_{optimization_seed}_sigma_{args.sigma}.pkl') if not args.no_save else None
    return MPC_ORD(world, car, init_states, env['eval_horizon'], num_samples=env['num_eval_samples'],
                   save_path=save_path, verbose=verbose), car


def run_opt(env, init_states, args, optimization_seed, verbose=True):
    """One optimisation over `init_states` (reference run_opt, :92-122).  Returns the MPC_ORD."""
    mpc_ord, car = make_ord(env, init_states, args, optimization_seed, verbose)
    if args.optimizer == 'random':
        mpc_ord.optimize_random_search(n_iter=400, seed=optimization_seed)
    elif args.optimizer == 'vis':
        true_w = car.weights
        r_true = -mpc_ord.eval_weights(true_w)
        r_tuned = -mpc_ord.eval_weights(env["tuned_weights"])
        print(f'{args.scenario}: return of the true weights {fmt(true_w)} = {r_true}')
        print(f'{args.scenario}: return of the tuned weights {fmt(env["tuned_weights"])} = {r_tuned}')
    else:
        assert args.optimizer == 'cmaes'
        stop = {} if args.max_evals is None else dict(maxfevals=args.max_evals)
        mpc_ord.optimize_cmaes(sigma0=args.sigma, seed=optimization_seed, **stop)
    return mpc_ord


def run_opts_lockstep(env, groups, args, optimization_seed, verbose=True):
    """`--one_by_one` with CMA-ES: the reference gives every init group its own worker process
    (`Pool(len(init_states_groups))`, :83-90); here the groups' CMA-ES runs advance in lock step and generation g of
    all of them is one episode launch (sharded over the ranks under torchrun).  Returns the MPC_ORDs."""
    ords = [make_ord(env, g, args, optimization_seed, verbose)[0] for g in groups]
    stop = {} if args.max_evals is None else dict(maxfevals=args.max_evals)
    optimize_cmaes_lockstep(ords, [optimization_seed] * len(ords), sigma0=args.sigma, **stop)
    return ords


def _init_distributed():
    """Under torchrun (WORLD_SIZE > 1): one process per GPU, NCCL process group.  -> rank or None."""
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return None
    import torch
    import torch.distributed as dist
    from ..runtime import set_default_device
    from ..parallel import bind_rank_cpus
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind_rank_cpus(local, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
    torch.cuda.set_device(local)
    set_default_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist.get_rank()


def main(argv=None):
    parser = ArgumentParser()
    parser.add_argument('scenario', type=str, choices=['local_opt', 'finite_horizon', 'replanning'],
                        help='Which scenario to run reward weight optimization for.')
    parser.add_argument('optimizer', type=str, choices=['random', 'cmaes', 'vis'],
                        help='Algorithm used for weight optimization.')
    parser.add_argument('--n_inits', type=int, default=1)
    parser.add_argument('--seed', type=int, default=None)
    parser.add_argument('--one_by_one', action='store_true',
                        help='Runs single init optimization separately for each init.')
    parser.add_argument('--rand_inits', action='store_true')
    parser.add_argument('--sigma', type=float, default=0.05)
    parser.add_argument('--max_evals', type=int, default=None, help='bound on CMA-ES evaluations (engine extension)')
    parser.add_argument('--opt_seed', type=int, default=None, help='fix the optimiser seed (engine extension)')
    parser.add_argument('--no_save', action='store_true', help='do not pickle the history (engine extension)')
    parser.add_argument('--quiet', action='store_true')
    args = parser.parse_args(argv)

    assert args.n_inits >= 1
    assert args.seed != 0, 'CMA doesn\'t accept 0 seed'
    if args.n_inits == 1:
        args.one_by_one = False
    env = envs[args.scenario]
    optimization_seed = np.random.randint(0, 2 ** 32) if args.opt_seed is None else args.opt_seed
    rank = _init_distributed()
    if rank is not None:
        # every rank runs the same optimiser on the same (all-gathered) returns: share the seed, and
        # let rank 0 alone talk and write the history
        import torch
        import torch.distributed as dist
        seed_t = torch.tensor([optimization_seed], dtype=torch.int64, device="cuda")
        dist.broadcast(seed_t, 0)
        optimization_seed = int(seed_t.item())
        if rank != 0:
            args.quiet, args.no_save = True, True
    if args.seed is None:
        args.seed = optimization_seed
    env_seeds = [(args.seed * 1000000 + i) % (2 ** 32) for i in range(args.n_inits)]
    car, world, init_states = env['make_env'](env_seeds=env_seeds, debug=True)
    groups = [[s] for s in init_states] if args.one_by_one else [init_states]
    if not args.quiet:
        print('init_states:', groups)
    if args.optimizer == 'cmaes' and len(groups) > 1:
        return run_opts_lockstep(env, groups, args, optimization_seed, verbose=not args.quiet)
    return [run_opt(env, g, args, optimization_seed, verbose=not args.quiet) for g in groups]


if __name__ == '__main__':
    main()
