"""A small (mu/mu_w, lambda)-CMA-ES with pycma's default strategy parameters, ask/tell style.

The reference calls `cma.evolution_strategy.fmin2(eval_weights, x0, sigma0, {'seed': seed})`
(interact_drive/reward_design/mpc_ord.py:41) with an un-pinned pycma that is not installable here,
so this is a restatement of the published algorithm (Hansen, "The CMA Evolution Strategy: A Tutorial",
2016: default lambda = 4 + floor(3 ln N), mu = lambda/2 with log weights, c_sigma, d_sigma, c_c, c_1,
c_mu as in its Table 1), NOT a bit-for-bit pycma: parity of the sampled candidates is unpinned.  What
matters for the engine is the interface: `ask()` returns the whole population so that MPC_ORD can
evaluate it in ONE launch."""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np


class CMAES:
    def __init__(self, x0: Sequence[float], sigma0: float, seed: Optional[int] = None, popsize: Optional[int] = None):
        self.N = N = len(x0)
        self.mean = np.asarray(x0, dtype=np.float64).copy()
        self.sigma = float(sigma0)
        self.rng = np.random.RandomState(None if seed is None else int(seed) % (2 ** 32))
        self.lam = popsize or 4 + int(3 * math.log(N))
        self.mu = self.lam // 2
        w = math.log(self.lam / 2 + 0.5) - np.log(np.arange(1, self.mu + 1))
        self.weights = w / w.sum()
        self.mueff = 1.0 / np.sum(self.weights ** 2)
        self.cc = (4 + self.mueff / N) / (N + 4 + 2 * self.mueff / N)
        self.cs = (self.mueff + 2) / (N + self.mueff + 5)
        self.c1 = 2 / ((N + 1.3) ** 2 + self.mueff)
        self.cmu = min(1 - self.c1, 2 * (self.mueff - 2 + 1 / self.mueff) / ((N + 2) ** 2 + self.mueff))
        self.damps = 1 + 2 * max(0.0, math.sqrt((self.mueff - 1) / (N + 1)) - 1) + self.cs
        self.chiN = math.sqrt(N) * (1 - 1 / (4 * N) + 1 / (21 * N * N))
        self.pc, self.ps = np.zeros(N), np.zeros(N)
        self.C = np.eye(N)
        self.B, self.D = np.eye(N), np.ones(N)
        self.invsqrtC = np.eye(N)
        self.countiter = 0
        self.counteval = 0
        self._eigeneval = 0
        self._pop = None
        self.best_x, self.best_f = self.mean.copy(), np.inf
        self.fit_history = []

    def ask(self) -> np.ndarray:
        """-> population [lambda, N]."""
        z = self.rng.standard_normal((self.lam, self.N))
        self._pop = self.mean + self.sigma * (z * self.D) @ self.B.T
        return self._pop.copy()

    def tell(self, fitness: Sequence[float]) -> None:
        f = np.asarray(fitness, dtype=np.float64)
        N, pop = self.N, self._pop
        self.counteval += len(f)
        self.countiter += 1
        order = np.argsort(f, kind="stable")
        if f[order[0]] < self.best_f:
            self.best_f, self.best_x = float(f[order[0]]), pop[order[0]].copy()
        self.fit_history.append(float(f[order[0]]))
        sel = pop[order[: self.mu]]
        old = self.mean
        self.mean = self.weights @ sel
        y = (self.mean - old) / self.sigma
        self.ps = (1 - self.cs) * self.ps + math.sqrt(self.cs * (2 - self.cs) * self.mueff) * (self.invsqrtC @ y)
        hsig = (np.linalg.norm(self.ps) / math.sqrt(1 - (1 - self.cs) ** (2 * self.countiter)) / self.chiN
                < 1.4 + 2 / (N + 1))
        self.pc = (1 - self.cc) * self.pc + hsig * math.sqrt(self.cc * (2 - self.cc) * self.mueff) * y
        art = (sel - old) / self.sigma
        self.C = ((1 - self.c1 - self.cmu) * self.C
                  + self.c1 * (np.outer(self.pc, self.pc) + (1 - hsig) * self.cc * (2 - self.cc) * self.C)
                  + self.cmu * (art.T * self.weights) @ art)
        self.sigma *= math.exp((self.cs / self.damps) * (np.linalg.norm(self.ps) / self.chiN - 1))
        if self.counteval - self._eigeneval > self.lam / (self.c1 + self.cmu) / N / 10:
            self._eigeneval = self.counteval
            self.C = np.triu(self.C) + np.triu(self.C, 1).T
            d, self.B = np.linalg.eigh(self.C)
            self.D = np.sqrt(np.maximum(d, 1e-20))
            self.invsqrtC = (self.B / self.D) @ self.B.T

    def stop(self, maxfevals=np.inf, maxiter=None, tolfun=1e-11, tolx=1e-11) -> Optional[str]:
        """pycma's main default termination criteria."""
        if maxiter is None:
            maxiter = 100 + 150 * (self.N + 3) ** 2 // math.sqrt(self.lam)
        if self.counteval >= maxfevals:
            return "maxfevals"
        if self.countiter >= maxiter:
            return "maxiter"
        h = self.fit_history
        k = 10 + int(30 * self.N / self.lam)
        if len(h) >= k and max(h[-k:]) - min(h[-k:]) < tolfun:
            return "tolfun"
        if self.sigma * max(np.max(np.abs(self.pc)), math.sqrt(np.max(np.diag(self.C)))) < tolx:
            return "tolx"
        if self.D.max() > 1e7 * self.D.min():
            return "conditioncov"
        return None


def fmin2(objective: Callable, x0, sigma0, options: Optional[dict] = None, batch_objective: Optional[Callable] = None):
    """Minimise like `cma.fmin2`: returns (xbest, es).  `batch_objective(pop [lambda, N]) -> [lambda]`
    evaluates a whole generation at once when given; otherwise `objective(x)` is called per candidate."""
    options = dict(options or {})
    es = CMAES(x0, sigma0, seed=options.get("seed"), popsize=options.get("popsize"))
    stop_kw = {k: options[k] for k in ("maxfevals", "maxiter", "tolfun", "tolx") if k in options}
    while es.stop(**stop_kw) is None:
        pop = es.ask()
        fit = batch_objective(pop) if batch_objective is not None else [objective(list(x)) for x in pop]
        es.tell(fit)
    return es.best_x, es


def fmin2_lockstep(batch_objective_multi: Callable, x0s, sigma0, options_list):
    """R independent `fmin2` runs advanced in lock step.  `batch_objective_multi(pops)` gets one population per run
    (an empty list for a run that has stopped) and returns one fitness array per run -- so that all runs' generations
    can be evaluated in one launch.  Each run draws from its own RandomState, asks, tells and stops exactly as it
    would alone.  -> [(xbest, es)] per run."""
    runs = []
    for x0, opt in zip(x0s, options_list):
        opt = dict(opt or {})
        es = CMAES(x0, sigma0, seed=opt.get("seed"), popsize=opt.get("popsize"))
        runs.append((es, {k: opt[k] for k in ("maxfevals", "maxiter", "tolfun", "tolx") if k in opt}))
    active = [i for i, (es, kw) in enumerate(runs) if es.stop(**kw) is None]
    while active:
        pops = [[] for _ in runs]
        for i in active:
            pops[i] = runs[i][0].ask()
        fits = batch_objective_multi(pops)
        for i in active:
            runs[i][0].tell(fits[i])
        active = [i for i in active if runs[i][0].stop(**runs[i][1]) is None]
    return [(es.best_x, es) for es, _ in runs]
