"""A small (mu/mu_w, lambda)-CMA-ES with pycma's default strategy parameters, ask/tell style.

The reference calls `cma.evolution_strategy.fmin2(eval_weights, x0, sigma0, {'seed': seed})`
(interact_drive/reward_design/mpc_ord.py:41) with an un-pinned pycma that is not installable here,
so this is a restatement of the published algorithm (Hansen, "The CMA Evolution Strategy: A Tutorial",
2016: default lambda = 4 + floor(3 ln N), mu = lambda/2 with log weights, c_sigma, d_sigma, c_c, c_1,
c_mu as in its Table 1), NOT a bit-for-bit pycma: parity of the sampled candidates is unpinned.  What
matters for the engine is the interface: `ask()` returns the whole population so that MPC_ORD can
evaluate it in ONE launch.

The state lives in `CMAESBatch`: R independent runs of the same dimension and population size as stacked arrays
([R, N], [R, N, N] ...), every update one numpy call for all of them -- the reference runs such optimisations in
separate worker processes (experiments/run_mpc_ord.py:83-90); in lock step on one GPU their Python bookkeeping,
not the episode launch, is what a generation costs, so it must not be a loop over runs.  Every operation is
row-wise (stacked matmul / eigh, reductions along the last axes, one RandomState per run), so a run's trajectory
does not depend on which other runs share the batch, and it is bit for bit the trajectory of `CMAES`, the plain
single-run class."""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np


def _row_norm(x: np.ndarray) -> np.ndarray:
    """np.linalg.norm of every row: sqrt(x.dot(x)) -- vecdot runs the same BLAS dot over each row."""
    return np.sqrt(np.vecdot(x, x))


class CMAES:
    """One run, plain 2-D numpy: what `optimize_cmaes` drives.  `CMAESBatch` below performs the same operations on stacked
    arrays and reproduces this class bit for bit (tests/test_host_logic_cpu.py pins that on long runs; numpy applies the
    same BLAS / LAPACK routine to every matrix of a stack)."""

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


class CMAESBatch:
    """R independent CMA-ES runs with a common dimension N and population size, advanced together.
    `ask(idx)` / `tell(idx, fitness)` / `stop(idx, ...)` act on the runs listed in `idx` (default: all)."""

    def __init__(self, x0s, sigma0, seeds: Sequence[Optional[int]], popsize: Optional[int] = None):
        self.mean = np.array(x0s, dtype=np.float64, ndmin=2)
        self.R, self.N = R, N = self.mean.shape
        sig = np.broadcast_to(np.asarray(sigma0, dtype=np.float64), (R,))
        self.sigma = sig.copy()
        assert len(seeds) == R
        self.rngs = [np.random.RandomState(None if s is None else int(s) % (2 ** 32)) for s in seeds]
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
        self.pc, self.ps = np.zeros((R, N)), np.zeros((R, N))
        eye = np.broadcast_to(np.eye(N), (R, N, N))
        self.C, self.B, self.invsqrtC = eye.copy(), eye.copy(), eye.copy()
        self.D = np.ones((R, N))
        self.countiter = np.zeros(R, dtype=np.int64)
        self.counteval = np.zeros(R, dtype=np.int64)
        self._eigeneval = np.zeros(R, dtype=np.int64)
        self._pop = np.zeros((R, self.lam, N))
        self.best_x, self.best_f = self.mean.copy(), np.full(R, np.inf)
        self.fit_history = [[] for _ in range(R)]
        self._all = np.arange(R)

    def _idx(self, idx) -> np.ndarray:
        return self._all if idx is None else np.asarray(idx, dtype=np.int64)

    def ask(self, idx=None) -> np.ndarray:
        """-> populations [len(idx), lambda, N]."""
        idx = self._idx(idx)
        z = np.stack([self.rngs[r].standard_normal((self.lam, self.N)) for r in idx])
        y = (self.sigma[idx, None, None] * (z * self.D[idx, None, :])) @ self.B[idx].transpose(0, 2, 1)
        pop = self.mean[idx, None, :] + y
        self._pop[idx] = pop
        return pop.copy()

    def tell(self, fitness, idx=None) -> None:
        """fitness [len(idx), lambda] of the populations last asked for."""
        idx = self._idx(idx)
        f = np.asarray(fitness, dtype=np.float64).reshape(len(idx), self.lam)
        N, pop = self.N, self._pop[idx]
        rows = np.arange(len(idx))
        self.counteval[idx] += self.lam
        self.countiter[idx] += 1
        order = np.argsort(f, axis=1, kind="stable")
        fbest = f[rows, order[:, 0]]
        better = fbest < self.best_f[idx]
        if better.any():
            bi = idx[better]
            self.best_f[bi] = fbest[better]
            self.best_x[bi] = pop[rows[better], order[better, 0]]
        for r, v in zip(idx, fbest):
            self.fit_history[r].append(float(v))
        sel = np.take_along_axis(pop, order[:, : self.mu, None], axis=1)           # [A, mu, N]
        old, sigma = self.mean[idx], self.sigma[idx]
        mean = np.matmul(self.weights, sel)                                          # [A, N]
        y = (mean - old) / sigma[:, None]
        ps = (1 - self.cs) * self.ps[idx] + math.sqrt(self.cs * (2 - self.cs) * self.mueff) * \
            np.matmul(self.invsqrtC[idx], y[:, :, None])[:, :, 0]
        psn = _row_norm(ps)
        damp = np.array([math.sqrt(1 - (1 - self.cs) ** (2 * int(c))) for c in self.countiter[idx]])
        hsig = (psn / damp / self.chiN < 1.4 + 2 / (N + 1)).astype(np.float64)
        pc = (1 - self.cc) * self.pc[idx] + (hsig * math.sqrt(self.cc * (2 - self.cc) * self.mueff))[:, None] * y
        art = (sel - old[:, None, :]) / sigma[:, None, None]
        Cold = self.C[idx]
        C = ((1 - self.c1 - self.cmu) * Cold
             + self.c1 * (pc[:, :, None] * pc[:, None, :] + ((1 - hsig) * self.cc * (2 - self.cc))[:, None, None] * Cold)
             + np.matmul(self.cmu * (art.transpose(0, 2, 1) * self.weights), art))
        self.mean[idx], self.ps[idx], self.pc[idx], self.C[idx] = mean, ps, pc, C
        self.sigma[idx] = sigma * np.array([math.exp(v) for v in (self.cs / self.damps) * (psn / self.chiN - 1)])
        due = self.counteval[idx] - self._eigeneval[idx] > self.lam / (self.c1 + self.cmu) / N / 10
        if due.any():
            di = idx[due]
            self._eigeneval[di] = self.counteval[di]
            Cd = np.triu(C[due]) + np.triu(C[due], 1).transpose(0, 2, 1)
            d, B = np.linalg.eigh(Cd)
            D = np.sqrt(np.maximum(d, 1e-20))
            self.C[di], self.B[di], self.D[di] = Cd, B, D
            self.invsqrtC[di] = np.matmul(B / D[:, None, :], B.transpose(0, 2, 1))

    def stop(self, idx=None, maxfevals=np.inf, maxiter=None, tolfun=1e-11, tolx=1e-11) -> list:
        """pycma's main default termination criteria, per run: a reason string or None.  The four limits may be
        scalars or one value per run in `idx`."""
        idx = self._idx(idx)
        A = len(idx)
        if maxiter is None:
            maxiter = 100 + 150 * (self.N + 3) ** 2 // math.sqrt(self.lam)
        per = lambda v: np.broadcast_to(np.asarray(v, dtype=np.float64), (A,))
        maxfevals, maxiter, tolfun, tolx = per(maxfevals), per(maxiter), per(tolfun), per(tolx)
        k = 10 + int(30 * self.N / self.lam)
        spread = np.full(A, np.inf)
        lens = [len(self.fit_history[r]) for r in idx]
        if A and min(lens) == max(lens):                # runs in lock step have equally long histories: one reduction
            if lens[0] >= k:
                win = np.array([self.fit_history[r][-k:] for r in idx])
                spread = win.max(axis=1) - win.min(axis=1)
        else:
            for a, r in enumerate(idx):
                h = self.fit_history[r]
                if len(h) >= k:
                    spread[a] = max(h[-k:]) - min(h[-k:])
        C = self.C[idx]
        size = self.sigma[idx] * np.maximum(np.abs(self.pc[idx]).max(axis=1),
                                            np.sqrt(np.diagonal(C, axis1=1, axis2=2).max(axis=1)))
        D = self.D[idx]
        # the first criterion that holds names the reason (same order as the single-run chain of ifs)
        tests = (("maxfevals", self.counteval[idx] >= maxfevals), ("maxiter", self.countiter[idx] >= maxiter),
                 ("tolfun", spread < tolfun), ("tolx", size < tolx), ("conditioncov", D.max(axis=1) > 1e7 * D.min(axis=1)))
        out = [None] * A
        hit = np.zeros(A, dtype=bool)
        for name, cond in tests:
            for a in np.nonzero(cond & ~hit)[0]:
                out[a] = name
            hit |= cond
        return out


class _Run:
    """One run of a CMAESBatch seen through the single-run attribute names."""

    def __init__(self, batch: CMAESBatch, r: int):
        self._b, self._r = batch, r

    N = property(lambda s: s._b.N)
    lam = property(lambda s: s._b.lam)
    mu = property(lambda s: s._b.mu)
    weights = property(lambda s: s._b.weights)
    mueff = property(lambda s: s._b.mueff)
    cc = property(lambda s: s._b.cc)
    cs = property(lambda s: s._b.cs)
    c1 = property(lambda s: s._b.c1)
    cmu = property(lambda s: s._b.cmu)
    damps = property(lambda s: s._b.damps)
    chiN = property(lambda s: s._b.chiN)
    mean = property(lambda s: s._b.mean[s._r])
    sigma = property(lambda s: float(s._b.sigma[s._r]))
    pc = property(lambda s: s._b.pc[s._r])
    ps = property(lambda s: s._b.ps[s._r])
    C = property(lambda s: s._b.C[s._r])
    B = property(lambda s: s._b.B[s._r])
    D = property(lambda s: s._b.D[s._r])
    countiter = property(lambda s: int(s._b.countiter[s._r]))
    counteval = property(lambda s: int(s._b.counteval[s._r]))
    best_x = property(lambda s: s._b.best_x[s._r].copy())
    best_f = property(lambda s: float(s._b.best_f[s._r]))
    fit_history = property(lambda s: s._b.fit_history[s._r])


_STOP_KEYS = ("maxfevals", "maxiter", "tolfun", "tolx")


def fmin2(objective: Callable, x0, sigma0, options: Optional[dict] = None, batch_objective: Optional[Callable] = None):
    """Minimise like `cma.fmin2`: returns (xbest, es).  `batch_objective(pop [lambda, N]) -> [lambda]`
    evaluates a whole generation at once when given; otherwise `objective(x)` is called per candidate."""
    options = dict(options or {})
    es = CMAES(x0, sigma0, seed=options.get("seed"), popsize=options.get("popsize"))
    stop_kw = {k: options[k] for k in _STOP_KEYS if k in options}
    while es.stop(**stop_kw) is None:
        pop = es.ask()
        fit = batch_objective(pop) if batch_objective is not None else [objective(list(x)) for x in pop]
        es.tell(fit)
    return es.best_x, es


def fmin2_lockstep(batch_objective_multi: Callable, x0s, sigma0, options_list):
    """R independent `fmin2` runs advanced in lock step.  `batch_objective_multi(pops)` gets one population per run
    (an empty list for a run that has stopped) and returns one fitness array per run -- so that all runs' generations
    can be evaluated in one launch.  Each run draws from its own RandomState, asks, tells and stops exactly as it
    would alone; runs of the same dimension and population size share one CMAESBatch, so a generation's bookkeeping
    is one set of numpy calls for all of them.  -> [(xbest, es)] per run."""
    opts = [dict(o or {}) for o in options_list]
    R = len(opts)
    groups = {}
    for r, (x0, o) in enumerate(zip(x0s, opts)):
        groups.setdefault((len(x0), o.get("popsize")), []).append(r)
    batches, views = [], [None] * R
    defaults = dict(maxfevals=np.inf, maxiter=None, tolfun=1e-11, tolx=1e-11)
    for (N, popsize), members in groups.items():
        b = CMAESBatch([list(x0s[r]) for r in members], float(sigma0), [opts[r].get("seed") for r in members], popsize)
        dflt_iter = 100 + 150 * (N + 3) ** 2 // math.sqrt(b.lam)
        lim = {k: np.array([float(opts[r].get(k, dflt_iter if k == "maxiter" else defaults[k])) for r in members])
               for k in _STOP_KEYS}
        batches.append((b, np.asarray(members), lim))
        for j, r in enumerate(members):
            views[r] = _Run(b, j)

    def still_going(b, members, lim, local):
        reasons = b.stop(local, **{k: v[local] for k, v in lim.items()})
        return local[[x is None for x in reasons]]

    active = [still_going(b, m, lim, np.arange(len(m))) for b, m, lim in batches]
    while any(len(a) for a in active):
        pops = [[] for _ in range(R)]
        for (b, members, _), local in zip(batches, active):
            if len(local):
                for r, p in zip(members[local], b.ask(local)):
                    pops[r] = p
        fits = batch_objective_multi(pops)
        for (b, members, _), local in zip(batches, active):
            if len(local):
                b.tell(np.stack([np.asarray(fits[r], dtype=np.float64) for r in members[local]]), local)
        active = [still_going(b, m, lim, local) if len(local) else local for (b, m, lim), local in zip(batches, active)]
    return [(v.best_x, v) for v in views]
