"""ctypes front-end of the CPU oracle (``libocd_oracle.so``).  TEST INFRASTRUCTURE ONLY.

Reference lines restated by each C function are cited in ``ocd_oracle_impl.inc``.  The
scenario constants below restate (paths relative to /root/reference)
``interact_drive/reward_design/mpc_ord.py:162-207`` (finite_horizon),
``experiments/local_opt_scenario.py:6-54`` (local_opt) and
``experiments/replanning_world.py:38-95`` (replanning), plus ``experiments/run_mpc_ord.py:19-44``
(eval horizon / samples / tuned weights).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libocd_oracle.so"

MAX_LANES, MAX_CARS, MAX_PLAN, MAX_H, MAX_S = 4, 8, 16, 64, 6


class _Params(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("C", C.c_int32), ("L", C.c_int32), ("n_iter", C.c_int32),
        ("num_lanes", C.c_int32), ("other_mode", C.c_int32), ("extra_inits", C.c_int32),
        ("optimizer", C.c_int32),
        ("lr", C.c_double), ("dt", C.c_double), ("friction", C.c_double),
        ("target_speed", C.c_double), ("lane_x", C.c_double * MAX_LANES),
    ]


class _Scenario(C.Structure):
    _fields_ = [
        ("n_other", C.c_int32), ("critical_t", C.c_int32),
        ("kind", C.c_int32 * MAX_CARS), ("plan_len", C.c_int32 * MAX_CARS),
        ("init_state", (C.c_double * 4) * MAX_CARS),
        ("friction", C.c_double * MAX_CARS),
        ("control", (C.c_double * 2) * MAX_CARS),
        ("plan", ((C.c_double * 2) * MAX_PLAN) * MAX_CARS),
        ("teleport_state", C.c_double * 4),
    ]


@dataclass
class OracleParams:
    H: int = 5
    C: int = 2
    lane_x: Sequence[float] = (-0.1, 0.0, 0.1)
    n_iter: int = 100
    num_lanes: int = 3
    other_mode: int = 0
    extra_inits: bool = False
    lr: float = 0.1
    dt: float = 0.1
    friction: float = 0.2
    target_speed: float = 1.0
    optimizer: int = 0            # 1: the engine's opt-in L-BFGS (no running reference counterpart)

    @property
    def L(self) -> int:
        return len(self.lane_x)

    @property
    def K(self) -> int:
        return self.L + 4

    @property
    def S(self) -> int:
        return 6 if self.extra_inits else 3

    def c_struct(self) -> _Params:
        p = _Params()
        p.H, p.C, p.L, p.n_iter = self.H, self.C, self.L, self.n_iter
        p.num_lanes, p.other_mode, p.extra_inits = self.num_lanes, self.other_mode, int(self.extra_inits)
        p.optimizer = int(self.optimizer)
        p.lr, p.dt, p.friction, p.target_speed = self.lr, self.dt, self.friction, self.target_speed
        for i, x in enumerate(self.lane_x):
            p.lane_x[i] = float(x)
        return p


@dataclass
class OracleScenario:
    """Scripted cars (cars 1..C-1) and the replanning teleport."""
    init_state: Sequence[Sequence[float]] = ()
    kind: Sequence[int] = ()                # 0 fixed control / fixed velocity, 1 fixed plan
    friction: Sequence[float] = ()
    control: Sequence[Sequence[float]] = ()  # fixed / default control
    plan: Sequence[Sequence[Sequence[float]]] = ()
    critical_t: int = 0
    teleport_state: Sequence[float] = (10.0, 0.0, 0.0, 0.0)

    def c_struct(self) -> _Scenario:
        s = _Scenario()
        n = len(self.init_state)
        s.n_other, s.critical_t = n, self.critical_t
        for j in range(n):
            s.kind[j] = int(self.kind[j])
            s.friction[j] = float(self.friction[j])
            for c in range(4):
                s.init_state[j][c] = float(self.init_state[j][c])
            for c in range(2):
                s.control[j][c] = float(self.control[j][c])
            pl = self.plan[j] if j < len(self.plan) else ()
            s.plan_len[j] = len(pl)
            for t, u in enumerate(pl):
                s.plan[j][t][0], s.plan[j][t][1] = float(u[0]), float(u[1])
        for c in range(4):
            s.teleport_state[c] = float(self.teleport_state[c])
        return s


# --------------------------------------------------------------------------------------
_lib: Optional[C.CDLL] = None


def build(force: bool = False) -> Path:
    """Compile the oracle with the system gcc (``make -C oracle``)."""
    srcs = [_HERE / "ocd_oracle.c", _HERE / "ocd_oracle_impl.inc", _HERE / "ocd_oracle.h"]
    stale = (not _SO.exists()) or any(s.stat().st_mtime > _SO.stat().st_mtime for s in srcs)
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(_HERE)], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.ocdo_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().ocdo_max_threads())


def _suffix(dtype) -> str:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(x, dtype, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(x, dtype=dtype))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _real(dtype):
    return C.c_float if np.dtype(dtype) == np.float32 else C.c_double


# ---- primitives ------------------------------------------------------------------------
def dynamics_step(state, control, dt: float, friction: float, dtype=np.float32) -> np.ndarray:
    s, u = _arr(state, dtype, (4,)), _arr(control, dtype, (2,))
    out = np.empty(4, dtype)
    fn = getattr(lib(), f"ocdo_dynamics_step_{_suffix(dtype)}")
    fn.restype = None
    fn(_ptr(s), _ptr(u), C.c_double(dt), C.c_double(friction), _ptr(out))
    return out


def smooth_f(x: float, shape: float = 5.0, dtype=np.float32) -> float:
    fn = getattr(lib(), f"ocdo_f_{_suffix(dtype)}")
    fn.restype = _real(dtype)
    fn.argtypes = [_real(dtype), _real(dtype)]
    return float(fn(x, shape))


def smooth_threshold(z: float, threshold: float, width: float, c: float = 5.0, dtype=np.float32) -> float:
    fn = getattr(lib(), f"ocdo_smooth_threshold_{_suffix(dtype)}")
    fn.restype = _real(dtype)
    fn.argtypes = [_real(dtype), C.c_double, C.c_double, C.c_double]
    return float(fn(z, threshold, width, c))


def smooth_bump(z: float, start: float, end: float, dtype=np.float32) -> float:
    fn = getattr(lib(), f"ocdo_smooth_bump_{_suffix(dtype)}")
    fn.restype = _real(dtype)
    fn.argtypes = [_real(dtype)] * 3
    return float(fn(z, start, end))


def features(p: OracleParams, world, dtype=np.float32, jac: bool = False):
    w = _arr(world, dtype, (p.C, 4))
    phi = np.empty(p.K, dtype)
    J = np.empty((p.K, 4), dtype) if jac else None
    ps = p.c_struct()
    fn = getattr(lib(), f"ocdo_features_{_suffix(dtype)}")
    fn.restype = None
    fn(C.byref(ps), _ptr(w), _ptr(phi), _ptr(J))
    return (phi, J) if jac else phi


def _other_controls(p: OracleParams, other_controls, dtype, batch=None):
    """Accepts [C-1][H][2] (cars 1..C-1) and pads the unused robot row in front."""
    if other_controls is None:
        return None
    oc = np.asarray(other_controls, dtype=dtype)
    lead = () if batch is None else (batch,)
    if oc.shape == lead + (p.C - 1, p.H, 2):
        pad = np.zeros(lead + (1, p.H, 2), dtype)
        oc = np.concatenate([pad, oc], axis=len(lead))
    return _arr(oc, dtype, lead + (p.C, p.H, 2))


def mpc_reward(p: OracleParams, init_world, controls, weights, other_controls=None,
               dtype=np.float32, grad: bool = True):
    iw, u, w = _arr(init_world, dtype, (p.C, 4)), _arr(controls, dtype, (p.H, 2)), _arr(weights, dtype, (p.K,))
    oc = _other_controls(p, other_controls, dtype)
    R = np.zeros(1, dtype)
    g = np.zeros((p.H, 2), dtype) if grad else None
    ps = p.c_struct()
    fn = getattr(lib(), f"ocdo_mpc_reward_{_suffix(dtype)}")
    fn.restype = C.c_int
    rc = fn(C.byref(ps), _ptr(iw), _ptr(u), _ptr(oc), _ptr(w), _ptr(R), _ptr(g))
    if rc:
        raise ValueError(f"oracle mpc_reward rejected the problem (rc={rc})")
    return (R[0], g) if grad else R[0]


def generate_plan(p: OracleParams, init_world, weights, other_controls=None, cur_speed=None,
                  dtype=np.float32, all_plans: bool = False):
    iw, w = _arr(init_world, dtype, (p.C, 4)), _arr(weights, dtype, (p.K,))
    oc = _other_controls(p, other_controls, dtype)
    plan, losses = np.zeros((p.H, 2), dtype), np.zeros(p.S, dtype)
    best = C.c_int32(0)
    ap = np.zeros((p.S, p.H, 2), dtype) if all_plans else None
    ps = p.c_struct()
    cs = iw[0, 2] if cur_speed is None else cur_speed
    fn = getattr(lib(), f"ocdo_generate_plan_{_suffix(dtype)}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 4 + [_real(dtype)] + [C.c_void_p] * 4
    rc = fn(C.addressof(ps), _ptr(iw), _ptr(oc), _ptr(w), float(cs), _ptr(plan), _ptr(losses),
            C.addressof(best), _ptr(ap))
    if rc:
        raise ValueError(f"oracle generate_plan rejected the problem (rc={rc})")
    out = dict(plan=plan, losses=losses, best=int(best.value))
    if all_plans:
        out["all_plans"] = ap
    return out


def generate_plan_batch(p: OracleParams, init_world, weights, other_controls=None,
                        dtype=np.float32, nthreads: Optional[int] = None):
    iw = _arr(init_world, dtype)
    B = iw.shape[0]
    iw = iw.reshape(B, p.C, 4)
    w = _arr(np.broadcast_to(np.asarray(weights, dtype), (B, p.K)), dtype)
    oc = _other_controls(p, other_controls, dtype, batch=B)
    plan, losses = np.zeros((B, p.H, 2), dtype), np.zeros((B, p.S), dtype)
    best = np.zeros(B, np.int32)
    ps = p.c_struct()
    fn = getattr(lib(), f"ocdo_generate_plan_batch_{_suffix(dtype)}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6 + [C.c_int]
    rc = fn(C.addressof(ps), B, _ptr(iw), _ptr(oc), _ptr(w), _ptr(plan), _ptr(losses), _ptr(best),
            int(nthreads or max_threads()))
    if rc:
        raise ValueError(f"oracle generate_plan_batch failed (rc={rc})")
    return dict(plan=plan, losses=losses, best=best)


def episode(p: OracleParams, sc: OracleScenario, robot_init, w_plan, w_true, T: int,
            unlucky_idx: int = 0, dtype=np.float32):
    ri, wp, wt = _arr(robot_init, dtype, (4,)), _arr(w_plan, dtype, (p.K,)), _arr(w_true, dtype, (p.K,))
    ret = np.zeros(1, dtype)
    tc, tb = np.zeros((T, 2), dtype), np.zeros(T, np.int32)
    ts, sr = np.zeros((T, p.C, 4), dtype), np.zeros(T, dtype)
    ps, ss = p.c_struct(), sc.c_struct()
    fn = getattr(lib(), f"ocdo_episode_{_suffix(dtype)}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int32] + [C.c_void_p] * 5
    rc = fn(C.addressof(ps), C.addressof(ss), _ptr(ri), _ptr(wp), _ptr(wt), int(unlucky_idx), int(T),
            _ptr(ret), _ptr(tc), _ptr(tb), _ptr(ts), _ptr(sr))
    if rc:
        raise ValueError(f"oracle episode failed (rc={rc})")
    return dict(ret=ret[0], controls=tc, best=tb, states=ts, step_rewards=sr)


def episode_batch(p: OracleParams, sc: OracleScenario, robot_init, w_plan, w_true, T: int,
                  unlucky_idx=None, dtype=np.float32, nthreads: Optional[int] = None) -> np.ndarray:
    ri = _arr(robot_init, dtype)
    B = ri.shape[0]
    ri = ri.reshape(B, 4)
    wp = _arr(np.broadcast_to(np.asarray(w_plan, dtype), (B, p.K)), dtype)
    wt = _arr(w_true, dtype, (p.K,))
    ul = None if unlucky_idx is None else _arr(unlucky_idx, np.int32, (B,))
    ret = np.zeros(B, dtype)
    ps, ss = p.c_struct(), sc.c_struct()
    fn = getattr(lib(), f"ocdo_episode_batch_{_suffix(dtype)}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p, C.c_int]
    rc = fn(C.addressof(ps), C.addressof(ss), B, _ptr(ri), _ptr(wp), _ptr(wt), _ptr(ul), int(T), _ptr(ret),
            int(nthreads or max_threads()))
    if rc:
        raise ValueError(f"oracle episode_batch failed (rc={rc})")
    return ret


# ---- scenario constants ------------------------------------------------------------------
_PI2 = float(np.pi / 2)


@dataclass
class ScenarioSpec:
    name: str
    params: OracleParams
    scenario: OracleScenario
    designer_weights: np.ndarray          # before normalisation
    tuned_weights: np.ndarray
    eval_horizon: int
    num_samples: int
    example_init: np.ndarray = field(default_factory=lambda: np.zeros(4))


def scenario_params(name: str, horizon: int = 5, extra_inits: bool = False) -> ScenarioSpec:
    if name == "finite_horizon":      # mpc_ord.py:162-207 ; run_mpc_ord.py:29-36
        return ScenarioSpec(
            name,
            OracleParams(H=horizon, C=2, lane_x=(-0.1, 0.0, 0.1), n_iter=200 if horizon == 6 else 100,
                         num_lanes=3, other_mode=0, extra_inits=extra_inits, target_speed=1.0),
            OracleScenario(init_state=[[0.0, -0.6, 0.5, _PI2]], kind=[0], friction=[0.0], control=[[0.0, 0.0]]),
            np.array([-5, 0., 0., 0., -6., -50, -50]),
            np.array([-0.21963165, -0.01184596, 0.34379187, -0.04687411, -0.06364365, -0.54138792, -0.7308079]),
            15, 1, np.array([0.084157771, -0.893098184, 0.784348475, _PI2]))
    if name == "local_opt":           # local_opt_scenario.py:6-54 ; run_mpc_ord.py:20-27
        return ScenarioSpec(
            name,
            OracleParams(H=5, C=2, lane_x=(-0.1, 0.0, 0.1), n_iter=100, num_lanes=3, other_mode=0,
                         extra_inits=extra_inits, target_speed=1.0),
            OracleScenario(init_state=[[0.0, -0.9, 1.0, _PI2]], kind=[0], friction=[0.0], control=[[0.0, 0.0]]),
            np.array([-5, 0., 0., -10, 0, -50, -50]),
            np.array([-0.09686739, 0.25720383, -0.58355971, -0.23075428, -0.41237239, -0.4758984, -0.36625558]),
            15, 1, np.array([-0.088658804, -0.886196368, 0.984348475, _PI2]))
    if name == "replanning":          # replanning_world.py:11-95 ; run_mpc_ord.py:37-43
        plan1 = [[0., 0.], [0.7, 2.7], [0., 0.], [0.0, -2.7]]
        plan2 = [[0., 0.], [0.7, -2.7], [0., 0.], [0.0, 2.7]]
        return ScenarioSpec(
            name,
            OracleParams(H=5, C=3, lane_x=(-0.05, 0.05), n_iter=100, num_lanes=2, other_mode=1,
                         extra_inits=False, target_speed=1.2),
            OracleScenario(init_state=[[0., -0.7, 0.8, _PI2]] * 2, kind=[1, 1], friction=[0.2, 0.2],
                           control=[[0., 0.]] * 2, plan=[plan1, plan2], critical_t=4),
            np.array([-3, 0, 0, -2, -10, -10], dtype=np.float32),
            np.array([-0.55899817, -0.4436692, -0.3724511, -0.19964276, -0.5438697, 0.12770043]),
            20, 2, np.array([0.004881371, -0.886196368, 0.973891383, _PI2]))
    raise KeyError(name)


SCENARIOS = ("finite_horizon", "local_opt", "replanning")
