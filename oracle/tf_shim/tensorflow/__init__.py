"""Minimal stand-in for the handful of TensorFlow 2.1 primitives the reference's hot path calls.

TEST INFRASTRUCTURE ONLY (lives under oracle/).  Purpose: let the reference's UNMODIFIED
Python sources (/root/reference/interact_drive/..., experiments/...) run in this container,
where TensorFlow cannot be installed, so that tests/golden/make_golden.py can record golden
vectors from the reference's own control flow (multi-start loop, world.step ordering,
check_plans quirk, replanning teleport, MPC_ORD return accumulation ...).

Backed by float32 torch CPU tensors and torch.autograd.  Where torch's gradient convention
differs from TensorFlow 2.1's (python/ops/math_grad.py) the op is a custom autograd Function
that follows TF:
  * minimum / maximum   -> whole gradient to x where x <= y / x >= y, else to y
                           (_MinimumGrad / _MaximumGrad)
  * reduce_min / max    -> gradient split evenly among tied extrema (_MinOrMaxGrad)
  * abs                 -> grad * sign(x) (_AbsGrad)            [same as torch]
  * where               -> gradient only to the selected branch [same as torch]
  * keras SGD.minimize  -> var -= lr * grad, momentum 0 (ResourceApplyGradientDescent)
It is NOT TensorFlow: kernel-level numerics (sin/cos/exp rounding, reduce_sum order) are
torch's.  Anything not needed by the hot path raises on use.
"""
from __future__ import annotations

import logging
from typing import Any, Iterable

import numpy as np
import torch

torch.set_num_threads(1)
# TF eager records gradients only under a tape (or inside optimizer.minimize): mirror that.
torch.set_grad_enabled(False)

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
__version__ = "2.1.0-shim"


def _conv_arg(a):
    if isinstance(a, np.ndarray):
        t = torch.from_numpy(np.array(a, copy=True, order='C'))
        return t.to(torch.float32) if t.is_floating_point() else t
    if isinstance(a, np.generic):
        return a.item() if not isinstance(a, np.floating) else float(a)
    if isinstance(a, (list, tuple)):
        return type(a)(_conv_arg(x) for x in a)
    return a


# Plain torch.Tensor is used as tf.Tensor / tf.Variable (a subclass with __torch_function__
# costs ~4x in eager dispatch).  The few TF methods the reference calls are patched onto
# torch.Tensor -- acceptable because this module is only ever imported by the golden-vector
# generator process (never by the product, never inside the pytest process).
Tensor = torch.Tensor
_torch_numpy = torch.Tensor.numpy


def _tf_numpy(self, *a, **k):  # tf returns a copy of the value
    return _torch_numpy(self.detach()).copy()


def _tf_assign(self, value):
    with torch.no_grad():
        self.copy_(_to_tensor(value, self.dtype))
    return self


torch.Tensor.numpy = _tf_numpy
torch.Tensor.assign = _tf_assign


def _install_numpy_binops():
    """`ndarray <op> Tensor` defers to the reflected op; accept numpy operands there."""
    def wrap(name):
        base = getattr(torch.Tensor, name)

        def op(self, other):
            return base(self, _conv_arg(other))
        op.__name__ = name
        return op
    for name in ("__mul__", "__rmul__", "__add__", "__radd__", "__sub__", "__rsub__",
                 "__truediv__", "__rtruediv__"):
        setattr(torch.Tensor, name, wrap(name))


_install_numpy_binops()


def _to_tensor(x: Any, dtype=None) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, (list, tuple)) and any(isinstance(e, torch.Tensor) for e in _flatten(x)):
        t = torch.stack([_to_tensor(e, dtype) for e in x])
    else:
        a = np.asarray(x)
        if a.dtype == np.float64 and dtype is None:
            dtype = torch.float32          # TF: Python floats become float32
        t = torch.from_numpy(np.array(a, copy=True, order='C'))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def _flatten(x):
    for e in x:
        if isinstance(e, (list, tuple)):
            yield from _flatten(e)
        else:
            yield e


def _like(x, ref: torch.Tensor) -> torch.Tensor:
    """Python / numpy scalars take the dtype of the tensor they meet (TF auto-conversion)."""
    if isinstance(x, torch.Tensor):
        return x
    return _to_tensor(x, ref.dtype)


def constant(value, dtype=None):
    t = _to_tensor(value, dtype)
    return t.detach().clone()


def convert_to_tensor(value, dtype=None):
    return _to_tensor(value, dtype)


def Variable(initial_value, dtype=None, **_):
    t = _to_tensor(initial_value, dtype).detach().clone()
    t.requires_grad_(True)
    return t


def identity(x):
    return _to_tensor(x).clone()


def function(fn=None, **_):
    if fn is None:
        return lambda f: f
    return fn


def get_logger():
    return logging.getLogger("tensorflow-shim")


def print(*args, **kwargs):  # noqa: A001  (tf.print)
    import builtins
    builtins.print(*args, **kwargs)


# ---- elementwise -----------------------------------------------------------------------
def cos(x):
    return torch.cos(_to_tensor(x))


def sin(x):
    return torch.sin(_to_tensor(x))


def exp(x):
    return torch.exp(_to_tensor(x))


def square(x):
    x = _to_tensor(x)
    return x * x


def zeros_like(x):
    return torch.zeros_like(_to_tensor(x))


def ones_like(x):
    return torch.ones_like(_to_tensor(x))


def abs(x):  # noqa: A001
    return torch.abs(_to_tensor(x))


def less(x, y):
    x = _to_tensor(x)
    return x < _like(y, x)


class _MinimumTF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        mask = x <= y
        ctx.save_for_backward(mask)
        ctx.xs, ctx.ys = x.shape, y.shape
        return torch.where(mask, x, y)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        z = torch.zeros_like(g)
        return torch.where(mask, g, z).sum_to_size(ctx.xs), torch.where(mask, z, g).sum_to_size(ctx.ys)


class _MaximumTF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        mask = x >= y
        ctx.save_for_backward(mask)
        ctx.xs, ctx.ys = x.shape, y.shape
        return torch.where(mask, x, y)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        z = torch.zeros_like(g)
        return torch.where(mask, g, z).sum_to_size(ctx.xs), torch.where(mask, z, g).sum_to_size(ctx.ys)


def _binary_args(x, y):
    if isinstance(x, torch.Tensor):
        return x, _like(y, x)
    if isinstance(y, torch.Tensor):
        return _like(x, y), y
    return _to_tensor(x), _to_tensor(y)


def minimum(x, y):
    x, y = _binary_args(x, y)
    return _MinimumTF.apply(x, y)


def maximum(x, y):
    x, y = _binary_args(x, y)
    return _MaximumTF.apply(x, y)


def where(cond, x, y):
    if isinstance(x, torch.Tensor):
        y = _like(y, x)
    elif isinstance(y, torch.Tensor):
        x = _like(x, y)
    else:
        x, y = _to_tensor(x), _to_tensor(y)
    return torch.where(cond, x, y)


# ---- shape ops / reductions ----------------------------------------------------------------
def stack(values, axis=0):
    vals = list(values)
    ref = next((v for v in vals if isinstance(v, torch.Tensor)), None)
    vals = [_to_tensor(v, ref.dtype if ref is not None else None) for v in vals]
    return torch.stack(vals, dim=axis)


def concat(values, axis=0):
    return torch.cat([_to_tensor(v) for v in values], dim=axis)


def reshape(x, shape):
    if isinstance(x, (list, tuple)):
        x = stack(x)
    return torch.reshape(_to_tensor(x), tuple(shape) if isinstance(shape, Iterable) else (shape,))


def _as_stacked(x):
    if isinstance(x, (list, tuple)):
        return stack(x)
    return _to_tensor(x)


def reduce_sum(x, axis=None):
    x = _as_stacked(x)
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis)


class _ReduceExtremumTF(torch.autograd.Function):
    """reduce_min / reduce_max over one axis with TF's even split among ties."""

    @staticmethod
    def forward(ctx, x, dim, is_max):
        y = torch.amax(x, dim=dim) if is_max else torch.amin(x, dim=dim)
        ind = (x == y.unsqueeze(dim)).to(x.dtype)
        ctx.save_for_backward(ind)
        ctx.dim = dim
        return y

    @staticmethod
    def backward(ctx, g):
        (ind,) = ctx.saved_tensors
        num = ind.sum(dim=ctx.dim, keepdim=True)
        return ind / num * g.unsqueeze(ctx.dim), None, None


def reduce_min(x, axis=None):
    x = _as_stacked(x)
    if axis is None:
        x, axis = x.reshape(-1), 0
    return _ReduceExtremumTF.apply(x, axis, False)


def reduce_max(x, axis=None):
    x = _as_stacked(x)
    if axis is None:
        x, axis = x.reshape(-1), 0
    return _ReduceExtremumTF.apply(x, axis, True)


# ---- autodiff / optimiser ------------------------------------------------------------------
class GradientTape:
    """Only what NaivePlanner-style code needs: tape.gradient(target, sources)."""

    def __init__(self, persistent=False, watch_accessed_variables=True):
        self.persistent = persistent

    def __enter__(self):
        self._prev = torch.is_grad_enabled()
        torch.set_grad_enabled(True)
        return self

    def __exit__(self, *exc):
        torch.set_grad_enabled(self._prev)
        return False

    def watch(self, t):
        if isinstance(t, (list, tuple)):
            for e in t:
                self.watch(e)
        elif isinstance(t, torch.Tensor) and not t.requires_grad:
            t.requires_grad_(True)

    def gradient(self, target, sources):
        single = not isinstance(sources, (list, tuple))
        srcs = [sources] if single else list(sources)
        with torch.enable_grad():
            grads = torch.autograd.grad(target, srcs, retain_graph=self.persistent, allow_unused=True)
        grads = [None if g is None else g for g in grads]
        return grads[0] if single else grads


class _SGD:
    """tf.keras.optimizers.SGD(learning_rate) with momentum 0: var -= lr * grad."""

    def __init__(self, learning_rate=0.01, momentum=0.0, **_):
        if momentum:
            raise NotImplementedError("shim SGD: momentum 0 only")
        self.learning_rate = learning_rate

    def minimize(self, loss, var_list):
        var_list = list(var_list)
        with torch.enable_grad():
            value = loss() if callable(loss) else loss
            grads = torch.autograd.grad(value, var_list, allow_unused=True)
        lr = torch.tensor(self.learning_rate, dtype=torch.float32)
        with torch.no_grad():
            for v, g in zip(var_list, grads):
                if g is not None:
                    v.sub_(lr.to(v.dtype) * g)


class _Unavailable:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise NotImplementedError(f"tensorflow shim: {self._name} is outside the hot path")

    def __getattr__(self, item):
        return _Unavailable(f"{self._name}.{item}")


class _Optimizers:
    SGD = _SGD
    Adam = _Unavailable("keras.optimizers.Adam")


class _Keras:
    optimizers = _Optimizers


keras = _Keras
linalg = _Unavailable("linalg")
nn = _Unavailable("nn")


def __getattr__(name):  # anything else: importable, unusable
    return _Unavailable(name)
