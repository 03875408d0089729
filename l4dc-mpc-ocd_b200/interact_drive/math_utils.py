"""Smooth helper functions of the reward features: mirror of interact_drive/math_utils.py
(`_f` :7-31, `smooth_threshold` :59-97, `smooth_bump` :135-180), evaluated by the engine's
`ocd_smooth_batch` operator.  Each constructor returns a callable like the reference's."""
from __future__ import annotations

import numpy as np

from .. import _native as N
from ..runtime import as_f32, get_engine


def _out(z, r):
    r = r.cpu().numpy()
    return np.float32(r) if np.ndim(z) == 0 else r


def _f(x, shape=5.0):
    """exp(-1/(shape*x)) for x > 0, else 0 (reference math_utils.py:7-31)."""
    return _out(x, get_engine().smooth(N.SMOOTH_F, as_f32(x), float(shape)))


def smooth_threshold(threshold, width=0.01, c=5.0):
    """Smooth step rising from 0 at threshold-width to 1 at threshold (reference math_utils.py:59-97).
    Only the reference's default sharpness c=5 is built into the kernel."""
    if float(c) != 5.0:
        raise ValueError("smooth_threshold: only c=5 is supported by the engine")
    thr, wd = float(threshold), float(width)

    def fn(z):
        return _out(z, get_engine().smooth(N.SMOOTH_THRESHOLD, as_f32(z), thr, wd))

    return fn


def smooth_bump(start, end):
    """Compactly supported bump equal to 1 at the centre of [start, end] (reference
    math_utils.py:135-180)."""
    a, b = float(as_f32(start)), float(as_f32(end))

    def fn(z):
        return _out(z, get_engine().smooth(N.SMOOTH_BUMP, as_f32(z), a, b))

    return fn
