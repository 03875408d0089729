"""Process-wide engine handle used by the drop-in classes (one per CUDA device)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .engine import Engine, HostContext

_engines: Dict[int, Engine] = {}
_host_contexts: Dict[int, HostContext] = {}
_default_device: Optional[int] = None


def set_default_device(device: int) -> None:
    """Device the drop-in classes plan on (one process per GPU: pass LOCAL_RANK)."""
    global _default_device
    _default_device = int(device)


def get_engine(device: Optional[int] = None) -> Engine:
    """The engine of `device` (default: set_default_device(), else the current torch device).
    Raises OcdCudaError when no GPU is visible -- there is no CPU planner to fall back to."""
    if device is None:
        device = _default_device
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]


def get_host_context(device: Optional[int] = None) -> HostContext:
    """The host-buffer context (`ocd_ctx`) of `device`: numpy in, numpy out, staging buffers and the captured
    episode-call graph kept between calls."""
    idx = get_engine(device).device.index or 0
    if idx not in _host_contexts:
        _host_contexts[idx] = HostContext(idx)
    return _host_contexts[idx]


def as_f32(x, shape=None) -> np.ndarray:
    """Host float32 copy of an array-like or tensor (the reference casts everything with
    tf.constant(..., dtype=tf.float32): car.py:55,118)."""
    if torch.is_tensor(x):
        x = x.detach().cpu().numpy()
    a = np.array(x, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a
