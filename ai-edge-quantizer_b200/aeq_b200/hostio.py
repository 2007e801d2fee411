"""Host <-> device movement for the NumPy-facing plug-in functions.

Inputs are read-only NumPy views onto the mmap'd flatbuffer (reference
utils/tfl_flatbuffer_utils.py:258-263; possibly unaligned, never written);
outputs are fresh NumPy arrays owned by the returned UniformQuantParams.
torch is the device-memory container.  No CUDA device => RuntimeError: there is
no CPU implementation to fall back to.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch


def device() -> torch.device:
  if not torch.cuda.is_available():
    raise RuntimeError(
        "aeq_b200 computes on a CUDA device only (no CPU fallback) and"
        " torch.cuda.is_available() is False")
  return torch.device("cuda", torch.cuda.current_device())


def to_device(arr, dtype=None) -> torch.Tensor:
  """Copies a NumPy array (any alignment, read-only allowed) to the current device."""
  if isinstance(arr, torch.Tensor):
    t = arr.to(device())
    return t if dtype is None else t.to(_torch_dtype(dtype))
  a = np.asarray(arr)
  if dtype is not None and a.dtype != np.dtype(dtype):
    a = a.astype(dtype)
  a = np.ascontiguousarray(a)
  with warnings.catch_warnings():
    warnings.simplefilter("ignore", UserWarning)  # non-writable mmap views
    t = torch.from_numpy(a)
  return t.to(device(), non_blocking=False)


def to_host(t: torch.Tensor) -> np.ndarray:
  """Device tensor -> freshly allocated NumPy array (synchronises)."""
  return t.detach().cpu().numpy()


def _torch_dtype(dtype):
  return {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32,
          np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16}[np.dtype(dtype)]
