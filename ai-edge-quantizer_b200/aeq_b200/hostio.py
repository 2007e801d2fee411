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


# Above this size a transfer goes through the library's pinned ring and staging workers
# (aeqb_host_copy_in / aeqb_host_copy_out, csrc/host_pipeline.cu): a pageable array moves at
# link speed instead of the one-thread staging `tensor.to(device)` does.
_STAGED_COPY_MIN_BYTES = 1 << 20


def to_device(arr, dtype=None) -> torch.Tensor:
  """Copies a NumPy array (any alignment, read-only allowed) to the current device."""
  if isinstance(arr, torch.Tensor):
    t = arr.to(device())
    return t if dtype is None else t.to(_torch_dtype(dtype))
  a = np.asarray(arr)
  if dtype is not None and a.dtype != np.dtype(dtype):
    a = a.astype(dtype)
  a = np.ascontiguousarray(a)
  dev = device()
  if a.nbytes >= _STAGED_COPY_MIN_BYTES and a.dtype in _TORCH_DTYPES:
    from . import _lib
    t = torch.empty(a.shape, dtype=_TORCH_DTYPES[a.dtype], device=dev)
    _lib.call("aeqb_host_copy_in", t.data_ptr(), a.ctypes.data, a.nbytes,
              torch.cuda.current_stream().cuda_stream)
    return t
  with warnings.catch_warnings():
    warnings.simplefilter("ignore", UserWarning)  # non-writable mmap views
    t = torch.from_numpy(a)
  return t.to(dev, non_blocking=False)


def to_host(t: torch.Tensor) -> np.ndarray:
  """Device tensor -> freshly allocated NumPy array (synchronises)."""
  t = t.detach()
  nbytes = t.numel() * t.element_size()
  if t.is_cuda and nbytes >= _STAGED_COPY_MIN_BYTES and t.dtype in _NUMPY_DTYPES:
    from . import _lib
    t = t.contiguous()
    out = np.empty(tuple(t.shape), dtype=_NUMPY_DTYPES[t.dtype])
    with torch.cuda.device(t.device):
      _lib.call("aeqb_host_copy_out", out.ctypes.data, t.data_ptr(), nbytes,
                torch.cuda.current_stream().cuda_stream)
    return out
  return t.cpu().numpy()


_TORCH_DTYPES = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                 np.dtype(np.int32): torch.int32, np.dtype(np.int8): torch.int8,
                 np.dtype(np.uint8): torch.uint8, np.dtype(np.int16): torch.int16,
                 np.dtype(np.float16): torch.float16, np.dtype(np.int64): torch.int64}
_NUMPY_DTYPES = {v: k for k, v in _TORCH_DTYPES.items()}


def _torch_dtype(dtype):
  return {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32,
          np.dtype(np.int8): torch.int8, np.dtype(np.int16): torch.int16}[np.dtype(dtype)]
