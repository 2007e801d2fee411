"""NumPy in -> NumPy out through the C ABI's host-buffer entry points.

This is the drop-in call: host arrays (read-only views onto the mmap'd
flatbuffer in the reference, utils/tfl_flatbuffer_utils.py:254-263) go in, fresh
host arrays come out; the chunked H2D -> kernel -> D2H pipeline lives in
libaeqb200.so (csrc/host_pipeline.cu).  `pinned_empty` hands out page-locked
arrays for callers that want the DMA to run in place.
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Optional, Sequence

import numpy as np

from . import _lib


def _require_gpu() -> None:
  import torch
  if not torch.cuda.is_available():
    raise RuntimeError(
        "aeq_b200 computes on a CUDA device only (no CPU fallback) and"
        " torch.cuda.is_available() is False")


class _PinnedBlock:
  def __init__(self, nbytes: int):
    self.ptr = _lib.load().aeqb_host_alloc(nbytes)
    if not self.ptr:
      raise MemoryError(f"aeqb_host_alloc({nbytes}) failed")
    weakref.finalize(self, _lib.load().aeqb_host_free, self.ptr)


def pinned_empty(shape, dtype) -> np.ndarray:
  """Page-locked NumPy array (freed when the array and its views are collected)."""
  _require_gpu()
  dtype = np.dtype(dtype)
  n = int(np.prod(shape, dtype=np.int64))
  block = _PinnedBlock(max(1, n * dtype.itemsize))
  buf = (ctypes.c_char * max(1, n * dtype.itemsize)).from_address(block.ptr)
  buf._aeqb_owner = block  # keep the allocation alive as long as the ctypes buffer is
  return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def set_devices(devices=None) -> list[int]:
  """GPUs the host-buffer calls fan out over: a list of device indices, "all" (every visible
  GPU) or None (default: the current device only).  Returns the list in effect."""
  _require_gpu()
  import torch
  if devices == "all":
    devices = list(range(torch.cuda.device_count()))
  devices = [int(d) for d in (devices or [])]
  arr = (ctypes.c_int * max(1, len(devices)))(*devices)
  _lib.call("aeqb_host_set_devices", ctypes.cast(arr, ctypes.c_void_p), len(devices))
  return devices or [torch.cuda.current_device()]


def _f32_2d(w: np.ndarray) -> np.ndarray:
  if w.dtype != np.float32:
    raise ValueError(f"only float32 tensors are quantised, got {w.dtype}")
  if w.ndim != 2:
    raise ValueError("expected a 2-D [rows, cols] array")
  return w if w.flags.c_contiguous else np.ascontiguousarray(w)


def _p(a: Optional[np.ndarray]):
  return None if a is None else a.ctypes.data


def requant_rows(ws: Sequence[np.ndarray], bits: int, symmetric: bool = True,
                 want_q: bool = True, want_packed: bool = False, outs=None, alloc=np.empty):
  """Per-channel requantisation of host arrays; returns [(q, packed, scale[rows,1], zp[rows,1])]."""
  _require_gpu()
  ws = [_f32_2d(w) for w in ws]
  if outs is None:
    outs = []
    for w in ws:
      r, c = w.shape
      outs.append((alloc((r, c), np.int8) if want_q else None,
                   alloc((r * c * bits // 8,), np.uint8) if want_packed else None,
                   alloc((r, 1), np.float32), alloc((r, 1), np.int32)))
  jobs = (_lib.RowsJob * len(ws))()
  for i, (w, o) in enumerate(zip(ws, outs)):
    jobs[i] = _lib.RowsJob(_p(w), w.shape[0], w.shape[1], None, _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]))
  _lib.call("aeqb_host_requant_rows_batch_f32", ctypes.cast(jobs, ctypes.c_void_p), len(ws), bits,
            int(symmetric))
  return outs


def requant_mse_rows(ws: Sequence[np.ndarray], bits: int, multiplier: float, want_packed: bool = False,
                     outs=None, alloc=np.empty):
  """MSE per channel (scale = multiplier * RMS(row), zero point 0) of host arrays through the same
  pipeline; returns [(q, packed, scale[rows,1], zp[rows,1])].  Rows must hold a multiple of 128
  values and at most 96 KiB (the fused kernel's RMS statistic)."""
  _require_gpu()
  ws = [_f32_2d(w) for w in ws]
  if outs is None:
    outs = []
    for w in ws:
      r, c = w.shape
      outs.append((alloc((r, c), np.int8), alloc((r * c * bits // 8,), np.uint8) if want_packed else None,
                   alloc((r, 1), np.float32), alloc((r, 1), np.int32)))
  jobs = (_lib.RowsJob * len(ws))()
  for i, (w, o) in enumerate(zip(ws, outs)):
    jobs[i] = _lib.RowsJob(_p(w), w.shape[0], w.shape[1], None, _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]))
  _lib.call("aeqb_host_requant_mse_rows_batch_f32", ctypes.cast(jobs, ctypes.c_void_p), len(ws), bits,
            float(multiplier))
  return outs


def requant_blocks(ws: Sequence[np.ndarray], block: int, bits: int, want_q: bool = True,
                   want_packed: bool = False, want_scale: bool = True,
                   want_scale_f16: bool = False, outs=None, alloc=np.empty):
  """Blockwise requantisation of host arrays; returns [(q, packed, scale, scale_f16)]."""
  _require_gpu()
  ws = [_f32_2d(w) for w in ws]
  if outs is None:
    outs = []
    for w in ws:
      r, c = w.shape
      if c % block:
        raise ValueError(
            f"Quantized dimension {c} in tensor shape {w.shape} is not"
            f" divisible by block size {block}.")
      outs.append((alloc((r, c), np.int8) if want_q else None,
                   alloc((r * c // 2,), np.uint8) if want_packed else None,
                   alloc((r, c // block), np.float32) if want_scale else None,
                   alloc((r, c // block), np.float16) if want_scale_f16 else None))
  jobs = (_lib.BlocksJob * len(ws))()
  for i, (w, o) in enumerate(zip(ws, outs)):
    jobs[i] = _lib.BlocksJob(_p(w), w.shape[0], w.shape[1], None, _p(o[0]), _p(o[1]), _p(o[2]), _p(o[3]))
  _lib.call("aeqb_host_requant_blocks_batch_f32", ctypes.cast(jobs, ctypes.c_void_p), len(ws), block,
            bits)
  return outs
