"""Device-resident API: torch CUDA tensors in, torch CUDA tensors out.

torch is used only as the device-memory container and for the current stream;
every number is produced by libaeqb200.so through the C ABI (include/aeqb200.h).
These wrappers enqueue on torch's current stream and do not synchronise.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else t.data_ptr()


def _stream():
  return torch.cuda.current_stream().cuda_stream


def _check_f32_2d(x: torch.Tensor) -> None:
  if not x.is_cuda:
    raise ValueError("aeq_b200 needs a CUDA tensor (no CPU fallback)")
  if x.dtype != torch.float32:
    raise ValueError(f"only float32 tensors are quantised, got {x.dtype}")
  if x.dim() != 2 or not x.is_contiguous():
    raise ValueError("expected a C-contiguous 2-D [rows, cols] tensor")


class Requantized(NamedTuple):
  """Outputs of a fused requantisation, all on the device."""
  q: Optional[torch.Tensor]          # int8 [rows, cols]
  packed: Optional[torch.Tensor]     # uint8 [rows*cols*bits/8]
  scale: torch.Tensor                # fp32 [rows, 1] / [rows, cols/block] / [1, 1]
  zero_point: Optional[torch.Tensor]  # int32, scale's shape (None for blockwise)
  scale_f16: Optional[torch.Tensor] = None  # fp16 [rows, cols/block], blockwise only


def requant_rows(x: torch.Tensor, bits: int, symmetric: bool = True,
                 clip: Optional[torch.Tensor] = None, want_q: bool = True,
                 want_packed: bool = False) -> Requantized:
  """Per-channel min/max -> scale/zp -> quantise, one HBM pass (aeqb_requant_rows_f32)."""
  _check_f32_2d(x)
  rows, cols = x.shape
  dev = x.device
  q = torch.empty((rows, cols), dtype=torch.int8, device=dev) if want_q else None
  packed = None
  if want_packed:
    if (rows * cols * bits) % 8:
      raise ValueError("packed output needs rows*cols*bits to be a multiple of 8")
    packed = torch.empty(rows * cols * bits // 8, dtype=torch.uint8, device=dev)
  scale = torch.empty((rows, 1), dtype=torch.float32, device=dev)
  zp = torch.empty((rows, 1), dtype=torch.int32, device=dev)
  _lib.call("aeqb_requant_rows_f32", _ptr(x), rows, cols, bits, int(symmetric),
            _ptr(clip), _ptr(q), _ptr(packed), _ptr(scale), _ptr(zp), _stream())
  return Requantized(q, packed, scale, zp)


def requant_given_minmax(x: torch.Tensor, mn: torch.Tensor, mx: torch.Tensor,
                         bits: int, symmetric: bool, per_row: bool,
                         clip: Optional[torch.Tensor] = None, want_q: bool = True,
                         want_packed: bool = False) -> Requantized:
  """Scale/zp from caller-supplied min/max, then quantise (QSV / per-tensor path)."""
  _check_f32_2d(x)
  rows, cols = x.shape
  dev = x.device
  n = rows if per_row else 1
  if mn.numel() != n or mx.numel() != n:
    raise ValueError(f"min/max must hold {n} values")
  q = torch.empty((rows, cols), dtype=torch.int8, device=dev) if want_q else None
  packed = (torch.empty(rows * cols * bits // 8, dtype=torch.uint8, device=dev)
            if want_packed else None)
  scale = torch.empty((n, 1), dtype=torch.float32, device=dev)
  zp = torch.empty((n, 1), dtype=torch.int32, device=dev)
  _lib.call("aeqb_requant_given_minmax_f32", _ptr(x), rows, cols, bits,
            int(symmetric), _ptr(mn), _ptr(mx), _ptr(clip), int(per_row), _ptr(q),
            _ptr(packed), _ptr(scale), _ptr(zp), _stream())
  return Requantized(q, packed, scale, zp)


def requant_blocks(x: torch.Tensor, block: int, bits: int,
                   clip: Optional[torch.Tensor] = None, want_q: bool = True,
                   want_packed: bool = False,
                   want_scale_f16: bool = True) -> Requantized:
  """Blockwise symmetric requantisation with fp16-rounded scales (aeqb_requant_blocks_f32)."""
  _check_f32_2d(x)
  rows, cols = x.shape
  if cols % block:
    raise ValueError(
        f"Quantized dimension {cols} in tensor shape {tuple(x.shape)} is not"
        f" divisible by block size {block}.")
  dev = x.device
  nb = cols // block
  q = torch.empty((rows, cols), dtype=torch.int8, device=dev) if want_q else None
  packed = (torch.empty(rows * cols // 2, dtype=torch.uint8, device=dev)
            if want_packed else None)
  scale = torch.empty((rows, nb), dtype=torch.float32, device=dev)
  f16 = torch.empty((rows, nb), dtype=torch.float16, device=dev) if want_scale_f16 else None
  _lib.call("aeqb_requant_blocks_f32", _ptr(x), rows, cols, block, bits, _ptr(clip),
            _ptr(q), _ptr(packed), _ptr(scale), _ptr(f16), _stream())
  return Requantized(q, packed, scale, None, f16)
