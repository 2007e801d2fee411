"""`dequantized_weight_recovery`: recovers the integer weights of a fake-quantised (QAT) tensor.

Mirror of ai_edge_quantizer/algorithms/uniform_quantize/dequantized_weight_recovery.py
(`get_zp_scale_from_dequantized_symmetric_weights` :132-217, `get_tensor_quant_params`
:220-304, `_validate_recovered_weights` :36-63, `_check_unique_values` :80-129).

Device flow: one upload; `aeqb_dwr_scales_f32` sorts |w| per group (tensor / channel / block)
in shared memory and takes the smallest step above 1e-9; `aeqb_quantize_f32` re-quantises with
that scale; the recovery check is `aeqb_dequantize_f32` + `aeqb_max_abs_diff_f32` (one float
comes back).  Only the failure path (the unique-value diagnosis of the error message) runs on
the host, like the reference's own Python loop.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Optional

import numpy as np

from ... import hostio
from ... import qtyping
from ..utils import common_utils
from . import naive_min_max_quantize
from . import uniform_quantize_tensor as uqt

ALGORITHM_KEY = "dequantized_weight_recovery"


def _groups(tensor_content: np.ndarray, quantized_dimension, block_size: int):
  """(2-D [groups, group_len] view/copy in group-major order, parameter shape)."""
  shape = tensor_content.shape
  if quantized_dimension is None:
    return tensor_content.reshape(1, -1), (1, 1)
  if block_size > 0:
    if quantized_dimension != tensor_content.ndim - 1:
      raise ValueError("blockwise quantisation cuts the last axis")
    target = list(shape)  # common_utils.get_blockwise_shape (:1221-1247)
    if target[quantized_dimension] % block_size != 0:
      raise ValueError(
          f"Dimension {target[quantized_dimension]} along axis {quantized_dimension} is not"
          f" divisible by block size {block_size}")
    target[quantized_dimension] //= block_size
    return tensor_content.reshape(-1, block_size), tuple(target)
  perm = [quantized_dimension] + [i for i in range(tensor_content.ndim) if i != quantized_dimension]
  rows = np.transpose(tensor_content, perm).reshape(shape[quantized_dimension], -1)
  target = [1] * tensor_content.ndim
  target[quantized_dimension] = shape[quantized_dimension]
  return rows, tuple(target)


def get_zp_scale_from_dequantized_symmetric_weights(
    dequant_vals: np.ndarray, quantized_dimension: Optional[int] = None, block_size: int = 0,
    min_scale: float = 1e-9) -> tuple[np.ndarray, np.ndarray]:
  """(zero_points, scales) of symmetric fake-quantised weights (reference :132-217)."""
  from ... import device
  if quantized_dimension not in (0, 1, None):
    raise ValueError(f"quantized_dimension must be 0, 1, or None. Got {quantized_dimension}")
  if min_scale != 1e-9:
    raise NotImplementedError("the device kernel implements the reference's default min_scale=1e-9")
  rows, target = _groups(np.asarray(dequant_vals), quantized_dimension, block_size)
  x = hostio.to_device(rows, np.float32)
  scales = hostio.to_host(device.dwr_scales(x, x.shape[0], x.shape[1])).reshape(target)
  if quantized_dimension is None:
    scales = scales.astype(np.float64)  # the reference builds np.array([[float]]) here (:160-161)
  return np.zeros_like(scales, dtype=np.int32), scales


def _check_unique_values(tensor_content, quantized_dimension, *, block_size: int, num_bits: int):
  """Error-message diagnosis only (reference :80-129)."""
  limit = 1 << num_bits
  rows, _ = _groups(tensor_content, quantized_dimension, block_size)
  max_found = 0
  for row in rows:
    max_found = max(max_found, np.unique(row).size)
    if max_found > limit:
      break
  return max_found, limit


def get_tensor_quant_params(
    op_info: qtyping.OpInfo,
    tensor_quant_config: qtyping.TensorQuantizationConfig,
    tensor_content: Optional[np.ndarray] = None,
    tensor_qsv: Optional[dict[str, Any]] = None,
) -> qtyping.UniformQuantParams:
  from ... import device
  cfg = tensor_quant_config
  if tensor_content is None:
    return naive_min_max_quantize.get_tensor_quant_params(op_info, cfg, tensor_content, tensor_qsv)
  block_size = uqt.extract_block_size_from_granularity(cfg.granularity)
  if not cfg.symmetric:
    raise ValueError("Only symmetric weights are supported for dequantized weight recovery.")
  if tensor_content.dtype != np.float32:
    raise ValueError(f"only float32 weights are quantised, got {tensor_content.dtype}")
  qdim = common_utils.get_weight_quantized_dim(op_info, tensor_content, cfg.granularity)
  if qdim not in (0, 1, None):
    raise ValueError(f"quantized_dimension must be 0, 1, or None. Got {qdim}")
  rows, target = _groups(tensor_content, qdim, block_size)
  x = hostio.to_device(rows, np.float32)
  groups, glen = x.shape
  scale_d = device.dwr_scales(x, groups, glen)
  q_d = device.quantize(x, scale_d, None, cfg.num_bits, True, groups, glen)
  scale = hostio.to_host(scale_d).reshape(target)
  if qdim is None:
    scale = scale.astype(np.float64)
  zp = np.zeros_like(scale, dtype=np.int32)
  params = qtyping.UniformQuantParams(
      scale=scale, zero_point=zp, num_bits=cfg.num_bits, symmetric=cfg.symmetric,
      quantized_dimension=qdim, block_size=block_size)
  if not op_info.op_quant_config.skip_checks:
    rec = device.dequantize(q_d, scale_d, None, groups, glen)
    max_diff = float(hostio.to_host(device.max_abs_diff(rec, x))[0])
    if max_diff > 1e-4:
      original = (
          "Failed to recover the original quantized values from dequantized"
          f" values. Max diff between recovered and original values: {max_diff}"
          " (tolerance: 0.0001)")
      max_found, limit = _check_unique_values(
          tensor_content, qdim, block_size=block_size, num_bits=cfg.num_bits)
      extra = (
          f"Detected a quantization group with {max_found} unique values, "
          f"which exceeds the limit of {limit} for"
          f" {cfg.num_bits}-bit quantization. This suggests"
          " the input tensor is NOT dequantized (fake-quantized) weights."
          " Please verify if you are using a QAT checkpoint."
          if max_found > limit else
          f"Max unique values in any group is {max_found} (limit: {limit})."
          " The recovery failed despite reasonable unique value count."
          " Check if the weights are symmetric or if tolerance is too"
          " tight.")
      raise RuntimeError(f"Failed to recover weights. Original error: {original}. {extra}")
  q = hostio.to_host(q_d)
  if qdim is not None and block_size == 0 and qdim != 0:  # undo the group-major transpose
    perm = [qdim] + [i for i in range(tensor_content.ndim) if i != qdim]
    moved = [tensor_content.shape[i] for i in perm]
    q = np.transpose(q.reshape(moved), np.argsort(perm))
  return dataclasses.replace(params, quantized_data=np.ascontiguousarray(q.reshape(tensor_content.shape)))


calibrate = naive_min_max_quantize.min_max_calibrate
