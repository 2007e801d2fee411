"""`OCTAV`: MSE-optimal clipping by Newton iterations, then symmetric quantisation.

Mirror of ai_edge_quantizer/algorithms/uniform_quantize/octav.py
(`_guess_clipping_with_octav` :30-112, `get_tensor_quant_params` :115-227).
Device flow for a constant weight: one upload, `aeqb_octav_clip_{rows,blocks}_f32`
(whole Newton trajectory of every row / block in one HBM pass + the reference's
global early stop), then the fused min/max -> clipped scale -> quantise kernel
with the clipping constants.
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

ALGORITHM_KEY = "OCTAV"
_Gran = qtyping.QuantGranularity
MAX_ITERATIONS = 10  # octav.py:189


def clipping_constants_device(x_dev, op_info: qtyping.OpInfo,
                              cfg: qtyping.TensorQuantizationConfig, shape):
  """Device tensor of clipping constants shaped like the weight's min/max ([R,1] / [R,C/B] / [1,1])."""
  from ... import device
  gran = cfg.granularity
  divisor = 3.0 if cfg.symmetric else 12.0
  block = uqt.extract_block_size_from_granularity(gran)
  probe = np.empty(shape, np.bool_)
  if block:
    uqt.reshape_data_for_blockwise(probe, op_info.op_name, gran)  # the reference's divisibility error
    return device.octav_clip_blocks(x_dev.reshape(-1, shape[-1]), block, cfg.num_bits,
                                    MAX_ITERATIONS, divisor)
  qdim = common_utils.get_weight_quantized_dim(op_info, probe, gran)
  if qdim is None:
    return device.octav_clip_rows(x_dev.reshape(1, -1), cfg.num_bits, MAX_ITERATIONS, divisor)
  return device.octav_clip_rows(device.channel_rows(x_dev.reshape(tuple(shape)), shape, qdim),
                                cfg.num_bits, MAX_ITERATIONS, divisor)


def guess_clipping_with_octav(x: np.ndarray, bits: int, axis, max_iterations: int = 10,
                              exponent_divisor: float = 3.0, early_stop: bool = True) -> np.ndarray:
  """NumPy-facing `_guess_clipping_with_octav` for the layouts the kernels cover:
  reduce all but axis 0, 3-D [R, C/B, B] reduce over axis 2, or axis None / all axes."""
  from ... import device
  x = np.asarray(x, dtype=np.float32)
  axes = None if axis is None else ((axis,) if isinstance(axis, int) else tuple(axis))
  xd = hostio.to_device(x, np.float32)
  if axes is None or len(axes) == x.ndim:
    clip = device.octav_clip_rows(xd.reshape(1, -1), bits, max_iterations, exponent_divisor, early_stop)
    out_shape = (1,) if axes is None else (1,) * x.ndim
  elif axes == (x.ndim - 1,) and x.ndim == 3 and x.shape[2] in (32, 64, 128, 256):
    clip = device.octav_clip_blocks(xd.reshape(x.shape[0], -1), x.shape[2], bits, max_iterations,
                                    exponent_divisor, early_stop)
    out_shape = (x.shape[0], x.shape[1], 1)
  elif axes == tuple(range(1, x.ndim)):
    clip = device.octav_clip_rows(xd.reshape(x.shape[0], -1), bits, max_iterations,
                                  exponent_divisor, early_stop)
    out_shape = (x.shape[0],) + (1,) * (x.ndim - 1)
  else:
    raise NotImplementedError(f"OCTAV reduction over axes {axes} of a rank-{x.ndim} tensor")
  return hostio.to_host(clip).reshape(out_shape)


def get_tensor_quant_params(
    op_info: qtyping.OpInfo,
    tensor_quant_config: qtyping.TensorQuantizationConfig,
    tensor_content: Optional[np.ndarray] = None,
    tensor_qsv: Optional[dict[str, Any]] = None,
    x_dev=None,
) -> qtyping.UniformQuantParams:
  """Quantisation parameters + quantised data; runtime tensors fall back to min-max."""
  cfg = tensor_quant_config
  if tensor_content is None:
    return naive_min_max_quantize.get_tensor_quant_params(op_info, cfg, tensor_content, tensor_qsv)
  if not cfg.symmetric:
    raise ValueError(
        f"Unsupported symmetry: {cfg.symmetric}. OCTAV"
        " supports symmetric quantization only for now.")
  if tensor_content.dtype != np.float32:
    raise ValueError(f"only float32 weights are quantised, got {tensor_content.dtype}")
  if x_dev is None:
    x_dev = hostio.to_device(tensor_content, np.float32)
  clip = clipping_constants_device(x_dev, op_info, cfg, tensor_content.shape)
  if not tensor_qsv or "min" not in tensor_qsv:
    return naive_min_max_quantize.quantize_weight(op_info, cfg, tensor_content, clip=clip,
                                                  x_dev=x_dev)
  if "max" not in tensor_qsv:
    raise ValueError(
        "min and max must be provided to produce tensor quantization"
        " parameters. Check if the correct calibration results are passed into"
        " the ParamsGenerator.")
  # Calibrated min / max supplied: unfused scale + quantise with the same constants.
  mn = np.asarray(tensor_qsv["min"])
  zp, scale = uqt.tensor_zp_scale_from_min_max(
      mn, tensor_qsv["max"], cfg.num_bits, cfg.symmetric, cfg.granularity,
      hostio.to_host(clip).reshape(mn.shape))
  params = qtyping.UniformQuantParams(
      scale=scale, zero_point=zp, num_bits=cfg.num_bits, symmetric=cfg.symmetric,
      quantized_dimension=common_utils.get_weight_quantized_dim(op_info, tensor_content,
                                                                cfg.granularity),
      block_size=uqt.extract_block_size_from_granularity(cfg.granularity))
  q = uqt.uniform_quantize(hostio.to_host(x_dev).reshape(tensor_content.shape), params,
                           uqt.is_blockwise(cfg.granularity))
  return dataclasses.replace(params, quantized_data=q)
