"""`MSE`: closed-form scale = k * RMS(row), then symmetric quantisation.

Mirror of ai_edge_quantizer/algorithms/uniform_quantize/mse.py
(`_MSE_QUANT_MULS` :30-33, `get_tensor_quant_params` :36-128).  Device flow: one
upload, then per channel the fused `aeqb_requant_mse_rows_f32` (sum of squares, scale and
integers in one pass); for a whole tensor `aeqb_mse_scale_rows_f32` followed by
`aeqb_quantize_f32` (zero point is int32 zeros, mse.py:109).
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from ... import hostio
from ... import qtyping
from ..utils import common_utils
from . import naive_min_max_quantize
from . import uniform_quantize_tensor as uqt

ALGORITHM_KEY = "MSE"
_MSE_QUANT_MULS = {8: 0.05408, 4: 0.37755}


def get_tensor_quant_params(
    op_info: qtyping.OpInfo,
    tensor_quant_config: qtyping.TensorQuantizationConfig,
    tensor_content: Optional[np.ndarray] = None,
    tensor_qsv: Optional[dict[str, Any]] = None,
) -> qtyping.UniformQuantParams:
  cfg = tensor_quant_config
  if uqt.is_blockwise(cfg.granularity):
    raise ValueError("Blockwise quantization is not supported for MSE quantization.")
  if tensor_content is None:
    return naive_min_max_quantize.get_tensor_quant_params(op_info, cfg, tensor_content, tensor_qsv)
  if not cfg.symmetric:
    raise ValueError(
        f"Unsupported symmetry: {cfg.symmetric}. MSE"
        " supports symmetric quantization only for now.")
  if tensor_qsv and "min" in tensor_qsv and "max" not in tensor_qsv:
    raise ValueError(
        "min and max must be provided to produce tensor quantization"
        " parameters. Check if the correct calibration results are passed into"
        " the ParamsGenerator.")
  if tensor_content.dtype != np.float32:
    raise ValueError(f"only float32 weights are quantised, got {tensor_content.dtype}")
  from ... import device
  multiplier = _MSE_QUANT_MULS[cfg.num_bits]  # KeyError for other widths, like the reference
  qdim = common_utils.get_weight_quantized_dim(op_info, tensor_content, cfg.granularity)
  shape = tensor_content.shape
  if qdim is None:
    x = hostio.to_device(tensor_content.reshape(1, -1), np.float32)
    pshape = [1] * tensor_content.ndim
  else:  # channels of any axis as rows (dim 0: a view; others: aeqb_swap_axes)
    x = device.channel_rows(hostio.to_device(tensor_content, np.float32), shape, qdim)
    pshape = [1] * tensor_content.ndim
    pshape[qdim] = shape[qdim]
  if qdim is not None:  # per channel: scale and integers in one pass over the weight
    out = device.requant_mse_rows(x, cfg.num_bits, multiplier)
    scale, q = out.scale, device.channel_rows_back(out.q, shape, qdim)
  else:          # whole tensor: grid-wide sum of squares first
    scale = device.mse_scale_rows(x, multiplier)
    q = device.quantize(x, scale.reshape(-1), None, cfg.num_bits, True, x.shape[0], x.shape[1])
  scale_np = hostio.to_host(scale).reshape(pshape)
  return qtyping.UniformQuantParams(
      scale=scale_np, zero_point=np.zeros_like(scale_np, dtype=np.int32), num_bits=cfg.num_bits,
      symmetric=cfg.symmetric, quantized_dimension=qdim,
      quantized_data=hostio.to_host(q).reshape(shape), block_size=0)
