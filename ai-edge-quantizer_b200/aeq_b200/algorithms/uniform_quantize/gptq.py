"""`GPTQ`: Hessian calibration + 64-column lazy-block OBS weight quantisation.

Mirror of ai_edge_quantizer/algorithms/uniform_quantize/gptq.py (`calibrate`
:55-108, `_prepare_hessian_inverse` :111-128, `_apply_gptq` :131-216,
`get_tensor_quant_params` :219-300).  Device flow for a weight: one upload ->
min/max at the configured granularity -> scale / zero point (no clipping) ->
damped Hessian inverse (fp64 Cholesky, fp32 triangular inverse, L^-T L^-1) ->
the column loop, all through libaeqb200.so.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Mapping, MutableMapping, Optional, Sequence

import numpy as np

from ... import hostio
from ... import qtyping
from ..utils import common_utils
from . import common_quantize
from . import uniform_quantize_tensor as uqt

ALGORITHM_KEY = "GPTQ"
_Gran = qtyping.QuantGranularity


def calibrate(tfl_op, graph_info: qtyping.GraphInfo, tensor_content_map: MutableMapping,
              inputs_to_ignore: Sequence[int] | None = None,
              outputs_to_ignore: Sequence[int] | None = None,
              valid_range: tuple[float, float] = (-3e38, 3e38),
              keep_on_device: bool = False, **kwargs) -> dict:
  """{tensor name: {min, max, num_samples, hessian}}; hessian = (2 / num_samples) * X^T X.

  `keep_on_device=True` leaves the float64 Hessian as a CUDA tensor (no K x K download per
  batch); `qsv_utils.gptq_and_moving_average_update` and `get_tensor_quant_params` accept both.
  """
  del kwargs
  from ... import device
  op_qsvs = {}
  ids = common_quantize.get_tensor_indices_requiring_calibration(
      tfl_op, graph_info, inputs_to_ignore, outputs_to_ignore)
  for name, content, qsv in common_quantize.collect_activation_statistics_batch(
      ids, graph_info, tensor_content_map, valid_range[0], valid_range[1]):
    x = hostio.to_device(content, np.float32)
    h = device.xtx(x, 2.0 / float(qsv["num_samples"]))
    qsv["hessian"] = h if keep_on_device else hostio.to_host(h)
    op_qsvs[name] = qsv
  return op_qsvs


def _prepare_hessian_inverse(hessian, damp_factor: float = 0.01):
  """NumPy-facing mirror: float32 inverse; leaves the damped diagonal in `hessian` like the reference."""
  from ... import device
  import torch
  if isinstance(hessian, torch.Tensor):
    h_dev = hessian if hessian.dtype == torch.float64 else hessian.double()
    return device.hessian_inverse(h_dev, damp_factor, keep_damped_diagonal=h_dev is hessian)
  h_np = np.asarray(hessian)
  h_dev = hostio.to_device(h_np.astype(np.float64, copy=False))
  hinv = device.hessian_inverse(h_dev, damp_factor, keep_damped_diagonal=True)
  if h_np.flags.writeable:  # np.diag view quirk, gptq.py:114,123
    np.fill_diagonal(h_np, hostio.to_host(torch.diagonal(h_dev)))
  return hostio.to_host(hinv)


def _scale_zp_device(x2, op_info, cfg, shape):
  """(scale, zp) device tensors from min/max at the configured granularity, no clipping."""
  from ... import device
  gran = cfg.granularity
  block = uqt.extract_block_size_from_granularity(gran)
  qdim = common_utils.get_weight_quantized_dim(op_info, np.empty(shape, np.bool_), gran)
  if block:
    uqt._blockwise_shape(shape, qdim, block)
    mn, mx = device.minmax_blocks(x2, block)
  elif gran == _Gran.CHANNELWISE and qdim == 0:
    mn, mx, _ = device.row_stats(x2)
  elif gran in (_Gran.TENSORWISE, _Gran.CHANNELWISE) and qdim is None:
    mm = device.minmax_tensor(x2)
    mn, mx = mm[0:1], mm[1:2]
  else:
    raise NotImplementedError(f"GPTQ along quantised dimension {qdim} is not on the accelerated path yet")
  zp, scale, _ = device.scale_zp_from_minmax(mn, mx, cfg.num_bits, bool(cfg.symmetric), bool(block))
  return scale, zp, qdim, block


def get_tensor_quant_params(
    op_info: qtyping.OpInfo,
    tensor_quant_config: qtyping.TensorQuantizationConfig,
    tensor_content: Optional[np.ndarray] = None,
    tensor_qsv: Optional[Mapping[str, Any]] = None,
) -> qtyping.UniformQuantParams:
  cfg = tensor_quant_config
  act_qsv = tensor_qsv.get("activation_tensor_qsv") if tensor_qsv else None
  have_minmax = tensor_qsv is not None and "min" in tensor_qsv
  if not have_minmax and tensor_content is None:
    raise ValueError(
        f"{op_info.op_name}(index: {op_info.subgraph_op_index}) not found in"
        " tensor_name_to_qsv. Check if the correct calibration results are"
        " passed into the ParamsGenerator.")
  if have_minmax and "max" not in tensor_qsv:
    raise ValueError(
        "min and max must be provided to produce tensor quantization"
        " parameters. Check if the correct calibration results are passed into"
        " the ParamsGenerator.")
  block = uqt.extract_block_size_from_granularity(cfg.granularity)
  if have_minmax:  # calibrated tensors (activations, or weights with a QSV)
    zp, scale = uqt.tensor_zp_scale_from_min_max(
        tensor_qsv["min"], tensor_qsv["max"], cfg.num_bits, cfg.symmetric, cfg.granularity, None)
    qdim = common_utils.get_weight_quantized_dim(op_info, tensor_content, cfg.granularity)
    params = qtyping.UniformQuantParams(
        scale=scale, zero_point=zp, num_bits=cfg.num_bits, symmetric=cfg.symmetric,
        quantized_dimension=qdim, block_size=block)
    if tensor_content is None or act_qsv is None or "hessian" not in act_qsv:
      return params  # the reference returns no quantized_data here either (gptq.py:290-294)
    x2 = hostio.to_device(tensor_content.reshape(-1, tensor_content.shape[-1])
                          if block or qdim is None else
                          tensor_content.reshape(tensor_content.shape[0], -1), np.float32)
    scale_d = hostio.to_device(np.asarray(scale, np.float32).reshape(-1))
    zp_d = hostio.to_device(np.asarray(zp).astype(np.int32).reshape(-1))
  else:
    if tensor_content.dtype != np.float32:
      raise ValueError(f"only float32 weights are quantised, got {tensor_content.dtype}")
    shape = tensor_content.shape
    x_dev = hostio.to_device(tensor_content, np.float32)
    gran = cfg.granularity
    qd = common_utils.get_weight_quantized_dim(op_info, tensor_content, gran)
    x2 = (x_dev.reshape(-1, shape[-1]) if block else
          x_dev.reshape(1, -1) if qd is None else x_dev.reshape(shape[0], -1))
    scale_d, zp_d, qdim, block = _scale_zp_device(x2, op_info, cfg, shape)
    pshape = (list(shape[:-1]) + [shape[-1] // block] if block else
              [1] * len(shape) if qdim is None else [shape[0]] + [1] * (len(shape) - 1))
    params = qtyping.UniformQuantParams(
        scale=hostio.to_host(scale_d).reshape(pshape),
        zero_point=hostio.to_host(zp_d).reshape(pshape).astype(uqt.numpy_dtype_for(cfg.num_bits)),
        num_bits=cfg.num_bits, symmetric=cfg.symmetric, quantized_dimension=qdim, block_size=block)
    if act_qsv is None or "hessian" not in act_qsv:
      return params
  return _apply_gptq_device(tensor_content, x2, params, scale_d, zp_d, act_qsv, cfg)


def _apply_gptq_device(tensor_content, x2, params, scale_d, zp_d, act_qsv, cfg):
  from ... import device
  if cfg.num_bits > 8:
    raise ValueError(f"device GPTQ supports num_bits <= 8, got {cfg.num_bits}")
  if tensor_content.ndim != 2:
    raise ValueError("GPTQ expects a 2-D FULLY_CONNECTED weight")
  rows, k = tensor_content.shape
  hinv = _prepare_hessian_inverse(act_qsv["hessian"])
  import torch
  if not isinstance(hinv, torch.Tensor):
    hinv = hostio.to_device(hinv, np.float32)
  if hinv.shape[0] != k:
    raise ValueError(f"Hessian order {hinv.shape[0]} does not match the weight's {k} input features")
  q = device.gptq_quantize(x2.reshape(rows, k), hinv, scale_d.reshape(-1), zp_d.reshape(-1),
                           params.block_size, cfg.num_bits, bool(cfg.symmetric))
  return dataclasses.replace(params, quantized_data=hostio.to_host(q).reshape(tensor_content.shape))
