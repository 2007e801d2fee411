"""`OSCAR`: activation-aware channel scaling + exact optimal clipping (FULLY_CONNECTED only).

Mirror of ai_edge_quantizer/algorithms/uniform_quantize/oscar.py (`_compute_channel_scales`
:192-251, `calibrate` :254-305, `get_clip_bounds` :308-364, `get_tensor_quant_params`
:461-526, `materialize_fully_connected` :635-682).

Split of the work: everything O(rows x cols) runs in float64 on the device (csrc/oscar.cu):
column second moments, the objective / arg-max passes over W, the per-group sort +
breakpoint scan, scale and quantisation of the scaled weight.  What stays here is the O(cols)
vector algebra between passes (geometric-mean normalisation, clamping, the damped
fixed-point update), written with the reference's NumPy expressions so that those vectors
are bit-identical given the same reductions.
"""
from __future__ import annotations

import dataclasses
import logging
from typing import Any, Optional

import numpy as np

from ... import hostio
from ... import qtyping
from ...utils import tfl_flatbuffer_utils
from ..utils import common_utils
from . import common_quantize
from . import naive_min_max_quantize
from . import uniform_quantize_tensor as uqt

ALGORITHM_KEY = "OSCAR"
_Op = qtyping.TFLOperationName
_QT = qtyping.QuantTransformation
_Gran = qtyping.QuantGranularity

_EPS = 1e-12
_SCALE_CLAMP = (1e-4, 1e4)


def _floor_positive(mu2: np.ndarray) -> np.ndarray:
  mu2 = np.asarray(mu2, np.float64)
  return np.maximum(mu2, float(np.max(mu2)) * 1e-8 + _EPS)


def _dev64(a: np.ndarray):
  import torch
  return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(hostio.device())


def _objective(w_dev, s: np.ndarray, mu2: np.ndarray, group: int, want_a_eff: bool):
  """(objective value, a_eff or None) for channel scales s (reference :172-189, :223-230)."""
  from ... import device
  group_sq, a_eff = device.oscar_pass(w_dev, _dev64(s), group, want_a_eff)
  group_sq = hostio.to_host(group_sq)
  m = mu2 / (s * s)
  total = 0.0
  for gi in range(group_sq.size):  # the reference accumulates group by group in this order
    total += float(group_sq[gi]) * float(m[gi * group:(gi + 1) * group].sum())
  return total, (None if a_eff is None else hostio.to_host(a_eff))


def compute_channel_scales(w_dev, mu2: np.ndarray, block_size: int = 0, num_iters: int = 3):
  """(s or None, gain): per-input-channel scales of `_compute_channel_scales` (:192-251)."""
  from ... import device
  in_ch = mu2.size
  mu2 = _floor_positive(mu2)
  mu = np.sqrt(mu2)
  d = w_dev.shape[1]
  group = block_size if (block_size and d % block_size == 0) else d

  def normalized(v):
    v = v / np.exp(np.mean(np.log(v)))
    return np.clip(v, *_SCALE_CLAMP)

  a_base = hostio.to_host(device.colsq(w_dev, 1.0)) + _EPS
  identity_loss, _ = _objective(w_dev, np.ones(in_ch), mu2, group, False)
  s = normalized(np.sqrt(mu / np.sqrt(a_base)))
  loss, a_eff = _objective(w_dev, s, mu2, group, True)
  best = (loss, s)
  for it in range(num_iters):
    a_eff = np.maximum(a_eff, 0.25 * a_base)
    s_cand = normalized(np.sqrt(mu / np.sqrt(a_eff)))
    s = normalized(np.sqrt(s * s_cand))
    loss, a_eff = _objective(w_dev, s, mu2, group, it + 1 < num_iters)
    if loss < best[0]:
      best = (loss, s)
  if best[0] >= identity_loss:
    return None, 1.0
  return best[1], identity_loss / max(best[0], _EPS)


def calibrate(tfl_op, graph_info: qtyping.GraphInfo, tensor_content_map, inputs_to_ignore=None,
              outputs_to_ignore=None, valid_range: tuple[float, float] = (-3e38, 3e38), **kwargs):
  """min / max / num_samples plus mu2 = mean(x * x) per trailing-axis channel (:254-305)."""
  del kwargs
  from ... import device
  ids = common_quantize.get_tensor_indices_requiring_calibration(
      tfl_op, graph_info, inputs_to_ignore, outputs_to_ignore)
  out = {}
  for name, content, qsv in common_quantize.collect_activation_statistics_batch(
      ids, graph_info, tensor_content_map, valid_range[0], valid_range[1]):
    x = hostio.to_device(content, np.float32)
    rows = max(x.numel() // x.shape[-1], 1)
    qsv["mu2"] = hostio.to_host(device.colsq(x, 1.0 / rows))
    out[name] = qsv
  return out


def _extract_mu2(tensor_qsv):
  if not tensor_qsv:
    return None
  if "mu2" in tensor_qsv:
    return tensor_qsv["mu2"]
  if "activation_tensor_qsv" in tensor_qsv and tensor_qsv["activation_tensor_qsv"]:
    return tensor_qsv["activation_tensor_qsv"].get("mu2")
  return None


def _clip_bounds(w_dev, s: np.ndarray, mu2_scaled, num_bits: int, granularity, block_size: int):
  """Device bounds of `get_clip_bounds` (:308-364), flat, one per group; and the group length."""
  from ... import device
  n, d = w_dev.shape
  qmax = 2 ** (num_bits - 1) - 1
  col_mu2 = np.ones(d) if mu2_scaled is None else np.asarray(mu2_scaled, np.float64).ravel()
  col_mu2 = _floor_positive(col_mu2)
  s_dev, m_dev = _dev64(s), _dev64(col_mu2)
  if granularity == _Gran.TENSORWISE:
    mass = float(np.tile(col_mu2, n).sum()) + _EPS
    return device.oscar_clip(w_dev, s_dev, m_dev, n * d, qmax, mass0=mass), n * d
  if granularity == _Gran.CHANNELWISE:
    mass = float(col_mu2.sum()) + _EPS
    return device.oscar_clip(w_dev, s_dev, m_dev, d, qmax, mass0=mass), d
  if block_size:
    if d % block_size != 0:
      raise ValueError(
          f"Block size {block_size} must divide the reduction dimension "
          f"{d} of {_Op.FULLY_CONNECTED} weights with shape {(n, d)}.")
    masses = np.array([float(col_mu2[b * block_size:(b + 1) * block_size].sum()) + _EPS
                       for b in range(d // block_size)])
    return device.oscar_clip(w_dev, s_dev, m_dev, block_size, qmax, mass=_dev64(masses)), block_size
  raise ValueError(f"Unsupported granularity: {granularity}")


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
  if not cfg.symmetric:
    raise ValueError(
        "OSCAR supports symmetric weight quantization only, got asymmetric"
        f" config for op {op_info.op_name}.")
  if op_info.op_name != _Op.FULLY_CONNECTED:
    raise ValueError(f"OSCAR supports FULLY_CONNECTED only, got: {op_info.op_name}")
  if tensor_content.ndim != 2:
    raise ValueError(f"OSCAR expects 2-D weights for {op_info.op_name}, got {tensor_content.shape}")
  if tensor_content.dtype != np.float32:
    raise ValueError(f"only float32 weights are quantised, got {tensor_content.dtype}")
  n, d = tensor_content.shape
  gran = cfg.granularity
  block_size = uqt.extract_block_size_from_granularity(gran)
  mu2 = _extract_mu2(tensor_qsv)
  w_dev = hostio.to_device(tensor_content, np.float32)
  if mu2 is not None:
    mu2_arr = np.asarray(mu2, np.float64).ravel()
    if mu2_arr.size != d:
      raise ValueError(
          f"OSCAR: activation mu2 has {mu2_arr.size} channels but"
          f" {op_info.op_name} weights of shape {(n, d)} expect {d}.")
    s, _ = compute_channel_scales(w_dev, mu2_arr, block_size)
    if s is None:
      s = np.ones(d, dtype=np.float64)
    mu2_scaled = mu2_arr / (s * s)
  else:
    logging.warning(
        "OSCAR: no activation second moments (mu2) found for op %s"
        " (index %d); falling back to unscaled optimal clipping.",
        op_info.op_name, op_info.subgraph_op_index)
    s = np.ones(d, dtype=np.float64)
    mu2_scaled = None
  qmax = 2 ** (cfg.num_bits - 1) - 1
  bounds, glen = _clip_bounds(w_dev, s, mu2_scaled, cfg.num_bits, gran, block_size)
  scale_d = device.oscar_scale(bounds, qmax, bool(block_size))
  q = device.oscar_quantize(w_dev, _dev64(s), scale_d, glen, cfg.num_bits)
  scale = hostio.to_host(scale_d)
  if block_size:
    scale = scale.astype(np.float32).reshape(n, d // block_size)
  elif gran == _Gran.CHANNELWISE:
    scale = scale.reshape(n, 1)
  else:
    scale = scale.reshape(1, 1)
  qdim = common_utils.get_weight_quantized_dim(op_info, tensor_content, gran)
  return qtyping.UniformQuantParams(
      scale=scale, zero_point=np.zeros(scale.shape, dtype=uqt.numpy_dtype_for(cfg.num_bits)),
      num_bits=cfg.num_bits, symmetric=cfg.symmetric, quantized_dimension=qdim,
      block_size=block_size, quantized_data=hostio.to_host(q),
      custom_algorithm_param={"multiplier": (1.0 / s).astype(np.float32)})


def materialize_fully_connected(op_info: qtyping.OpInfo, graph_info: qtyping.GraphInfo,
                                tensor_name_to_qsv=None,
                                tensor_quant_params_cache: common_utils.TensorQuantParamsCache = None):
  """INSERT_MULTIPLY on the activation, QUANTIZE_TENSOR on the scaled weight (:635-682)."""
  wcfg = op_info.op_quant_config.weight_tensor_config
  if wcfg is None:
    raise ValueError("Weight tensor quantization config is not provided for OSCAR quantization.")
  if op_info.op_name != _Op.FULLY_CONNECTED:
    raise ValueError(f"OSCAR supports FULLY_CONNECTED only, got: {op_info.op_name}")
  tensors = graph_info.subgraph_tensors
  name = lambda t: tfl_flatbuffer_utils.get_tensor_name(t)
  input_tensor, weight_tensor = tensors[op_info.op.inputs[0]], tensors[op_info.op.inputs[1]]
  mu2 = None
  if tensor_name_to_qsv and name(input_tensor) in tensor_name_to_qsv:
    mu2 = tensor_name_to_qsv[name(input_tensor)].get("mu2")
  params = tensor_quant_params_cache.lookup(weight_tensor.buffer, wcfg)
  if not params:
    data = tfl_flatbuffer_utils.get_tensor_data(weight_tensor, graph_info.buffers)
    params = get_tensor_quant_params(op_info, wcfg, data, tensor_qsv={"mu2": mu2})
    tensor_quant_params_cache.insert(weight_tensor.buffer, wcfg, params)
  sid = op_info.subgraph_op_index
  out = [
      qtyping.TensorTransformationParams(name(input_tensor), consumers=[
          qtyping.OpToTensorParams(sid, [_QT.INSERT_MULTIPLY], params)]),
      qtyping.TensorTransformationParams(name(weight_tensor), consumers=[
          qtyping.OpToTensorParams(sid, [_QT.QUANTIZE_TENSOR], params)]),
  ]
  if len(op_info.op.inputs) > 2 and op_info.op.inputs[2] >= 0:
    out.append(qtyping.TensorTransformationParams(name(tensors[op_info.op.inputs[2]]), consumers=[
        qtyping.OpToTensorParams(sid, [_QT.NO_QUANTIZE])]))
  out.append(qtyping.TensorTransformationParams(
      name(tensors[op_info.op.outputs[0]]), producer=qtyping.OpToTensorParams(sid, [_QT.NO_QUANTIZE])))
  return out
