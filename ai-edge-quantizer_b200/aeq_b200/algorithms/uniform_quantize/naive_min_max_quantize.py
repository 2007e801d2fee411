"""`min_max_uniform_quantize`: weight min/max -> scale -> quantise; activation min/max calibration.

Mirror of ai_edge_quantizer/algorithms/uniform_quantize/naive_min_max_quantize.py
(`get_tensor_quant_params` :34-110, `min_max_calibrate` :181-226) with the
arithmetic in the fused sm_100a kernels:
  CHANNELWISE (quantised dim 0) -> aeqb_requant_rows_f32        (one HBM pass)
  CHANNELWISE (other dims: DEPTHWISE_CONV_2D dim 3, BATCH_MATMUL) -> aeqb_swap_axes moves the
                                   channel axis to the front, same kernel, integers moved back
  BLOCKWISE_*                   -> aeqb_requant_blocks_f32      (one HBM pass)
  TENSORWISE / QSV min-max      -> aeqb_minmax_tensor_f32 + aeqb_requant_given_minmax_f32
"""
from __future__ import annotations

import dataclasses
from typing import Any, Optional

import numpy as np

from ... import host
from ... import hostio
from ... import qtyping
from ...transformations import quantize_tensor
from ..utils import common_utils
from . import common_quantize
from . import uniform_quantize_tensor as uqt

ALGORITHM_KEY = "min_max_uniform_quantize"
_Gran = qtyping.QuantGranularity


def quantize_weight(op_info: qtyping.OpInfo, cfg: qtyping.TensorQuantizationConfig,
                    tensor_content: np.ndarray, clip=None, x_dev=None) -> qtyping.UniformQuantParams:
  """Fused min/max -> scale -> quantise of a constant tensor.

  `clip` is a device tensor of clipping constants (OCTAV); `x_dev` an optional
  device copy of `tensor_content` (same element order) that is already resident.
  """
  from ... import device
  if tensor_content.dtype != np.float32:
    raise ValueError(f"only float32 weights are quantised, got {tensor_content.dtype}")
  gran = cfg.granularity
  bits, sym = cfg.num_bits, bool(cfg.symmetric)
  qdim = common_utils.get_weight_quantized_dim(op_info, tensor_content, gran)
  block = uqt.extract_block_size_from_granularity(gran)
  shape = tensor_content.shape
  if block:
    if not sym:
      raise ValueError("blockwise quantisation is symmetric only")
    if qdim != tensor_content.ndim - 1:
      raise ValueError("blockwise quantisation cuts the last axis")
    uqt._blockwise_shape(shape, qdim, block)  # reference's divisibility error
    w2 = tensor_content.reshape(-1, shape[-1])
    if clip is None:  # host-buffer pipeline: chunked H2D -> fused kernel -> D2H
      q, packed, scale, _ = host.requant_blocks([w2], block, bits, want_packed=(bits == 4))[0]
      if packed is not None:  # QUANTIZE_TENSOR's pack step finds the bytes instead of redoing them
        quantize_tensor.remember_packed(q, bits, packed)
      scale = scale.reshape(*shape[:-1], shape[-1] // block)
      return qtyping.UniformQuantParams(
          num_bits=bits, quantized_dimension=qdim, scale=scale,
          zero_point=np.zeros(scale.shape, dtype=uqt.numpy_dtype_for(bits)), symmetric=sym,
          quantized_data=q.reshape(shape), block_size=block)
    xd = hostio.to_device(w2, np.float32) if x_dev is None else x_dev.reshape(w2.shape)
    out = device.requant_blocks(xd, block, bits, clip=clip, want_scale_f16=False)
    scale = hostio.to_host(out.scale).reshape(*shape[:-1], shape[-1] // block)
    zp = np.zeros(scale.shape, dtype=uqt.numpy_dtype_for(bits))
  elif gran == _Gran.CHANNELWISE and qdim is not None:
    pshape = [1] * tensor_content.ndim
    pshape[qdim] = shape[qdim]
    if clip is None and qdim == 0:
      w2 = tensor_content.reshape(shape[0], -1)
      fuse_pack = bits in (2, 4) and w2.shape[1] % (8 // bits) == 0
      q, packed, scale, zp = host.requant_rows([w2], bits, sym, want_packed=fuse_pack)[0]
      if packed is not None:
        quantize_tensor.remember_packed(q, bits, packed)
      return qtyping.UniformQuantParams(
          num_bits=bits, quantized_dimension=qdim, scale=scale.reshape(pshape),
          zero_point=zp.reshape(pshape).astype(uqt.numpy_dtype_for(bits)), symmetric=sym,
          quantized_data=q.reshape(shape), block_size=0)
    xd = hostio.to_device(tensor_content, np.float32) if x_dev is None else x_dev
    out = device.requant_rows(device.channel_rows(xd, shape, qdim), bits, sym, clip=clip)
    scale = hostio.to_host(out.scale).reshape(pshape)
    zp = hostio.to_host(out.zero_point).reshape(pshape).astype(uqt.numpy_dtype_for(bits))
    return qtyping.UniformQuantParams(
        num_bits=bits, quantized_dimension=qdim, scale=scale, zero_point=zp, symmetric=sym,
        quantized_data=hostio.to_host(device.channel_rows_back(out.q, shape, qdim)), block_size=0)
  elif gran in (_Gran.TENSORWISE, _Gran.CHANNELWISE):
    # CHANNELWISE on an op without a quantised-dim entry reduces over everything,
    # like get_reduce_dims(None) -> axis=None in the reference.
    x = (hostio.to_device(tensor_content.reshape(1, -1), np.float32) if x_dev is None
         else x_dev.reshape(1, -1))
    mm = device.minmax_tensor(x)
    out = device.requant_given_minmax(x, mm[0:1], mm[1:2], bits, sym, per_row=False, clip=clip)
    pshape = [1] * tensor_content.ndim
    scale = hostio.to_host(out.scale).reshape(pshape)
    zp = hostio.to_host(out.zero_point).reshape(pshape).astype(uqt.numpy_dtype_for(bits))
  else:
    raise ValueError(f"Unsupported granularity: {gran}")
  return qtyping.UniformQuantParams(
      num_bits=bits, quantized_dimension=qdim, scale=scale, zero_point=zp, symmetric=sym,
      quantized_data=hostio.to_host(out.q).reshape(shape), block_size=block)


def get_tensor_quant_params(
    op_info: qtyping.OpInfo,
    tensor_quant_config: qtyping.TensorQuantizationConfig,
    tensor_content: Optional[np.ndarray] = None,
    tensor_qsv: Optional[dict[str, Any]] = None,
) -> qtyping.UniformQuantParams:
  """Quantisation parameters (and quantised data for constants) of one tensor."""
  cfg = tensor_quant_config
  if tensor_qsv is None or "min" not in tensor_qsv:
    if tensor_content is None:
      raise ValueError(
          f"{op_info.op_name}(index: {op_info.subgraph_op_index}) not found in"
          " tensor_name_to_qsv. Check if the correct calibration results are"
          " passed into the ParamsGenerator.")
    return quantize_weight(op_info, cfg, tensor_content)

  if "min" not in tensor_qsv or "max" not in tensor_qsv:
    raise ValueError(
        "min and max must be provided to produce tensor quantization"
        " parameters. Check if the correct calibration results are passed into"
        " the ParamsGenerator.")
  zp, scale = uqt.tensor_zp_scale_from_min_max(
      tensor_qsv["min"], tensor_qsv["max"], cfg.num_bits, cfg.symmetric, cfg.granularity, None)
  qdim = common_utils.get_weight_quantized_dim(op_info, tensor_content, cfg.granularity)
  params = qtyping.UniformQuantParams(
      num_bits=cfg.num_bits, quantized_dimension=qdim, scale=scale, zero_point=zp,
      symmetric=cfg.symmetric,
      block_size=uqt.extract_block_size_from_granularity(cfg.granularity))
  if tensor_content is None:
    return params
  quantized = uqt.uniform_quantize(tensor_content, params, uqt.is_blockwise(cfg.granularity))
  return dataclasses.replace(params, quantized_data=quantized)


def check_if_quantized(tensor: Any) -> bool:
  return tensor.quantization is not None and tensor.quantization.scale is not None


def min_max_calibrate(tfl_op, graph_info: qtyping.GraphInfo, tensor_content_map,
                      inputs_to_ignore=None, outputs_to_ignore=None,
                      valid_range: tuple[float, float] = (-3e38, 3e38), **kwargs) -> dict:
  """{tensor name: {min, max, num_samples}} for every runtime tensor the op touches.

  Values outside the open interval `valid_range` are ignored (bf16 -inf padding
  constants etc.); one batched device launch for all the op's tensors.
  """
  del kwargs
  ids = common_quantize.get_tensor_indices_requiring_calibration(
      tfl_op, graph_info, inputs_to_ignore, outputs_to_ignore)
  return {name: qsv for name, _, qsv in common_quantize.collect_activation_statistics_batch(
      ids, graph_info, tensor_content_map, valid_range[0], valid_range[1])}
