"""`float_casting`: fp16 weights behind a DEQUANTIZE op.

Mirror of ai_edge_quantizer/algorithms/nonlinear_quantize/float_casting.py
(`check_op_quantization_config` :44-103, `materialize_fc_conv` :106-197,
`materialize_embedding_lookup` :200-260): the weight becomes
`NonLinearQuantParams(num_bits=16, quantized_data=weight.astype(float16))` with
ADD_DEQUANTIZE, every other tensor NO_QUANTIZE.  The cast runs on the device
(`aeqb_cast_f32_f16`, round to nearest even like NumPy's astype).
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from ... import hostio
from ... import qtyping
from ...utils import tfl_flatbuffer_utils
from ..utils import common_utils

ALGORITHM_KEY = "float_casting"
_Op = qtyping.TFLOperationName
_QT = qtyping.QuantTransformation

_FP16_QUANT_CONFIG = qtyping.TensorQuantizationConfig(num_bits=16, dtype=qtyping.TensorDataType.FLOAT)

SUPPORTED_WEIGHT_QUANT_OPS = frozenset([
    _Op.FULLY_CONNECTED, _Op.CONV_2D, _Op.DEPTHWISE_CONV_2D, _Op.CONV_2D_TRANSPOSE,
    _Op.EMBEDDING_LOOKUP])

# position of the constant weight among the op's inputs
_WEIGHT_INPUT = {_Op.FULLY_CONNECTED: 1, _Op.CONV_2D: 1, _Op.DEPTHWISE_CONV_2D: 1,
                 _Op.CONV_2D_TRANSPOSE: 1, _Op.EMBEDDING_LOOKUP: 1}


def check_op_quantization_config(op_name, op_quant_config, config_check_policy=None) -> None:
  if config_check_policy is not None and config_check_policy:
    raise ValueError(f"Config check isn't implemented yet for op: {op_name}.")
  if op_quant_config.compute_precision != qtyping.ComputePrecision.FLOAT:
    raise ValueError(
        "Currently, only Weight-Only is supported for float casting"
        " quantization. Got unsupported execution mode:"
        f" {op_quant_config.compute_precision} for op: {op_name}")
  if op_quant_config.activation_tensor_config is not None:
    raise ValueError(
        "Activation tensor quantization is not supported for float casting quantization.")
  if op_name not in SUPPORTED_WEIGHT_QUANT_OPS:
    raise ValueError(f"Unsupported op: {op_name} for float casting quantization.")
  w = op_quant_config.weight_tensor_config
  if w is None:
    raise ValueError(
        "Weight tensor quantization config is required for float casting quantization.")
  if w.num_bits != 16 or w.dtype != qtyping.TensorDataType.FLOAT:
    raise ValueError(
        "Currently, float casting quantization config requires number of bits"
        f" to be set as 16, dtype as float, got {w.num_bits} and {w.dtype} .")


def cast_weight(weight_content: np.ndarray) -> qtyping.NonLinearQuantParams:
  """NonLinearQuantParams holding the float16 copy of a float32 weight."""
  from ... import device
  if weight_content.dtype != np.float32:
    raise ValueError(f"only float32 weights are cast, got {weight_content.dtype}")
  x = hostio.to_device(weight_content, np.float32)
  return qtyping.NonLinearQuantParams(num_bits=16, quantized_data=hostio.to_host(device.cast_f16(x)))


def _no_quantize(op_info, name: str, inbound: bool) -> qtyping.TensorTransformationParams:
  o2t = qtyping.OpToTensorParams(op_info.subgraph_op_index, [_QT.NO_QUANTIZE])
  return (qtyping.TensorTransformationParams(name, consumers=[o2t]) if inbound
          else qtyping.TensorTransformationParams(name, producer=o2t))


def materialize_weight_op(op_info: qtyping.OpInfo, graph_info: qtyping.GraphInfo,
                          tensor_name_to_qsv: Optional[dict[str, Any]],
                          tensor_quant_params_cache: common_utils.TensorQuantParamsCache):
  """materialize_fc_conv / materialize_embedding_lookup of the reference in one function."""
  del tensor_name_to_qsv
  wpos = _WEIGHT_INPUT[op_info.op_name]
  out = []
  for inbound, ids in ((True, op_info.op.inputs), (False, op_info.op.outputs)):
    for pos, tid in enumerate(ids):
      if tid == -1:
        continue
      tensor = graph_info.subgraph_tensors[tid]
      name = tfl_flatbuffer_utils.get_tensor_name(tensor)
      data = tfl_flatbuffer_utils.get_tensor_data(tensor, graph_info.buffers) if (
          inbound and pos == wpos) else None
      if data is None:
        out.append(_no_quantize(op_info, name, inbound))
        continue
      params = tensor_quant_params_cache.lookup(tensor.buffer, _FP16_QUANT_CONFIG)
      if not params:
        params = cast_weight(data)
        tensor_quant_params_cache.insert(tensor.buffer, _FP16_QUANT_CONFIG, params)
      o2t = qtyping.OpToTensorParams(op_info.subgraph_op_index, [_QT.ADD_DEQUANTIZE], params)
      out.append(qtyping.TensorTransformationParams(name, consumers=[o2t]))
  return out
