"""Layout helpers shared by the algorithms (subset of the reference's
ai_edge_quantizer/algorithms/utils/common_utils.py: TensorQuantParamsCache :48-77,
get_weight_quantized_dim :1162-1193, get_reduce_dims :1196-1207,
get_bmm_weight_quantized_dim :1210-1218)."""
from __future__ import annotations

from typing import Optional, Sequence

from ... import qtyping
from ...utils import tfl_flatbuffer_utils

_Op = qtyping.TFLOperationName


def is_blockwise(granularity) -> bool:
  return "BLOCKWISE" in str(granularity)


class TensorQuantParamsCache:
  """(buffer id, TensorQuantizationConfig) -> UniformQuantParams, shared between ops."""

  def __init__(self):
    self._cache: dict = {}

  def lookup(self, buffer_id: int, quant_config: qtyping.TensorQuantizationConfig):
    return self._cache.get((buffer_id, quant_config), None)

  def insert(self, buffer_id: int, quant_config: qtyping.TensorQuantizationConfig, quant_params):
    self._cache[(buffer_id, quant_config)] = quant_params
    return quant_params

  def __len__(self):
    return len(self._cache)


def get_bmm_weight_quantized_dim(weight_tensor_data, adj_y: bool) -> int:
  rank = len(weight_tensor_data.shape)
  return rank - 2 if adj_y else rank - 1


def get_weight_quantized_dim(op_info: qtyping.OpInfo, tensor_data, granularity) -> Optional[int]:
  if granularity == qtyping.QuantGranularity.CHANNELWISE:
    if op_info.op_name == _Op.BATCH_MATMUL:
      return get_bmm_weight_quantized_dim(tensor_data, adj_y=op_info.op.builtinOptions.adjY)
    return tfl_flatbuffer_utils.TFL_OP_TO_WEIGHT_QUANTIZED_DIM.get(op_info.op_name, None)
  if is_blockwise(granularity):
    return tfl_flatbuffer_utils.TFL_OP_TO_BLOCKWISE_WEIGHT_QUANTIZED_DIM[op_info.op_name]
  return None


def get_reduce_dims(quantized_dim: Optional[int], tensor_shape: Sequence[int]):
  if quantized_dim is None:
    return None
  return tuple(d for d in range(len(tensor_shape)) if d != quantized_dim)


# ------------------------------------------------------------------ per-op bookkeeping
# Compact restatement of the reference's caller chain for the ops whose constant
# weights feed the hot path (materialize_standard_op :878-1065,
# _get_tensor_transformation_params_wrapper :219-290, get_tensor_transformations
# :1068-1121, _get_tensor_qsv_val :182-216).  Scale constraints between tensors,
# bias quantisation and the other ~45 op materialisers are the reference's own
# bookkeeping and are reached through `aeq_b200.plugin.install` instead.
_QT = qtyping.QuantTransformation
_DRQ_OR_WEIGHT_ONLY_OPS = frozenset([
    _Op.FULLY_CONNECTED, _Op.CONV_2D, _Op.BATCH_MATMUL, _Op.EMBEDDING_LOOKUP,
    _Op.DEPTHWISE_CONV_2D, _Op.CONV_2D_TRANSPOSE])


def get_tensor_transformations(op_quant_config: qtyping.OpQuantizationConfig,
                               is_inbounding_tensor: bool, is_constant: bool):
  integer = op_quant_config.compute_precision == qtyping.ComputePrecision.INTEGER
  if integer and op_quant_config.activation_tensor_config is not None:  # SRQ
    if not is_inbounding_tensor:
      return [_QT.ADD_DEQUANTIZE]
    return [_QT.QUANTIZE_TENSOR] if is_constant else [_QT.ADD_QUANTIZE]
  if integer:  # DRQ
    return [_QT.QUANTIZE_TENSOR] if (is_inbounding_tensor and is_constant) else [_QT.NO_QUANTIZE]
  if (op_quant_config.compute_precision == qtyping.ComputePrecision.FLOAT
      and op_quant_config.explicit_dequantize):  # weight-only
    return [_QT.ADD_DEQUANTIZE] if (is_inbounding_tensor and is_constant) else [_QT.NO_QUANTIZE]
  raise ValueError("Unsupported compute precision: %s" % op_quant_config.compute_precision)


def _tensor_qsv(tensor_name, op_info, graph_info, tensor_name_to_qsv):
  """The tensor's QSV with the op's input-0 QSV nested as `activation_tensor_qsv`."""
  qsv = tensor_name_to_qsv.get(tensor_name)
  if op_info.op and len(op_info.op.inputs):
    act = tfl_flatbuffer_utils.get_tensor_name(graph_info.subgraph_tensors[op_info.op.inputs[0]])
    act_qsv = tensor_name_to_qsv.get(act)
    if act_qsv is not None:
      qsv = dict(qsv or {})
      qsv["activation_tensor_qsv"] = act_qsv
  return qsv


def materialize_weight_op(get_tensor_quant_params_fn, op_info: qtyping.OpInfo,
                          graph_info: qtyping.GraphInfo, tensor_name_to_qsv,
                          tensor_quant_params_cache: TensorQuantParamsCache,
                          inputs_to_ignore=()):
  """TensorTransformationParams for every tensor of a weight-carrying op.

  Constants use the weight config (and the (buffer, config) cache); runtime
  tensors use the activation config when there is one (SRQ) and pass through
  otherwise.  Weights smaller than `min_weight_elements` stay float.
  """
  cfg = op_info.op_quant_config
  out = []
  for inbound, ids in ((True, op_info.op.inputs), (False, op_info.op.outputs)):
    for pos, tid in enumerate(ids):
      if tid == -1:
        continue
      tensor = graph_info.subgraph_tensors[tid]
      name = tfl_flatbuffer_utils.get_tensor_name(tensor)
      data = tfl_flatbuffer_utils.get_tensor_data(tensor, graph_info.buffers)
      constant = data is not None
      tcfg = cfg.activation_tensor_config
      # Only the weight operand (input 1 of FC / CONV / TRANSPOSE_CONV / EMBEDDING_LOOKUP) takes
      # the weight config; a constant bias stays float unless activations are quantised too
      # (common_quantize.materialize_fc_conv: bias is handled by the SRQ path only).
      if constant and inbound and pos == 1 and op_info.op_name in _DRQ_OR_WEIGHT_ONLY_OPS:
        tcfg = cfg.weight_tensor_config
      elif constant:
        if tcfg is not None:
          raise NotImplementedError(
              f"static-range quantisation of the constant '{name}' (bias: int32 with scale ="
              " input scale x weight scale) is part of the reference's materialize_fc_conv; use"
              " aeq_b200.plugin.install with the reference for SRQ recipes")
        tcfg = None
      skip = (inbound and pos in inputs_to_ignore) or tcfg is None or (
          constant and (data.dtype != "float32" or data.size < cfg.min_weight_elements))
      params, transformations = None, [_QT.NO_QUANTIZE]
      if not skip:
        params = tensor_quant_params_cache.lookup(tensor.buffer, tcfg) if constant else None
        if params is None:
          try:
            params = get_tensor_quant_params_fn(
                op_info, tcfg, data, _tensor_qsv(name, op_info, graph_info, tensor_name_to_qsv))
          except Exception as e:
            raise ValueError(
                f"Failed to get quantization parameters for tensor: {name}. Error: {e}") from e
          if constant:
            tensor_quant_params_cache.insert(tensor.buffer, tcfg, params)
        transformations = get_tensor_transformations(cfg, inbound, constant)
      o2t = qtyping.OpToTensorParams(op_info.subgraph_op_index, transformations, params)
      out.append(qtyping.TensorTransformationParams(name, consumers=[o2t]) if inbound
                 else qtyping.TensorTransformationParams(name, producer=o2t))
  return out
