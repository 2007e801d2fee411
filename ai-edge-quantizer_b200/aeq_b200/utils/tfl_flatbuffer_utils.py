"""The two weight-layout tables the hot path reads, plus constant-tensor access.

Mirror of ai_edge_quantizer/utils/tfl_flatbuffer_utils.py:95-106 (tables) and
:242-263 (`get_tensor_data`).  Graph objects are duck-typed: a tensor needs
`.buffer`, `.shape`, `.type`, `.name`; a buffer needs `.data` (bytes-like or a
uint8 array) — what the flatbuffer object API and our synthetic graphs both give.
"""
from __future__ import annotations

import numpy as np

from .. import qtyping
from . import tfl_model

_Op = qtyping.TFLOperationName

# Model I/O (reference :118-160 goes through ai_edge_litert's flatbuffer_utils; here the TFL3
# wire format is read and written by aeq_b200.utils.tfl_model / flatbuffer_lite).
read_model = tfl_model.read_model
read_model_from_bytes = tfl_model.read_model_from_bytes
write_model = tfl_model.write_model
write_model_to_bytes = tfl_model.write_model_to_bytes

# Builtin operator code -> op name, for the ops whose constants reach the hot path
# (reference TFL_OP_CODE_TO_NAME, :32-92, restricted to the weight-carrying ops).
TFL_OP_CODE_TO_NAME = qtyping.FrozenParams({
    tfl_model.BuiltinOperator.FULLY_CONNECTED: _Op.FULLY_CONNECTED,
    tfl_model.BuiltinOperator.CONV_2D: _Op.CONV_2D,
    tfl_model.BuiltinOperator.DEPTHWISE_CONV_2D: _Op.DEPTHWISE_CONV_2D,
    tfl_model.BuiltinOperator.EMBEDDING_LOOKUP: _Op.EMBEDDING_LOOKUP,
    tfl_model.BuiltinOperator.TRANSPOSE_CONV: _Op.CONV_2D_TRANSPOSE,
    tfl_model.BuiltinOperator.BATCH_MATMUL: _Op.BATCH_MATMUL,
})


def get_op_scope(op, subgraph_tensors) -> str:
  """Output tensor names joined with ';' — what recipe regexes match (reference :371-415)."""
  def names(ids):
    return [n for n in (get_tensor_name(subgraph_tensors[i]) for i in ids if i != -1) if n]
  found = names(op.outputs) or names(op.inputs)
  return ";".join(found) + (";" if found else "")


# Per-channel quantised dimension of the weight (TFLite quantisation spec).
TFL_OP_TO_WEIGHT_QUANTIZED_DIM = qtyping.FrozenParams({
    _Op.FULLY_CONNECTED: 0,
    _Op.DEPTHWISE_CONV_2D: 3,
    _Op.CONV_2D: 0,
    _Op.EMBEDDING_LOOKUP: 0,
    _Op.CONV_2D_TRANSPOSE: 0,
})

# Axis that is cut into blocks for BLOCKWISE_* granularities.
TFL_OP_TO_BLOCKWISE_WEIGHT_QUANTIZED_DIM = qtyping.FrozenParams({
    _Op.FULLY_CONNECTED: 1,
    _Op.EMBEDDING_LOOKUP: 1,
})

# TensorType codes of the TFLite schema that this path can meet.
TENSOR_TYPE_TO_NUMPY = {0: np.float32, 1: np.float16, 2: np.int32, 3: np.uint8,
                        4: np.int64, 6: np.bool_, 7: np.int16, 9: np.int8, 10: np.float64,
                        12: np.uint64, 15: np.uint32, 16: np.uint16}


def get_tensor_name(tensor) -> str:
  name = tensor.name
  return name.decode("utf-8") if isinstance(name, (bytes, bytearray)) else str(name)


def get_tensor_data(tensor, buffers):
  """Zero-copy NumPy view of a constant tensor, or None for runtime tensors."""
  if tensor.buffer is None or tensor.buffer < 0 or tensor.buffer >= len(buffers):
    return None
  raw = getattr(buffers[tensor.buffer], "data", None)
  if raw is None or len(raw) == 0:
    return None
  dtype = TENSOR_TYPE_TO_NUMPY.get(getattr(tensor, "type", 0))
  if dtype is None:  # packed INT4 / strings / resources: a constant, but not one this path reads
    return np.frombuffer(raw, dtype=np.uint8)
  data = np.frombuffer(raw, dtype=dtype)
  return data if tensor.shape is None else data.reshape(tuple(tensor.shape))
