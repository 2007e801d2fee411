"""The two weight-layout tables the hot path reads, plus constant-tensor access.

Mirror of ai_edge_quantizer/utils/tfl_flatbuffer_utils.py:95-106 (tables) and
:242-263 (`get_tensor_data`).  Graph objects are duck-typed: a tensor needs
`.buffer`, `.shape`, `.type`, `.name`; a buffer needs `.data` (bytes-like or a
uint8 array) — what the flatbuffer object API and our synthetic graphs both give.
"""
from __future__ import annotations

import numpy as np

from .. import qtyping

_Op = qtyping.TFLOperationName

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
                        4: np.int64, 7: np.int16, 9: np.int8}


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
  if dtype is None:
    raise ValueError(f"unsupported tensor type code {tensor.type}")
  return np.frombuffer(raw, dtype=dtype).reshape(tuple(tensor.shape))
