"""QUANTIZE_TENSOR and ADD_DEQUANTIZE applied to the model object tree.

Mirror of ai_edge_quantizer/transformations/quantize_tensor.py (`quantize_tensor` :149-230,
`_perform_channelwise_quantization` :76-104, `_perform_blockwise_quantization` :107-147),
transformations/dequant_insert.py (`insert_dequant` :27-115) and
transformations/transformation_utils.py (`pack_data` :293-353, `add_new_constant_tensor`).
Bit packing of INT4 / INT2 payloads runs on the device (`aeqb_pack_bits`).
"""
from __future__ import annotations

import weakref

import ml_dtypes
import numpy as np

from .. import hostio
from .. import qtyping
from ..utils import tfl_model as tm


# The fused requantisation kernels write the packed INT4 / INT2 bytes in the same pass that
# produces the one-value-per-byte `quantized_data` the reference's dataclass carries; the packed
# copy is remembered here, keyed by the identity of the integer array, so that QUANTIZE_TENSOR's
# pack step (quantize_tensor.py:195-200) finds it instead of shipping the integers to the device
# again.  Entries die with their array.
_packed_by_array: dict = {}


def _owner(a: np.ndarray) -> np.ndarray:
  """The ndarray at the end of the `.base` chain (views collapse onto it)."""
  while isinstance(a.base, np.ndarray):
    a = a.base
  return a


def remember_packed(quantized_data: np.ndarray, bitwidth: int, packed: np.ndarray) -> None:
  own = _owner(quantized_data)
  if own.size != quantized_data.size:
    return  # a slice of something larger: identity would not identify it
  key = id(own)
  _packed_by_array[key] = (weakref.ref(own, lambda _r, k=key: _packed_by_array.pop(k, None)),
                           bitwidth, packed)


def lookup_packed(data: np.ndarray, bitwidth: int):
  """Packed bytes remembered for the array `data` is a whole (flat, reinterpreted) view of."""
  own = _owner(data)
  hit = _packed_by_array.get(id(own))
  if hit is not None and hit[0]() is own and hit[1] == bitwidth and own.size == data.size:
    return hit[2]
  return None


def pack_data(bitwidth: int, flattened_data: np.ndarray) -> np.ndarray:
  """INT4: two values per byte (even index low); INT2: four per byte; other widths pass."""
  flattened_data = flattened_data.reshape(-1)
  if bitwidth not in (2, 4):
    return flattened_data
  hit = lookup_packed(flattened_data, bitwidth)
  if hit is not None:
    return hit
  from .. import device
  q = hostio.to_device(np.ascontiguousarray(flattened_data).view(np.int8), np.int8)
  return hostio.to_host(device.pack_bits(q, bitwidth))


def quant_params_to_tflite_type(bitwidth: int) -> int:
  """TFLite integer type of a bit width, the reference's table (quantize_tensor.py:29-55):
  exactly 2 -> INT2, exactly 4 -> INT4 (the two packed widths), any other width up to 8 ->
  INT8 (3-, 5-, 6-, 7-bit values travel one per byte), then INT16 / INT32 / INT64."""
  if bitwidth == 2:
    return tm.TensorType.INT2
  if bitwidth == 4:
    return tm.TensorType.INT4
  if 1 < bitwidth <= 8:
    return tm.TensorType.INT8
  if 8 < bitwidth <= 16:
    return tm.TensorType.INT16
  if 16 < bitwidth <= 32:
    return tm.TensorType.INT32
  if 32 < bitwidth <= 64:
    return tm.TensorType.INT64
  raise ValueError(f"Unsupported bitwidth {bitwidth}.I")


def add_new_constant_tensor(name: bytes, data: np.ndarray, tensor_type: int, subgraph, model) -> int:
  """Appends a buffer + constant tensor; returns the tensor index."""
  model.buffers.append(tm.BufferT(data=np.frombuffer(np.ascontiguousarray(data).tobytes(), dtype=np.uint8)))
  subgraph.tensors.append(tm.TensorT(shape=np.asarray(data.shape, np.int32), type=tensor_type,
                                     buffer=len(model.buffers) - 1, name=name))
  return len(subgraph.tensors) - 1


def _channelwise(params: qtyping.UniformQuantParams) -> tm.QuantizationParametersT:
  q = tm.QuantizationParametersT()
  q.scale = np.ravel(params.scale).astype(np.float32, copy=False)
  if params.zero_point is not None:
    q.zeroPoint = np.ravel(params.zero_point).astype(np.int64, copy=False)
  if params.quantized_dimension is not None:
    q.quantizedDimension = int(params.quantized_dimension)
  return q


def _blockwise(params: qtyping.UniformQuantParams, tensor, subgraph, model) -> tm.QuantizationParametersT:
  q = tm.QuantizationParametersT(detailsType=tm.QuantizationDetails.BlockwiseQuantization)
  scales = params.scale.astype(ml_dtypes.bfloat16).astype(np.float16)  # exact: already 7-bit mantissas
  sid = add_new_constant_tensor(tensor.name + b"_scales", scales, tm.TensorType.FLOAT16, subgraph, model)
  q.details = tm.BlockwiseQuantizationT(scales=sid, zeroPoints=-1, blockSize=int(params.block_size))
  q.quantizedDimension = 0  # hard-coded in the reference (:145)
  return q


def quantize_tensor(model, subgraph, tensor_id: int, params, buffer_origin: dict) -> None:
  """Stores the quantised payload and parameters on tensor `tensor_id`."""
  tensor = subgraph.tensors[tensor_id]
  buffer_id = tensor.buffer
  if buffer_id and params.quantized_data is not None:
    origin = buffer_origin.get(buffer_id)
    if origin is not params:  # shared buffers are packed once (reference :176-185)
      buffer_origin[buffer_id] = params
      model.buffers[buffer_id].data = pack_data(
          params.num_bits, np.ravel(np.asarray(params.quantized_data)).view(np.uint8))
  if isinstance(params, qtyping.UniformQuantParams):
    tensor.quantization = (_channelwise(params) if params.block_size == 0
                           else _blockwise(params, tensor, subgraph, model))
    tensor.type = quant_params_to_tflite_type(params.num_bits)
  elif isinstance(params, qtyping.NonLinearQuantParams):
    if params.num_bits == 16:
      tensor.type = tm.TensorType.FLOAT16
    elif params.num_bits == 32:
      tensor.type = tm.TensorType.FLOAT32
    else:
      raise ValueError(f"Unsupported nonlinear params: {params.num_bits}")


def _opcode_index(model, builtin: int) -> int:
  for i, c in enumerate(model.operatorCodes):
    if tm.builtin_code(c) == builtin:
      return i
  model.operatorCodes.append(tm.OperatorCodeT(
      deprecatedBuiltinCode=min(builtin, tm.BuiltinOperator.PLACEHOLDER_FOR_GREATER_OP_CODES),
      builtinCode=builtin, version=1))
  return len(model.operatorCodes) - 1


def insert_dequant(model, subgraph, tensor_id: int, params, consumers: list, buffer_origin: dict) -> int:
  """Quantises the constant, then feeds `consumers` (operator objects of this subgraph) from a
  DEQUANTIZE op that restores a float tensor.  Returns the number of ops added (1)."""
  quantize_tensor(model, subgraph, tensor_id, params, buffer_origin)
  src = subgraph.tensors[tensor_id]
  subgraph.tensors.append(tm.TensorT(shape=None if src.shape is None else np.array(src.shape, np.int32),
                                     type=tm.TensorType.FLOAT32, buffer=0, name=src.name + b"_dequant",
                                     shapeSignature=src.shapeSignature))
  new_id = len(subgraph.tensors) - 1
  op = tm.OperatorT(opcodeIndex=_opcode_index(model, tm.BuiltinOperator.DEQUANTIZE),
                    inputs=np.array([tensor_id], np.int32), outputs=np.array([new_id], np.int32))
  first = min(i for i, o in enumerate(subgraph.operators) if any(o is c for c in consumers))
  for cons in consumers:
    ins = np.array(cons.inputs, np.int32)
    ins[ins == tensor_id] = new_id
    cons.inputs = ins
  subgraph.operators.insert(first, op)
  return 1
