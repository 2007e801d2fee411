"""Data contracts that cross the plug-in boundary.

Same names, fields, defaults and equality rules as the reference's
`ai_edge_quantizer/qtyping.py` (enums :80-200, UniformQuantParams :205-313,
TensorQuantizationConfig :384-445, OpQuantizationConfig :465-551, OpInfo :567,
GraphInfo :554, GetTensorQuantParamsFuncSignature :702-710), re-declared without
the `ai_edge_litert` schema aliases (:37-79): on this path graph objects are
opaque handles that are only forwarded.
"""
from __future__ import annotations

import collections
import copy
import dataclasses
import enum
from typing import Any, Callable, Mapping, MutableMapping, Optional, Union

import numpy as np

QSV = MutableMapping[str, Any]
ModelQuantizationRecipe = list  # list[dict[str, Any]]

# Opaque graph handles (flatbuffer object-API instances in the reference).
TensorT = Any
OperatorT = Any
BufferT = Any
SubGraphT = Any
ModelT = Any

_OP_NAMES = (
    "INPUT OUTPUT FULLY_CONNECTED BATCH_MATMUL DEPTHWISE_CONV_2D CONV_2D"
    " CONV_2D_TRANSPOSE AVERAGE_POOL_2D RESHAPE CUSTOM_OP EMBEDDING_LOOKUP"
    " SOFTMAX TANH TRANSPOSE GELU ADD SUB MUL MEAN RSQRT CONCATENATION"
    " STRIDED_SLICE SPLIT LOGISTIC SLICE SUM SELECT SELECT_V2"
    " DYNAMIC_UPDATE_SLICE STABLEHLO_COMPOSITE PAD SQUARED_DIFFERENCE"
    " MAX_POOL_2D RESIZE_BILINEAR RESIZE_NEAREST_NEIGHBOR GATHER_ND PACK UNPACK"
    " DIV BROADCAST_TO SQRT GATHER HARD_SWISH MAXIMUM PADV2 REDUCE_MIN EQUAL"
    " NOT_EQUAL MIRROR_PAD SPACE_TO_DEPTH RELU"
).split()


class _StrEnum(str, enum.Enum):
  pass


# qtyping.py:80-137 — TFLite op names; value == name except the wildcard.
TFLOperationName = _StrEnum(
    "TFLOperationName",
    [("ALL_SUPPORTED", "*")] + [(n, n) for n in _OP_NAMES],
    module=__name__,
)


class QuantizeMode(enum.Enum):
  CALIBRATE = 2
  MATERIALIZE = 3


class OpExecutionMode(str, enum.Enum):
  WEIGHT_ONLY = "WEIGHT_ONLY"
  DRQ = "DRQ"
  SRQ = "SRQ"


class ComputePrecision(str, enum.Enum):
  INTEGER = "INTEGER"
  FLOAT = "FLOAT"


class TensorDataType(str, enum.Enum):
  INT = "INT"
  FLOAT = "FLOAT"


class QuantGranularity(str, enum.Enum):
  TENSORWISE = "TENSORWISE"
  CHANNELWISE = "CHANNELWISE"
  BLOCKWISE_32 = "BLOCKWISE_32"
  BLOCKWISE_64 = "BLOCKWISE_64"
  BLOCKWISE_128 = "BLOCKWISE_128"
  BLOCKWISE_256 = "BLOCKWISE_256"


class QuantTransformation(enum.Enum):
  NO_QUANTIZE = 0
  ADD_QUANTIZE = 1
  ADD_DEQUANTIZE = 2
  QUANTIZE_TENSOR = 3
  EMULATED_SUBCHANNEL = 4
  DUPLICATE_BUFFER = 5
  DUPLICATE_TENSOR = 6
  INSERT_HADAMARD_ROTATION = 7
  INSERT_DECOMPOSED_HADAMARD_ROTATION = 8
  INSERT_MULTIPLY = 9


def _arrays_match(a, b) -> bool:
  if a is None or b is None:
    return a is None and b is None
  return a is b or np.array_equal(a, b)


def _values_match(a, b) -> bool:
  a_arr, b_arr = isinstance(a, np.ndarray), isinstance(b, np.ndarray)
  if a_arr or b_arr:
    return a_arr and b_arr and np.array_equal(a, b)
  return a == b


def _param_dicts_match(a, b) -> bool:
  if a is None or b is None:
    return a is None and b is None
  return a.keys() == b.keys() and all(_values_match(v, b[k]) for k, v in a.items())


class HadamardRotationParams:
  """`random_binary_vector` (always ones today) and the Hadamard block size."""

  def __init__(self, random_binary_vector: np.ndarray, hadamard_size: int):
    self.random_binary_vector = random_binary_vector
    self.hadamard_size = hadamard_size

  def __eq__(self, other):
    if other.__class__ is not self.__class__:
      return NotImplemented
    return self is other or (
        self.hadamard_size == other.hadamard_size
        and np.array_equal(self.random_binary_vector, other.random_binary_vector))

  __hash__ = None

  def __repr__(self):
    return f"HadamardRotationParams(hadamard_size={self.hadamard_size})"


@dataclasses.dataclass(frozen=True, eq=False)
class UniformQuantParams:
  """Result of quantising one tensor; arrays are owned and immutable by contract."""
  HadamardRotationParams = HadamardRotationParams  # reference nests the class

  num_bits: int
  quantized_dimension: Optional[int]
  scale: np.ndarray
  zero_point: np.ndarray
  symmetric: bool = True
  quantized_data: Optional[np.ndarray] = None
  block_size: int = 0
  hadamard: Optional[HadamardRotationParams] = None
  custom_algorithm_param: Optional[dict] = None

  @classmethod
  def from_tfl_tensor_details(cls, tensor_detail) -> "UniformQuantParams":
    qp = tensor_detail["quantization_parameters"]
    bits_of = {np.int8: 8, np.int16: 16, np.int32: 32, np.int64: 64}
    dtype = tensor_detail["dtype"]
    if dtype not in bits_of:
      raise ValueError(
          f"Unsupported data type: {dtype}. Supported types are np.int8,"
          " np.int16, np.int32, np.int64.")
    return cls(
        num_bits=bits_of[dtype],
        quantized_dimension=qp["quantized_dimension"],
        scale=qp["scales"],
        zero_point=qp["zero_points"],
        symmetric=sum(abs(qp["zero_points"])) == 0,
        block_size=qp["block_size"],
    )

  def __eq__(self, other):
    if other.__class__ is not self.__class__:
      return NotImplemented
    if self is other:
      return True
    return (
        (self.num_bits, self.quantized_dimension, self.symmetric, self.block_size)
        == (other.num_bits, other.quantized_dimension, other.symmetric, other.block_size)
        and _arrays_match(self.scale, other.scale)
        and _arrays_match(self.zero_point, other.zero_point)
        and _arrays_match(self.quantized_data, other.quantized_data)
        and self.hadamard == other.hadamard
        and _param_dicts_match(self.custom_algorithm_param, other.custom_algorithm_param))

  __hash__ = None


@dataclasses.dataclass(frozen=True, eq=False)
class NonLinearQuantParams:
  num_bits: int
  quantized_data: Optional[np.ndarray]
  data_type: TensorDataType = TensorDataType.FLOAT

  def __eq__(self, other):
    if other.__class__ is not self.__class__:
      return NotImplemented
    return self is other or (
        self.num_bits == other.num_bits and self.data_type == other.data_type
        and _arrays_match(self.quantized_data, other.quantized_data))

  __hash__ = None


@dataclasses.dataclass(frozen=True)
class OpToTensorParams:
  subgraph_op_id: int
  transformations: list
  parameters: Union[None, UniformQuantParams, NonLinearQuantParams] = None


@dataclasses.dataclass
class TensorTransformationParams:
  tensor_name: str
  producer: Optional[OpToTensorParams] = None
  consumers: Optional[list] = None

  def __copy__(self):
    return TensorTransformationParams(
        self.tensor_name, self.producer,
        None if self.consumers is None else list(self.consumers))


class FrozenParams(dict):
  """Hashable read-only mapping for `algorithm_params` (immutabledict stand-in)."""

  def _ro(self, *a, **k):
    raise TypeError("algorithm_params is read-only")

  __setitem__ = __delitem__ = clear = pop = popitem = setdefault = update = _ro

  def __hash__(self):
    return hash(frozenset(self.items()))


_BLOCK_GRANULARITY = {
    32: QuantGranularity.BLOCKWISE_32,
    64: QuantGranularity.BLOCKWISE_64,
    128: QuantGranularity.BLOCKWISE_128,
    256: QuantGranularity.BLOCKWISE_256,
}


def _plain(d):
  """dataclasses.asdict factory: drops None / empty mappings, thaws FrozenParams."""
  out = {}
  for k, v in d:
    if v is None or (isinstance(v, Mapping) and not v):
      continue
    out[k] = dict(v) if isinstance(v, Mapping) and type(v) is not dict else v
  return out


@dataclasses.dataclass(frozen=True)
class TensorQuantizationConfig:
  """Per-tensor request; frozen + hashable because it keys TensorQuantParamsCache."""
  num_bits: int
  symmetric: bool = True
  granularity: QuantGranularity = QuantGranularity.TENSORWISE
  dtype: TensorDataType = TensorDataType.INT
  algorithm_params: Mapping[str, Any] = dataclasses.field(default_factory=FrozenParams)

  def __post_init__(self):
    if not isinstance(self.algorithm_params, FrozenParams):
      object.__setattr__(self, "algorithm_params", FrozenParams(self.algorithm_params))

  def to_dict(self) -> dict:
    return dataclasses.asdict(self, dict_factory=_plain)

  @classmethod
  def from_dict(cls, params: dict) -> "TensorQuantizationConfig":
    p = copy.deepcopy(params)
    block = p.pop("block_size", 0)  # legacy recipes (qtyping.py:448-462)
    if block > 0:
      if block not in _BLOCK_GRANULARITY:
        raise ValueError(f"Unsupported block size: {block}")
      p["granularity"] = _BLOCK_GRANULARITY[block]
    known = {f.name for f in dataclasses.fields(cls)}
    extra = p.pop("algorithm_params", {})
    for key in [k for k in p if k not in known]:
      extra[key] = p.pop(key)  # unknown keys fold into algorithm_params (:438-445)
    return cls(algorithm_params=extra, **p)


@dataclasses.dataclass(frozen=True)
class OpQuantizationConfig:
  activation_tensor_config: Optional[TensorQuantizationConfig] = None
  weight_tensor_config: Optional[TensorQuantizationConfig] = None
  compute_precision: ComputePrecision = ComputePrecision.FLOAT
  explicit_dequantize: bool = False
  skip_checks: bool = False
  min_weight_elements: int = 0

  def __post_init__(self):
    act, wgt = self.activation_tensor_config, self.weight_tensor_config
    if act is None or wgt is None:
      return
    if act.dtype == TensorDataType.INT and wgt.dtype == TensorDataType.FLOAT:
      raise ValueError(
          "An op can not be set to have integer activation but float weights!")
    if (act.dtype == TensorDataType.INT and wgt.dtype == TensorDataType.INT
        and self.compute_precision != ComputePrecision.INTEGER):
      raise ValueError(
          "Op execution mode must be SRQ (static range quantization) if both"
          " activation and weight tensors are quantized!")

  def to_dict(self) -> dict:
    return dataclasses.asdict(self, dict_factory=_plain)

  @classmethod
  def from_dict(cls, params: dict) -> "OpQuantizationConfig":
    p = copy.deepcopy(params)
    p["weight_tensor_config"] = TensorQuantizationConfig.from_dict(p["weight_tensor_config"])
    if "activation_tensor_config" in p:
      p["activation_tensor_config"] = TensorQuantizationConfig.from_dict(
          p["activation_tensor_config"])
    return cls(**p)


@dataclasses.dataclass(frozen=True)
class GraphInfo:
  subgraph_tensors: list
  buffers: list


@dataclasses.dataclass(frozen=True)
class OpInfo:
  op: OperatorT
  op_name: TFLOperationName
  subgraph_op_index: int
  op_quant_config: OpQuantizationConfig


@dataclasses.dataclass
class TransformationInst:
  transformation: QuantTransformation
  tensor_id: int
  producer: Optional[int]
  consumers: list
  parameters: Union[None, UniformQuantParams, NonLinearQuantParams] = None


@dataclasses.dataclass
class TensorTransformationInsts:
  tensor_name: str
  subgraph_id: int
  instructions: Optional[list]


@dataclasses.dataclass(frozen=True)
class TransformationInfo:
  op_id: int
  num_ops_added: int
  output_tensor_id: int


@dataclasses.dataclass(frozen=True)
class IOOperator:
  inputs: list
  outputs: list
  op_key: TFLOperationName


ConfigCheckPolicyDict = collections.OrderedDict

GetTensorQuantParamsFuncSignature = Callable[
    [OpInfo, TensorQuantizationConfig, Optional[np.ndarray], Optional[dict]],
    UniformQuantParams,
]
