"""TFLite (`TFL3`) model object tree, read from / written to FlatBuffer bytes.

Stands in for `ai_edge_litert.tools.flatbuffer_utils.read_model / write_model` and the generated
schema classes the reference aliases in qtyping.py:37-79 (`ModelT`, `SubGraphT`, `TensorT`,
`OperatorT`, `BufferT`, `QuantizationParametersT`, `BlockwiseQuantizationT`, ...), with the
object-API attribute names the reference code uses (`operatorCodes`, `opcodeIndex`,
`builtinCode`, `zeroPoint`, `quantizedDimension`, `detailsType`, ...).  Field ids follow the
public TFLite schema (schema.fbs, file identifier TFL3).

Scope: everything the quantizer reads or rewrites is parsed into attributes.  Operator option
tables (`builtin_options`) are carried through as opaque scalar-only tables, which is what all
but a handful of them are; the ones holding vectors are handled explicitly (Reshape, Squeeze,
ConcatEmbeddings, and StableHLOCompositeOptions of `builtin_options_2`) or rejected at write time
(VarHandle, Bucketize, the other StableHLO options), as are sparsity and variant tensors.  Constant data is exposed as
zero-copy views of the input bytes (mmap-friendly: nothing is copied at read time).
"""
from __future__ import annotations

import dataclasses
import mmap
import os
import struct
from typing import Any, Optional

import numpy as np

from . import flatbuffer_lite as fb

FILE_IDENTIFIER = b"TFL3"


class TensorType:
  """schema.fbs TensorType."""
  FLOAT32, FLOAT16, INT32, UINT8, INT64, STRING, BOOL, INT16, COMPLEX64, INT8 = range(10)
  FLOAT64, COMPLEX128, UINT64, RESOURCE, VARIANT, UINT32, UINT16, INT4, BFLOAT16, INT2 = range(10, 20)


class BuiltinOperator:
  """The builtin codes this package names (schema.fbs BuiltinOperator)."""
  ADD = 0
  CONV_2D = 3
  DEPTHWISE_CONV_2D = 4
  DEQUANTIZE = 6
  EMBEDDING_LOOKUP = 7
  FULLY_CONNECTED = 9
  MUL = 18
  TRANSPOSE_CONV = 67
  QUANTIZE = 114
  BATCH_MATMUL = 126
  PLACEHOLDER_FOR_GREATER_OP_CODES = 127


class QuantizationDetails:
  NONE, CustomQuantization, BlockwiseQuantization = 0, 1, 2


# builtin_options union members that are NOT scalar-only tables
_OPT_CONCAT_EMBEDDINGS, _OPT_RESHAPE, _OPT_SQUEEZE, _OPT_VAR_HANDLE, _OPT_BUCKETIZE = 3, 17, 30, 111, 115


@dataclasses.dataclass(eq=False)
class RawTable:
  """A scalar-only table carried through verbatim (vtable, inline bytes, position mod 8)."""
  vtable: bytes
  table: bytes
  pos_mod8: int = 0

  def scalar(self, field_id: int, kind: str, default=0):
    """Reads a field of the carried table (e.g. FullyConnectedOptions.keep_num_dims)."""
    import struct
    data = bytearray(self.vtable + self.table)
    struct.pack_into("<i", data, len(self.vtable), len(self.vtable))  # the vtable now sits at 0
    return fb.Table(bytes(data), len(self.vtable)).scalar(field_id, kind, default)


@dataclasses.dataclass(eq=False)
class IntVectorOptions:
  """Reshape / Squeeze options: a single [int] field 0."""
  values: Optional[np.ndarray] = None


@dataclasses.dataclass(eq=False)
class ConcatEmbeddingsOptions:
  numChannels: int = 0
  numColumnsPerChannel: Optional[np.ndarray] = None
  embeddingDimPerChannel: Optional[np.ndarray] = None


@dataclasses.dataclass(eq=False)
class StableHLOCompositeOptionsT:
  """builtin_options_2 member 21 (STABLEHLO_COMPOSITE ops: `odml.*` fused subgraphs)."""
  name: Optional[bytes] = None
  decompositionSubgraphIndex: int = 0
  compositeAttributes: Optional[np.ndarray] = None
  compositeAttributesFormat: int = 0
  version: int = 0


_OPT2_STABLEHLO_COMPOSITE = 21


@dataclasses.dataclass(eq=False)
class BlockwiseQuantizationT:
  scales: int = 0
  zeroPoints: int = 0
  blockSize: int = 0


@dataclasses.dataclass(eq=False)
class CustomQuantizationT:
  custom: Optional[np.ndarray] = None


@dataclasses.dataclass(eq=False)
class QuantizationParametersT:
  min: Optional[np.ndarray] = None
  max: Optional[np.ndarray] = None
  scale: Optional[np.ndarray] = None
  zeroPoint: Optional[np.ndarray] = None
  detailsType: int = 0
  details: Any = None
  quantizedDimension: int = 0


@dataclasses.dataclass(eq=False)
class TensorT:
  shape: Optional[np.ndarray] = None
  type: int = 0
  buffer: int = 0
  name: Optional[bytes] = None
  quantization: Optional[QuantizationParametersT] = None
  isVariable: bool = False
  shapeSignature: Optional[np.ndarray] = None
  hasRank: bool = False
  sparsity: Any = None          # carried only as "present" (rejected at write time)
  variantTensors: Any = None


@dataclasses.dataclass(eq=False)
class OperatorT:
  opcodeIndex: int = 0
  inputs: Optional[np.ndarray] = None
  outputs: Optional[np.ndarray] = None
  builtinOptionsType: int = 0
  builtinOptions: Any = None
  customOptions: Optional[np.ndarray] = None
  customOptionsFormat: int = 0
  mutatingVariableInputs: Optional[np.ndarray] = None
  intermediates: Optional[np.ndarray] = None
  largeCustomOptionsOffset: int = 0
  largeCustomOptionsSize: int = 0
  builtinOptions2Type: int = 0
  builtinOptions2: Any = None
  debugMetadataIndex: int = -1


@dataclasses.dataclass(eq=False)
class OperatorCodeT:
  deprecatedBuiltinCode: int = 0
  customCode: Optional[bytes] = None
  version: int = 1
  builtinCode: int = 0


@dataclasses.dataclass(eq=False)
class SubGraphT:
  tensors: list = dataclasses.field(default_factory=list)
  inputs: Optional[np.ndarray] = None
  outputs: Optional[np.ndarray] = None
  operators: list = dataclasses.field(default_factory=list)
  name: Optional[bytes] = None
  debugMetadataIndex: int = -1


@dataclasses.dataclass(eq=False)
class BufferT:
  data: Any = None   # bytes-like / uint8 array, or None
  offset: int = 0
  size: int = 0


@dataclasses.dataclass(eq=False)
class MetadataT:
  name: Optional[bytes] = None
  buffer: int = 0


@dataclasses.dataclass(eq=False)
class TensorMapT:
  name: Optional[bytes] = None
  tensorIndex: int = 0


@dataclasses.dataclass(eq=False)
class SignatureDefT:
  inputs: list = dataclasses.field(default_factory=list)
  outputs: list = dataclasses.field(default_factory=list)
  signatureKey: Optional[bytes] = None
  subgraphIndex: int = 0


@dataclasses.dataclass(eq=False)
class ModelT:
  version: int = 0
  operatorCodes: list = dataclasses.field(default_factory=list)
  subgraphs: list = dataclasses.field(default_factory=list)
  description: Optional[bytes] = None
  buffers: list = dataclasses.field(default_factory=list)
  metadataBuffer: Optional[np.ndarray] = None
  metadata: list = dataclasses.field(default_factory=list)
  signatureDefs: list = dataclasses.field(default_factory=list)


# ------------------------------------------------------------------------------ reading
def _read_quantization(t: fb.Table) -> QuantizationParametersT:
  q = QuantizationParametersT(
      min=t.scalar_vector(0, "float"), max=t.scalar_vector(1, "float"),
      scale=t.scalar_vector(2, "float"), zeroPoint=t.scalar_vector(3, "long"),
      detailsType=t.scalar(4, "ubyte"), quantizedDimension=t.scalar(6, "int"))
  d = t.table(5)
  if d is not None:
    if q.detailsType == QuantizationDetails.BlockwiseQuantization:
      q.details = BlockwiseQuantizationT(d.scalar(0, "int"), d.scalar(1, "int"), d.scalar(2, "int"))
    elif q.detailsType == QuantizationDetails.CustomQuantization:
      q.details = CustomQuantizationT(d.scalar_vector(0, "ubyte"))
  return q


def _read_tensor(t: fb.Table) -> TensorT:
  qt = t.table(4)
  return TensorT(
      shape=t.scalar_vector(0, "int"), type=t.scalar(1, "byte"), buffer=t.scalar(2, "uint"),
      name=t.string(3), quantization=None if qt is None else _read_quantization(qt),
      isVariable=bool(t.scalar(5, "bool", False)), shapeSignature=t.scalar_vector(7, "int"),
      hasRank=bool(t.scalar(8, "bool", False)),
      sparsity=True if t.has(6) else None, variantTensors=True if t.has(9) else None)


def _read_options(kind: int, t: Optional[fb.Table]):
  if t is None:
    return None
  if kind in (_OPT_RESHAPE, _OPT_SQUEEZE):
    return IntVectorOptions(t.scalar_vector(0, "int"))
  if kind == _OPT_CONCAT_EMBEDDINGS:
    return ConcatEmbeddingsOptions(t.scalar(0, "int"), t.scalar_vector(1, "int"), t.scalar_vector(2, "int"))
  return RawTable(*t.raw())


def _read_operator(t: fb.Table) -> OperatorT:
  kind = t.scalar(3, "ubyte")
  op = OperatorT(
      opcodeIndex=t.scalar(0, "uint"), inputs=t.scalar_vector(1, "int"), outputs=t.scalar_vector(2, "int"),
      builtinOptionsType=kind, builtinOptions=_read_options(kind, t.table(4)),
      customOptions=t.scalar_vector(5, "ubyte"), customOptionsFormat=t.scalar(6, "byte"),
      mutatingVariableInputs=t.scalar_vector(7, "bool"), intermediates=t.scalar_vector(8, "int"),
      largeCustomOptionsOffset=t.scalar(9, "ulong"), largeCustomOptionsSize=t.scalar(10, "ulong"),
      builtinOptions2Type=t.scalar(11, "ubyte"), debugMetadataIndex=t.scalar(13, "int", -1))
  if t.has(12):
    if op.builtinOptions2Type == _OPT2_STABLEHLO_COMPOSITE:
      o = t.table(12)
      op.builtinOptions2 = StableHLOCompositeOptionsT(
          o.string(0), o.scalar(1, "int"), o.scalar_vector(2, "ubyte"), o.scalar(3, "byte"), o.scalar(4, "int"))
    else:
      op.builtinOptions2 = True  # other StableHLO options: present, not modelled (rejected at write time)
  return op


def _read_subgraph(t: fb.Table) -> SubGraphT:
  return SubGraphT(
      tensors=[_read_tensor(x) for x in t.table_vector(0)], inputs=t.scalar_vector(1, "int"),
      outputs=t.scalar_vector(2, "int"), operators=[_read_operator(x) for x in t.table_vector(3)],
      name=t.string(4), debugMetadataIndex=t.scalar(5, "int", -1))


def read_model_from_bytes(buf) -> ModelT:
  """Parses TFL3 bytes (bytes, bytearray, mmap, memoryview, uint8 array).  Buffers whose data
  lives outside the flatbuffer (`offset` / `size`, models > 2 GB) are resolved to views too."""
  view = memoryview(buf)
  if len(view) < 8 or fb.file_identifier(view) != FILE_IDENTIFIER:
    raise ValueError("not a TFLite flatbuffer: the TFL3 file identifier is missing")
  root = fb.Table.root(view)
  m = ModelT(version=root.scalar(0, "uint"), description=root.string(3),
             metadataBuffer=root.scalar_vector(5, "int"))
  for t in root.table_vector(1):
    m.operatorCodes.append(OperatorCodeT(t.scalar(0, "byte"), t.string(1), t.scalar(2, "int", 1),
                                         t.scalar(3, "int")))
  m.subgraphs = [_read_subgraph(t) for t in root.table_vector(2)]
  for t in root.table_vector(4):
    b = BufferT(data=t.scalar_vector(0, "ubyte"), offset=t.scalar(1, "ulong"), size=t.scalar(2, "ulong"))
    if b.data is None and b.offset > 1 and b.size > 0:
      b.data = np.frombuffer(view, dtype=np.uint8, count=b.size, offset=b.offset)
      b.offset = b.size = 0   # rewritten inline (or re-externalised) by write_model
    m.buffers.append(b)
  m.metadata = [MetadataT(t.string(0), t.scalar(1, "uint")) for t in root.table_vector(6)]
  for t in root.table_vector(7):
    maps = lambda fid: [TensorMapT(x.string(0), x.scalar(1, "uint")) for x in t.table_vector(fid)]
    m.signatureDefs.append(SignatureDefT(maps(0), maps(1), t.string(2), t.scalar(4, "uint")))
  return m


def read_model(path: str) -> ModelT:
  """Memory-maps the file read-only: constant tensors stay views of the page cache."""
  with open(path, "rb") as f:
    if os.fstat(f.fileno()).st_size == 0:
      raise ValueError(f"{path} is empty")
    mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
  return read_model_from_bytes(mm)


def builtin_code(op_code: OperatorCodeT) -> int:
  """max(deprecated_builtin_code, builtin_code): codes < 127 may live in either field."""
  return max(op_code.deprecatedBuiltinCode, op_code.builtinCode)


# ------------------------------------------------------------------------------ writing
def _vec(b: fb.Builder, arr, dtype) -> int:
  return 0 if arr is None else b.create_numpy_vector(np.asarray(arr, dtype=dtype))


def _str(b: fb.Builder, s) -> int:
  return 0 if s is None else b.create_string(s)


def _write_quantization(b: fb.Builder, q: QuantizationParametersT) -> int:
  details = 0
  if q.details is not None:
    if q.detailsType == QuantizationDetails.BlockwiseQuantization:
      b.start_table()
      b.add_scalar(0, "int", int(q.details.scales))
      b.add_scalar(1, "int", int(q.details.zeroPoints))
      b.add_scalar(2, "int", int(q.details.blockSize))
      details = b.end_table()
    elif q.detailsType == QuantizationDetails.CustomQuantization:
      v = _vec(b, q.details.custom, np.uint8)
      b.start_table()
      b.add_offset(0, v)
      details = b.end_table()
  mn, mx = _vec(b, q.min, np.float32), _vec(b, q.max, np.float32)
  sc, zp = _vec(b, q.scale, np.float32), _vec(b, q.zeroPoint, np.int64)
  b.start_table()
  b.add_offset(0, mn)
  b.add_offset(1, mx)
  b.add_offset(2, sc)
  b.add_offset(3, zp)
  if details:
    b.add_scalar(4, "ubyte", int(q.detailsType))
    b.add_offset(5, details)
  b.add_scalar(6, "int", int(q.quantizedDimension))
  return b.end_table()


def _write_tensor(b: fb.Builder, t: TensorT) -> int:
  if t.sparsity is not None or t.variantTensors is not None:
    raise NotImplementedError("sparse / variant tensors are not serialised by aeq_b200")
  shape, sig = _vec(b, t.shape, np.int32), _vec(b, t.shapeSignature, np.int32)
  name = _str(b, t.name)
  quant = 0 if t.quantization is None else _write_quantization(b, t.quantization)
  b.start_table()
  b.add_offset(0, shape)
  b.add_scalar(1, "byte", int(t.type))
  b.add_scalar(2, "uint", int(t.buffer))
  b.add_offset(3, name)
  b.add_offset(4, quant)
  b.add_scalar(5, "bool", bool(t.isVariable), False)
  b.add_offset(7, sig)
  b.add_scalar(8, "bool", bool(t.hasRank), False)
  return b.end_table()


def _write_options(b: fb.Builder, kind: int, opt) -> int:
  if opt is None:
    return 0
  if isinstance(opt, RawTable):
    if kind in (_OPT_VAR_HANDLE, _OPT_BUCKETIZE):
      raise NotImplementedError(f"builtin options type {kind} holds vectors and is not serialised")
    return b.add_raw_table(opt.vtable, opt.table, opt.pos_mod8)
  if isinstance(opt, IntVectorOptions):
    v = _vec(b, opt.values, np.int32)
    b.start_table()
    b.add_offset(0, v)
    return b.end_table()
  if isinstance(opt, ConcatEmbeddingsOptions):
    v1, v2 = _vec(b, opt.numColumnsPerChannel, np.int32), _vec(b, opt.embeddingDimPerChannel, np.int32)
    b.start_table()
    b.add_scalar(0, "int", int(opt.numChannels))
    b.add_offset(1, v1)
    b.add_offset(2, v2)
    return b.end_table()
  raise NotImplementedError(f"cannot serialise builtin options {type(opt).__name__}")


def _write_operator(b: fb.Builder, op: OperatorT) -> int:
  opts2 = 0
  if isinstance(op.builtinOptions2, StableHLOCompositeOptionsT):
    o = op.builtinOptions2
    name, attrs = _str(b, o.name), _vec(b, o.compositeAttributes, np.uint8)
    b.start_table()
    b.add_offset(0, name)
    b.add_scalar(1, "int", int(o.decompositionSubgraphIndex))
    b.add_offset(2, attrs)
    b.add_scalar(3, "byte", int(o.compositeAttributesFormat))
    b.add_scalar(4, "int", int(o.version))
    opts2 = b.end_table()
  elif op.builtinOptions2 is not None:
    raise NotImplementedError(
        f"builtin_options_2 type {op.builtinOptions2Type} (StableHLO) is not serialised by aeq_b200")
  ins, outs = _vec(b, op.inputs, np.int32), _vec(b, op.outputs, np.int32)
  opts = _write_options(b, op.builtinOptionsType, op.builtinOptions)
  custom = _vec(b, op.customOptions, np.uint8)
  mut = _vec(b, op.mutatingVariableInputs, np.bool_)
  inter = _vec(b, op.intermediates, np.int32)
  b.start_table()
  b.add_scalar(0, "uint", int(op.opcodeIndex))
  b.add_offset(1, ins)
  b.add_offset(2, outs)
  if opts:
    b.add_scalar(3, "ubyte", int(op.builtinOptionsType))
    b.add_offset(4, opts)
  b.add_offset(5, custom)
  b.add_scalar(6, "byte", int(op.customOptionsFormat))
  b.add_offset(7, mut)
  b.add_offset(8, inter)
  b.add_scalar(9, "ulong", int(op.largeCustomOptionsOffset))
  b.add_scalar(10, "ulong", int(op.largeCustomOptionsSize))
  if opts2:
    b.add_scalar(11, "ubyte", int(op.builtinOptions2Type))
    b.add_offset(12, opts2)
  b.add_scalar(13, "int", int(op.debugMetadataIndex), -1)
  return b.end_table()


def _write_subgraph(b: fb.Builder, g: SubGraphT) -> int:
  tensors = b.create_offset_vector([_write_tensor(b, t) for t in g.tensors])
  ops = b.create_offset_vector([_write_operator(b, o) for o in g.operators])
  ins, outs = _vec(b, g.inputs, np.int32), _vec(b, g.outputs, np.int32)
  name = _str(b, g.name)
  b.start_table()
  b.add_offset(0, tensors)
  b.add_offset(1, ins)
  b.add_offset(2, outs)
  b.add_offset(3, ops)
  b.add_offset(4, name)
  b.add_scalar(5, "int", int(g.debugMetadataIndex), -1)
  return b.end_table()


# The reference's ModelModifier moves every buffer of 1 KiB or more out of the flatbuffer as soon
# as those buffers add up to 256 KiB (model_modifier.py:43-75, :202-215): payloads follow the
# flatbuffer, each 16-byte aligned, and the buffer tables carry `offset` / `size`.  Same rule
# here; it is also what keeps models above the 2 GB flatbuffer limit writable, and it means a
# payload is copied once (into the output) instead of twice (builder, then output).
EXTERNAL_MIN_BUFFER_BYTES = 1024
EXTERNAL_MIN_TOTAL_BYTES = 256 * 1024


def _payload(x: BufferT):
  if x.data is None or len(x.data) == 0:
    return None
  if isinstance(x.data, (bytes, bytearray, memoryview)):
    return x.data
  return np.ascontiguousarray(np.asarray(x.data).view(np.uint8) if isinstance(x.data, np.ndarray) else x.data)


def _uninitialised_bytearray(n: int) -> bytearray:
  """bytearray of n bytes WITHOUT the zero fill: `bytearray(n)` memsets on one thread, 0.55 s per GiB
  of fresh pages -- more than everything else `Quantizer.quantize()` does to a 4 GiB model -- while
  every byte of the result is written below anyway (payloads by the copy threads, the <= 15-byte
  alignment gaps explicitly).  CPython's own constructor with a NULL source allocates and does not
  touch the memory; anything unexpected falls back to the zero-filled one."""
  try:
    import ctypes
    f = ctypes.pythonapi.PyByteArray_FromStringAndSize
    f.restype = ctypes.py_object
    f.argtypes = [ctypes.c_char_p, ctypes.c_ssize_t]
    out = f(None, n)
    if isinstance(out, bytearray) and len(out) == n:
      return out
  except Exception:  # pylint: disable=broad-except
    pass
  return bytearray(n)


def _copy_pieces(out: bytearray, pieces) -> None:
  """out[pos : pos + len(src)] = src for every (pos, src); large models on a few threads (the
  destination is fresh memory: page faults and the copy both spread over the cores, NumPy drops
  the GIL while it copies)."""
  dst = np.frombuffer(out, dtype=np.uint8)
  jobs = []
  step = 8 << 20
  for pos, src in pieces:
    a = np.frombuffer(src, dtype=np.uint8) if not isinstance(src, np.ndarray) else src.reshape(-1).view(np.uint8)
    for o in range(0, a.size, step):
      jobs.append((pos + o, a[o:o + step]))
  total = sum(a.size for _, a in jobs)
  if total < (64 << 20):
    for pos, a in jobs:
      dst[pos:pos + a.size] = a
    return
  from concurrent.futures import ThreadPoolExecutor

  def put(job):
    dst[job[0]:job[0] + job[1].size] = job[1]
  with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
    list(ex.map(put, jobs))


def write_model_to_bytes(m: ModelT, external_buffers: Optional[bool] = None):
  """Serialises the object tree; returns bytes (or a bytearray when buffers are external).

  external_buffers: None = the reference's rule (see above), True / False = force.  Inline
  payloads are 16-byte aligned like the schema's `force_align: 16`; external ones follow the
  flatbuffer, 16-byte aligned, and `read_model_from_bytes` resolves them back into views."""
  payloads = [_payload(x) for x in m.buffers]
  big = sum(len(p) for p in payloads if p is not None and len(p) >= EXTERNAL_MIN_BUFFER_BYTES)
  forced = external_buffers is True
  if external_buffers is None:
    external_buffers = big >= EXTERNAL_MIN_TOTAL_BYTES
  # automatic: buffers of 1 KiB and more leave the flatbuffer; forced: every non-empty buffer does
  is_ext = [bool(external_buffers) and p is not None and (forced or len(p) >= EXTERNAL_MIN_BUFFER_BYTES)
            for p in payloads]
  inline_total = sum(len(p) for p, e in zip(payloads, is_ext) if p is not None and not e)
  if inline_total >= (1 << 31) - (1 << 24):
    raise ValueError("more than 2 GB of buffers would stay inside the flatbuffer")
  b = fb.Builder(max(1 << 16, int(inline_total * 1.05) + (1 << 16)))
  buffers = []
  for x, payload, ext in zip(m.buffers, payloads, is_ext):
    data = 0
    if payload is not None and not ext:
      data = b.create_byte_vector(payload, align=16)
    b.start_table()
    b.add_offset(0, data)
    if ext:
      b.add_scalar(1, "ulong", 1)  # placeholders: non-default, so the fields exist and can be
      b.add_scalar(2, "ulong", 1)  # patched once the payload positions are known
    else:
      b.add_scalar(1, "ulong", int(x.offset))
      b.add_scalar(2, "ulong", int(x.size))
    buffers.append(b.end_table())
  buffers_v = b.create_offset_vector(buffers)
  codes = []
  for c in m.operatorCodes:
    custom = _str(b, c.customCode)
    b.start_table()
    b.add_scalar(0, "byte", int(c.deprecatedBuiltinCode))
    b.add_offset(1, custom)
    b.add_scalar(2, "int", int(c.version), 1)
    b.add_scalar(3, "int", int(c.builtinCode))
    codes.append(b.end_table())
  codes_v = b.create_offset_vector(codes)
  subgraphs_v = b.create_offset_vector([_write_subgraph(b, g) for g in m.subgraphs])
  meta = []
  for x in m.metadata:
    name = _str(b, x.name)
    b.start_table()
    b.add_offset(0, name)
    b.add_scalar(1, "uint", int(x.buffer))
    meta.append(b.end_table())
  meta_v = b.create_offset_vector(meta) if meta else 0
  sigs = []
  for s in m.signatureDefs:
    def maps(items):
      offs = []
      for tm in items:
        name = _str(b, tm.name)
        b.start_table()
        b.add_offset(0, name)
        b.add_scalar(1, "uint", int(tm.tensorIndex))
        offs.append(b.end_table())
      return b.create_offset_vector(offs)
    i_v, o_v = maps(s.inputs), maps(s.outputs)
    key = _str(b, s.signatureKey)
    b.start_table()
    b.add_offset(0, i_v)
    b.add_offset(1, o_v)
    b.add_offset(2, key)
    b.add_scalar(4, "uint", int(s.subgraphIndex))
    sigs.append(b.end_table())
  sigs_v = b.create_offset_vector(sigs) if sigs else 0
  desc = _str(b, m.description)
  mbuf = _vec(b, m.metadataBuffer, np.int32)
  b.start_table()
  b.add_scalar(0, "uint", int(m.version))
  b.add_offset(1, codes_v)
  b.add_offset(2, subgraphs_v)
  b.add_offset(3, desc)
  b.add_offset(4, buffers_v)
  b.add_offset(5, mbuf)
  b.add_offset(6, meta_v)
  b.add_offset(7, sigs_v)
  head = b.finish(b.end_table(), FILE_IDENTIFIER)
  if not any(is_ext):
    return head
  r16 = lambda n: (n + 15) & ~15
  out = _uninitialised_bytearray(r16(len(head)) + sum(r16(len(p)) for p, e in zip(payloads, is_ext) if e))
  out[:len(head)] = head
  out[len(head):r16(len(head))] = bytes(r16(len(head)) - len(head))
  tables = fb.Table.root(head).table_vector(4)
  pos = r16(len(head))
  pieces = []
  for t, payload, ext in zip(tables, payloads, is_ext):
    if not ext:
      continue
    n = len(payload)
    struct.pack_into("<Q", out, t._field(1), pos)  # pylint: disable=protected-access
    struct.pack_into("<Q", out, t._field(2), n)    # pylint: disable=protected-access
    pieces.append((pos, payload))
    out[pos + n:r16(pos + n)] = bytes(r16(pos + n) - (pos + n))  # the alignment gap behind the payload
    pos = r16(pos + n)
  _copy_pieces(out, pieces)
  return out


def write_model(m: ModelT, path: str) -> None:
  with open(path, "wb") as f:
    f.write(write_model_to_bytes(m))
