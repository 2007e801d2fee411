"""CPU tests of the TFLite model reader / writer, the recipe manager and the op scope helper
(the data format either side of the hot path)."""
import dataclasses
import glob
import json
import os

import numpy as np
import pytest

from aeq_b200 import qtyping, recipe, recipe_manager
from aeq_b200.utils import flatbuffer_lite as fb
from aeq_b200.utils import tfl_flatbuffer_utils as fu
from aeq_b200.utils import tfl_model as T
from tests import tfl_fixtures

REF_MODELS = "/root/reference/ai_edge_quantizer/tests/models"
REF_RECIPES = "/root/reference/ai_edge_quantizer/recipes"


def _same(a, b, path="model"):
  if dataclasses.is_dataclass(a):
    assert type(a) is type(b), path
    for f in dataclasses.fields(a):
      if isinstance(a, T.RawTable) and f.name == "pos_mod8":
        continue
      x, y = getattr(a, f.name), getattr(b, f.name)
      if isinstance(a, T.RawTable) and f.name == "table":
        x, y = x[4:], y[4:]  # the first 4 bytes are the (position dependent) vtable offset
      _same(x, y, f"{path}.{f.name}")
  elif isinstance(a, list):
    assert len(a) == len(b), path
    for i, (x, y) in enumerate(zip(a, b)):
      _same(x, y, f"{path}[{i}]")
  elif isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
    if a is None or b is None:
      assert (a is None or a.size == 0) and (b is None or b.size == 0), path
    else:
      assert a.dtype == b.dtype and np.array_equal(a, b), path
  else:
    assert a == b, (path, a, b)


def test_builder_reader_roundtrip_all_field_kinds():
  b = fb.Builder(64)  # tiny initial buffer: exercises growth
  s = b.create_string("héllo")
  v64 = b.create_numpy_vector(np.array([-1, 2 ** 40], np.int64))
  vf = b.create_numpy_vector(np.array([1.5, -2.25, 3.0], np.float32))
  blob = b.create_byte_vector(bytes(range(37)), align=16)
  empty = b.create_numpy_vector(np.zeros(0, np.int32))
  b.start_table()
  b.add_scalar(0, "int", 7)
  b.add_scalar(1, "ulong", 2 ** 50)
  child = b.end_table()
  kids = b.create_offset_vector([child, child])
  b.start_table()
  b.add_scalar(0, "uint", 3)
  b.add_scalar(1, "byte", -5)
  b.add_scalar(2, "bool", True, False)
  b.add_scalar(3, "int", 0)          # default: not stored
  b.add_offset(4, s)
  b.add_offset(5, v64)
  b.add_offset(6, vf)
  b.add_offset(7, blob)
  b.add_offset(8, kids)
  b.add_offset(9, empty)
  b.add_scalar(12, "double", 0.125)
  data = b.finish(b.end_table(), b"TEST")
  assert fb.file_identifier(data) == b"TEST"
  t = fb.Table.root(data)
  assert t.scalar(0, "uint") == 3 and t.scalar(1, "byte") == -5 and t.scalar(2, "bool") is True
  assert not t.has(3) and t.scalar(3, "int", 42) == 42 and not t.has(10) and not t.has(99)
  assert t.string(4) == "héllo".encode()
  np.testing.assert_array_equal(t.scalar_vector(5, "long"), [-1, 2 ** 40])
  np.testing.assert_array_equal(t.scalar_vector(6, "float"), [1.5, -2.25, 3.0])
  blob_v = t.scalar_vector(7, "ubyte")
  assert bytes(blob_v) == bytes(range(37))
  assert (blob_v.__array_interface__["data"][0] - np.frombuffer(data, np.uint8).__array_interface__["data"][0]) % 16 == 0
  assert t.scalar_vector(5, "long").__array_interface__["data"][0] % 8 == np.frombuffer(data, np.uint8).__array_interface__["data"][0] % 8
  kids_t = t.table_vector(8)
  assert len(kids_t) == 2 and kids_t[0].pos == kids_t[1].pos
  assert kids_t[0].scalar(0, "int") == 7 and kids_t[0].scalar(1, "ulong") == 2 ** 50
  assert t.vector_len(9) == 0 and t.scalar_vector(9, "int").size == 0
  assert t.scalar(12, "double") == 0.125
  vt, tab, mod8 = kids_t[0].raw()   # scalar-only table: re-emit verbatim
  b2 = fb.Builder(32)
  b2.create_string("x")             # shift the alignment
  raw = b2.add_raw_table(vt, tab, mod8)
  b2.start_table()
  b2.add_offset(0, raw)
  d2 = b2.finish(b2.end_table())
  c2 = fb.Table.root(d2).table(0)
  assert c2.scalar(0, "int") == 7 and c2.scalar(1, "ulong") == 2 ** 50 and c2.pos % 8 == mod8


def test_synthetic_model_roundtrip():
  w0, w1 = np.arange(24, dtype=np.float32).reshape(4, 6), np.ones((3, 4), np.float32)
  m = tfl_fixtures.fc_stack([w0, w1], biases=[np.zeros(4, np.float32), None],
                            embedding=np.ones((10, 6), np.float32))
  data = T.write_model_to_bytes(m)
  assert data[4:8] == b"TFL3"
  m2 = T.read_model_from_bytes(data)
  _same(m, m2)
  g = m2.subgraphs[0]
  w = fu.get_tensor_data(g.tensors[int(g.operators[1].inputs[1])], m2.buffers)
  np.testing.assert_array_equal(w, w0)
  assert w.__array_interface__["data"][0] % 16 == np.frombuffer(data, np.uint8).__array_interface__["data"][0] % 16
  assert fu.get_op_scope(g.operators[1], g.tensors) == "layer0/out;"
  assert g.operators[1].builtinOptions.scalar(2, "bool", False) is False
  assert T.builtin_code(m2.operatorCodes[g.operators[0].opcodeIndex]) == T.BuiltinOperator.EMBEDDING_LOOKUP
  with pytest.raises(ValueError, match="TFL3"):
    T.read_model_from_bytes(b"\x00" * 64)


def test_external_buffers_bytes_do_not_depend_on_the_allocation(monkeypatch):
  """External payloads follow the flatbuffer in a bytearray that is allocated WITHOUT a zero fill
  (tfl_model._uninitialised_bytearray): every byte, alignment gaps included, must be written, so the
  result equals what a zero-filled and a 0xAA-filled allocation give, and reads back to the payloads."""
  from tests import tfl_fixtures
  rng = np.random.default_rng(3)
  ws = [rng.standard_normal((r, c)).astype(np.float32) for r, c in ((37, 129), (5, 1031), (64, 64), (3, 7))]
  m = tfl_fixtures.fc_stack(ws)
  for i, b in enumerate(m.buffers):  # odd payload sizes: gaps of every length behind them
    if b.data is not None and len(b.data) > 64:
      b.data = np.asarray(b.data).view(np.uint8)[:len(b.data) - (i % 16)].copy()
  got = T.write_model_to_bytes(m, external_buffers=True)
  monkeypatch.setattr(T, "_uninitialised_bytearray", lambda n: bytearray(n))
  zero = T.write_model_to_bytes(m, external_buffers=True)
  monkeypatch.setattr(T, "_uninitialised_bytearray", lambda n: bytearray(b"\xaa" * n))
  dirty = T.write_model_to_bytes(m, external_buffers=True)
  assert isinstance(got, bytearray) and got == zero == dirty
  back = T.read_model_from_bytes(got)
  for a, b in zip(m.buffers, back.buffers):
    if a.data is not None and len(a.data) > 0:
      assert bytes(np.asarray(a.data).view(np.uint8)) == bytes(np.asarray(b.data).view(np.uint8))


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference tree not mounted")
def test_reference_fixtures_roundtrip():
  """Every .tflite the reference ships parses and re-serialises to an identical object tree,
  the two StableHLO-composite models included."""
  files = sorted(glob.glob(os.path.join(REF_MODELS, "**", "*.tflite"), recursive=True))
  assert len(files) >= 70
  ok, rejected = 0, []
  for f in files:
    m = T.read_model(f)
    try:
      out = T.write_model_to_bytes(m)
    except NotImplementedError:
      rejected.append(os.path.basename(f))
      continue
    _same(m, T.read_model_from_bytes(out), os.path.basename(f))
    ok += 1
  assert rejected == [] and ok == len(files)
  comp = T.read_model(os.path.join(REF_MODELS, "sdpa_composite.tflite")).subgraphs[0].operators[5]
  assert comp.builtinOptions2Type == 21 and comp.builtinOptions2.name == b"odml.scaled_dot_product_attention"
  assert comp.builtinOptions2.decompositionSubgraphIndex == 1 and comp.builtinOptions2.compositeAttributes.size == 28
  m = T.read_model(os.path.join(REF_MODELS, "single_fc.tflite"))
  g = m.subgraphs[0]
  op = g.operators[0]
  assert T.builtin_code(m.operatorCodes[op.opcodeIndex]) == 9 and list(op.inputs) == [0, 2, 1]
  assert list(g.tensors[2].shape) == [16, 8] and len(m.buffers[g.tensors[2].buffer].data) == 512
  assert m.signatureDefs[0].signatureKey == b"serving_default"
  q = T.read_model(os.path.join(REF_MODELS, "mnist_quantized.tflite"))
  qt = [t for t in q.subgraphs[0].tensors if t.type == T.TensorType.INT8 and t.quantization is not None
        and t.quantization.scale is not None and t.quantization.scale.size > 1]
  assert qt and qt[0].quantization.zeroPoint.dtype == np.int64


def test_recipe_manager_and_recipes():
  got = recipe.dynamic_wi8_afp32()
  assert got == [{
      "regex": ".*", "operation": "*", "algorithm_key": "min_max_uniform_quantize",
      "op_config": {"weight_tensor_config": {"num_bits": 8, "symmetric": True,
                                             "granularity": "CHANNELWISE", "dtype": "INT"},
                    "compute_precision": "INTEGER", "explicit_dequantize": False,
                    "skip_checks": False, "min_weight_elements": 0}}]
  if os.path.isdir(REF_RECIPES):
    assert got == json.load(open(os.path.join(REF_RECIPES, "dynamic_wi8_afp32_recipe.json")))
  rm = recipe_manager.RecipeManager()
  rm.load_quantization_recipe(got)
  Op = qtyping.TFLOperationName
  alg, cfg = rm.get_quantization_configs(Op.FULLY_CONNECTED, "dense/out;")
  assert alg == "min_max_uniform_quantize" and cfg.weight_tensor_config.num_bits == 8
  assert rm.get_quantization_configs(Op.SOFTMAX, "x;")[0] == "no_quantize"
  # later scopes override earlier ones; NO_QUANTIZE carves a layer out
  rm.add_quantization_config("layer1/", Op.FULLY_CONNECTED, algorithm_key="no_quantize")
  rm.add_dynamic_config("layer2/", Op.FULLY_CONNECTED, 4, qtyping.QuantGranularity.BLOCKWISE_32)
  assert rm.get_quantization_configs(Op.FULLY_CONNECTED, "layer1/out;")[0] == "no_quantize"
  _, c2 = rm.get_quantization_configs(Op.FULLY_CONNECTED, "layer2/out;")
  assert c2.weight_tensor_config.granularity == qtyping.QuantGranularity.BLOCKWISE_32
  assert rm.get_quantization_configs(Op.FULLY_CONNECTED, "layer0/out;")[1].weight_tensor_config.num_bits == 8
  rm2 = recipe_manager.RecipeManager()
  rm2.load_quantization_recipe(rm.get_quantization_recipe())   # JSON round trip
  assert rm2.get_quantization_recipe() == rm.get_quantization_recipe()
  assert not rm.need_calibration()
  with pytest.raises(ValueError, match="Unsupported algorithm key: nope"):  # recipe_manager.py:113-116
    rm.add_quantization_config(".*", Op.FULLY_CONNECTED, algorithm_key="nope")
  # A `*` entry is resolved per op at lookup: a blockwise recipe leaves CONV_2D / CONV_2D_TRANSPOSE
  # float (their config check raises, recipe_manager.py:185-198) instead of handing them a
  # BLOCKWISE_32 config that fails later in the materialiser.
  rb = recipe_manager.RecipeManager()
  rb.load_quantization_recipe(recipe.dynamic_wi4b32_afp32())
  assert rb.get_quantization_configs(Op.FULLY_CONNECTED, "a;")[0] == "min_max_uniform_quantize"
  assert rb.get_quantization_configs(Op.EMBEDDING_LOOKUP, "a;")[0] == "min_max_uniform_quantize"
  for op in (Op.CONV_2D, Op.CONV_2D_TRANSPOSE, Op.DEPTHWISE_CONV_2D):
    alg, cfg = rb.get_quantization_configs(op, "a;")
    assert alg == "no_quantize" and cfg.weight_tensor_config is None, op
  assert rb.get_quantization_recipe() == recipe.dynamic_wi4b32_afp32()  # the `*` entry stays one entry
  # a named op replaces its own entry in place, a later `*` replaces the whole scope
  rb.add_dynamic_config(".*", Op.FULLY_CONNECTED, 8)
  assert rb.get_quantization_configs(Op.FULLY_CONNECTED, "a;")[1].weight_tensor_config.num_bits == 8
  assert len(rb.get_quantization_recipe()) == 2
  rb.add_dynamic_config(".*", Op.ALL_SUPPORTED, 4)
  assert len(rb.get_quantization_recipe()) == 1
  with pytest.raises(ValueError, match="Unsupported op for blockwise quantization"):
    rb.add_dynamic_config(".*", Op.CONV_2D, 4, qtyping.QuantGranularity.BLOCKWISE_32)
  assert recipe.dynamic_wi4b32_afp32()[0]["op_config"]["weight_tensor_config"]["granularity"] == "BLOCKWISE_32"
  assert recipe.weight_only_wi8_afp32()[0]["op_config"]["explicit_dequantize"] is True
