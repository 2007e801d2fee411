"""Host-side boundary: data contracts, registry semantics and error behaviour mirror the
reference (qtyping.py, algorithm_manager_api.py); the product path fails loudly without a GPU."""
import numpy as np
import pytest

from aeq_b200 import algorithm_manager as am
from aeq_b200 import algorithm_manager_api, qtyping
from aeq_b200.algorithms.uniform_quantize import naive_min_max_quantize as nmm
from aeq_b200.algorithms.uniform_quantize import uniform_quantize_tensor as uqt
from aeq_b200.algorithms.utils import common_utils
from aeq_b200.utils import qsv_utils
from tests import synthetic_graph as sg

Op = qtyping.TFLOperationName
G = qtyping.QuantGranularity


def test_enums_match_reference_values():
  assert Op.ALL_SUPPORTED.value == "*" and Op("FULLY_CONNECTED") is Op.FULLY_CONNECTED
  assert len(Op) == 52
  assert qtyping.QuantizeMode.CALIBRATE.value == 2 and qtyping.QuantizeMode.MATERIALIZE.value == 3
  assert [t.value for t in qtyping.QuantTransformation] == list(range(10))
  assert am.AlgorithmName.MIN_MAX_UNIFORM_QUANT == "min_max_uniform_quantize"
  assert {a.value for a in am.AlgorithmName} >= {"OCTAV", "MSE", "GPTQ", "HADAMARD_ROTATION",
                                                  "DECOMPOSED_HADAMARD_ROTATION", "no_quantize"}


def test_reference_enums_agree_when_mounted():
  from oracle import refshim
  if not refshim.available():
    pytest.skip("reference tree not mounted")
  rq = refshim.ref("qtyping")
  for name in ("TFLOperationName", "QuantizeMode", "OpExecutionMode", "ComputePrecision",
               "TensorDataType", "QuantGranularity", "QuantTransformation"):
    ours, theirs = getattr(qtyping, name), getattr(rq, name)
    assert {m.name: m.value for m in ours} == {m.name: m.value for m in theirs}, name
  ram = refshim.ref("algorithm_manager")
  assert {m.name: m.value for m in am.AlgorithmName} == {m.name: m.value for m in ram.AlgorithmName}


def test_tensor_config_is_hashable_cache_key_and_roundtrips():
  c1 = qtyping.TensorQuantizationConfig(4, True, G.BLOCKWISE_32, algorithm_params={"a": 1})
  c2 = qtyping.TensorQuantizationConfig.from_dict(
      {"num_bits": 4, "symmetric": True, "granularity": G.BLOCKWISE_32, "a": 1})
  assert c1 == c2 and hash(c1) == hash(c2)
  assert c1.to_dict()["algorithm_params"] == {"a": 1}
  legacy = qtyping.TensorQuantizationConfig.from_dict({"num_bits": 4, "block_size": 64})
  assert legacy.granularity == G.BLOCKWISE_64
  with pytest.raises(ValueError, match="Unsupported block size: 48"):
    qtyping.TensorQuantizationConfig.from_dict({"num_bits": 4, "block_size": 48})
  with pytest.raises(TypeError):
    c1.algorithm_params["b"] = 2


def test_op_config_validation_messages():
  act = qtyping.TensorQuantizationConfig(8, dtype=qtyping.TensorDataType.INT)
  with pytest.raises(ValueError, match="integer activation but float weights"):
    qtyping.OpQuantizationConfig(act, qtyping.TensorQuantizationConfig(16, dtype=qtyping.TensorDataType.FLOAT))
  with pytest.raises(ValueError, match="must be SRQ"):
    qtyping.OpQuantizationConfig(act, qtyping.TensorQuantizationConfig(8))
  ok = qtyping.OpQuantizationConfig(act, qtyping.TensorQuantizationConfig(8),
                                    compute_precision=qtyping.ComputePrecision.INTEGER)
  assert qtyping.OpQuantizationConfig.from_dict(ok.to_dict()) == ok


def test_uniform_quant_params_equality_is_array_aware():
  a = qtyping.UniformQuantParams(8, 0, np.array([1.0, 2.0]), np.array([0, 0]))
  b = qtyping.UniformQuantParams(8, 0, np.array([1.0, 2.0]), np.array([0, 0]))
  c = qtyping.UniformQuantParams(8, 0, np.array([1.0, 2.5]), np.array([0, 0]))
  assert a == b and a != c
  h1 = qtyping.UniformQuantParams.HadamardRotationParams(np.ones(4, np.int8), 4)
  h2 = qtyping.HadamardRotationParams(np.ones(4, np.int8), 4)
  assert h1 == h2
  with pytest.raises(Exception):
    a.scale = None  # frozen


def test_registry_roundtrip_and_errors():
  api = algorithm_manager_api.AlgorithmManagerApi()
  f = lambda *a, **k: "called"
  api.register_quantized_op("alg", Op.FULLY_CONNECTED, f, f, f)
  assert api.is_algorithm_registered("alg") and api.is_op_registered("alg", Op.FULLY_CONNECTED)
  assert not api.is_op_registered("alg", Op.CONV_2D) and not api.is_op_registered("nope", Op.CONV_2D)
  assert api.get_quantization_func("alg", Op.FULLY_CONNECTED, qtyping.QuantizeMode.MATERIALIZE) is f
  assert api.get_update_qsv_func("alg", Op.FULLY_CONNECTED) is qsv_utils.moving_average_update
  assert api.get_init_qsv_func("alg", Op.FULLY_CONNECTED) is f
  with pytest.raises(ValueError, match="Unsupported operation .* for Algorithm: alg"):
    api.get_quantization_func("alg", Op.CONV_2D, qtyping.QuantizeMode.MATERIALIZE)
  with pytest.raises(ValueError, match="Unregistered algorithm: nope"):
    api.get_supported_ops("nope")
  cfg = qtyping.OpQuantizationConfig(weight_tensor_config=qtyping.TensorQuantizationConfig(8))
  with pytest.raises(ValueError, match="Config checking function for  algorithm alg"):
    api.check_op_quantization_config("alg", Op.FULLY_CONNECTED, cfg)
  api.check_op_quantization_config(
      "alg", Op.FULLY_CONNECTED,
      qtyping.OpQuantizationConfig(weight_tensor_config=qtyping.TensorQuantizationConfig(8), skip_checks=True))


def test_module_registry_has_hot_path_ops():
  for alg in ("min_max_uniform_quantize",):
    ops = am.get_supported_ops(alg)
    assert Op.FULLY_CONNECTED in ops and Op.EMBEDDING_LOOKUP in ops
    m = am.get_quantization_func(alg, Op.FULLY_CONNECTED, qtyping.QuantizeMode.MATERIALIZE)
    assert m.func is common_utils.materialize_weight_op


def test_every_algorithm_key_is_registered():
  """All AlgorithmName keys of the reference (algorithm_manager.py:65-75) resolve here;
  `no_quantize` is a marker the callers test for, not a registered algorithm."""
  for alg in am.AlgorithmName:
    if alg == am.AlgorithmName.NO_QUANTIZE:
      continue
    assert am.is_algorithm_registered(alg), alg
    assert Op.FULLY_CONNECTED in am.get_supported_ops(alg)
  assert am.get_supported_ops(am.AlgorithmName.OSCAR) == [Op.FULLY_CONNECTED]
  assert am.get_update_qsv_func(am.AlgorithmName.OSCAR, Op.FULLY_CONNECTED) is (
      qsv_utils.oscar_and_moving_average_update)
  assert Op.DEPTHWISE_CONV_2D in am.get_supported_ops(am.AlgorithmName.FLOAT_CASTING)


def test_oscar_qsv_merge_matches_reference():
  from oracle import refshim
  a = {"min": np.float32([[-1.0]]), "max": np.float32([[2.0]]), "mu2": np.array([1.0, 4.0]), "num_samples": 3}
  b = {"min": np.float32([[-3.0]]), "max": np.float32([[1.0]]), "mu2": np.array([2.0, 0.5]), "num_samples": 1}
  got = qsv_utils.oscar_and_moving_average_update(a, b)
  np.testing.assert_array_equal(got["mu2"], (a["mu2"] * 3 + b["mu2"] * 1) / 4)
  assert got["num_samples"] == 4 and qsv_utils.oscar_and_moving_average_update({}, b) is b
  if refshim.available():
    want = refshim.ref("utils.qsv_utils").oscar_and_moving_average_update(a, b)
    for k in ("min", "max", "mu2"):
      np.testing.assert_array_equal(got[k], want[k])
    assert got["num_samples"] == want["num_samples"]


def test_blockwise_config_rules():
  blk = qtyping.TensorQuantizationConfig(4, True, G.BLOCKWISE_32)
  am.check_op_quantization_config("min_max_uniform_quantize", Op.FULLY_CONNECTED,
                                  qtyping.OpQuantizationConfig(weight_tensor_config=blk))
  with pytest.raises(ValueError, match="blockwise"):
    am.check_op_quantization_config("min_max_uniform_quantize", Op.CONV_2D,
                                    qtyping.OpQuantizationConfig(weight_tensor_config=blk))
  asym = qtyping.TensorQuantizationConfig(4, False, G.BLOCKWISE_32)
  with pytest.raises(ValueError, match="asymmetric"):
    am.check_op_quantization_config("min_max_uniform_quantize", Op.FULLY_CONNECTED,
                                    qtyping.OpQuantizationConfig(weight_tensor_config=asym))


def test_qsv_update_rules():
  q = qsv_utils.moving_average_update({}, {"min": np.array([[-1.0]]), "max": np.array([[1.0]])})
  assert q["min"][0, 0] == -1.0
  q = qsv_utils.moving_average_update({"min": np.float32(-10), "max": np.float32(10)},
                                      {"min": np.float32(-20), "max": np.float32(20)})
  assert q["min"] == pytest.approx(-10.5) and q["max"] == pytest.approx(10.5)
  u = qsv_utils.min_max_update({"min": np.array([-1.0]), "max": np.array([1.0])},
                               {"min": np.array([-0.5]), "max": np.array([2.0])})
  assert u["min"][0] == -1.0 and u["max"][0] == 2.0
  a = {"min": np.array([0.0]), "max": np.array([1.0]), "hessian": np.eye(2) * 2.0, "num_samples": 2}
  b = {"min": np.array([0.0]), "max": np.array([1.0]), "hessian": np.eye(2) * 8.0, "num_samples": 6}
  # the K x K Hessian merge is device work (aeqb_hessian_merge_f64): no GPU here -> it must raise
  with pytest.raises(RuntimeError, match="no CPU fallback"):
    qsv_utils.gptq_and_moving_average_update(a, b)
  assert qsv_utils.gptq_and_moving_average_update({}, b) is b


def test_layout_helpers():
  w = np.zeros((8, 64), np.float32)
  op, _ = sg.fc_graph(w)
  info = sg.op_info(op, qtyping.TensorQuantizationConfig(8, True, G.CHANNELWISE))
  assert common_utils.get_weight_quantized_dim(info, w, G.CHANNELWISE) == 0
  assert common_utils.get_weight_quantized_dim(info, w, G.BLOCKWISE_32) == 1
  assert common_utils.get_weight_quantized_dim(info, w, G.TENSORWISE) is None
  assert common_utils.get_reduce_dims(0, (8, 3, 3, 4)) == (1, 2, 3)
  assert common_utils.get_reduce_dims(None, (8, 4)) is None
  assert uqt.extract_block_size_from_granularity(G.BLOCKWISE_128) == 128
  assert uqt.get_quantized_range(uqt.IntType(4, True)) == (-8.0, 7.0)
  assert uqt.get_quantized_range(uqt.IntType(8, False)) == (0.0, 255.0)
  v, axis = uqt.reshape_data_for_blockwise(w, Op.FULLY_CONNECTED, G.BLOCKWISE_32)
  assert v.shape == (8, 2, 32) and axis == 2
  with pytest.raises(ValueError, match="is not divisible by block size 32"):
    uqt.reshape_data_for_blockwise(np.zeros((8, 48), np.float32), Op.FULLY_CONNECTED, G.BLOCKWISE_32)


def test_rank_and_shape_validation_messages():
  x = np.array([-3.0, 1.3, 2.4, 16.0], np.float32)
  bad = qtyping.UniformQuantParams(4, 0, np.array([[[1.2]]]), np.array([[-6]]))
  with pytest.raises(ValueError, match=r"Ranks of scales \(3\) and zps \(2\) must be the same as the tensor rank"):
    uqt._is_valid_quantization_params(x, bad)
  fixed = uqt.fix_quantization_params_rank(
      np.zeros((4, 8)), qtyping.UniformQuantParams(8, 0, np.ones(4), np.zeros(4, np.int8)))
  assert fixed.scale.shape == (4, 1) and fixed.zero_point.shape == (4, 1)


def test_missing_qsv_error_message():
  w = np.zeros((8, 64), np.float32)
  op, _ = sg.fc_graph(w)
  cfg = qtyping.TensorQuantizationConfig(8, True, G.CHANNELWISE)
  with pytest.raises(ValueError, match="not found in tensor_name_to_qsv"):
    nmm.get_tensor_quant_params(sg.op_info(op, cfg), cfg, None, None)


def test_product_path_fails_loudly_without_gpu():
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  w = np.ones((8, 64), np.float32)
  op, _ = sg.fc_graph(w)
  cfg = qtyping.TensorQuantizationConfig(8, True, G.CHANNELWISE)
  with pytest.raises(RuntimeError, match="no CPU fallback"):
    nmm.get_tensor_quant_params(sg.op_info(op, cfg), cfg, w, None)


def test_every_algorithm_and_the_quantizer_fail_loudly_without_gpu():
  """No algorithm has a host fallback: every weight path and the Quantizer front end raise on a
  box without a CUDA device instead of computing on the CPU."""
  import torch
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  from aeq_b200 import quantizer, recipe
  from aeq_b200.algorithms.nonlinear_quantize import float_casting
  from aeq_b200.algorithms.uniform_quantize import (dequantized_weight_recovery, gptq, hadamard_rotation,
                                                    mse, octav, oscar)
  from aeq_b200.utils import tfl_model as T
  from tests import tfl_fixtures
  w = np.ones((8, 64), np.float32)
  op, _ = sg.fc_graph(w)
  cfg = qtyping.TensorQuantizationConfig(8, True, G.CHANNELWISE)
  qsv = {"activation_tensor_qsv": {"hessian": np.eye(64), "num_samples": 1}}
  for mod, q in ((octav, None), (mse, None), (hadamard_rotation, None), (oscar, None),
                 (dequantized_weight_recovery, None), (gptq, qsv)):
    with pytest.raises((RuntimeError, ImportError), match="no CPU fallback|CUDA"):
      mod.get_tensor_quant_params(sg.op_info(op, cfg), cfg, w, q)
  with pytest.raises((RuntimeError, ImportError), match="no CPU fallback|CUDA"):
    float_casting.cast_weight(w)
  model = T.write_model_to_bytes(tfl_fixtures.fc_stack([w]))
  # the batched driver in front of the per-op walk is the first thing to need the device
  with pytest.raises((RuntimeError, ValueError), match="no CPU fallback|CUDA"):
    quantizer.Quantizer(model, recipe.dynamic_wi8_afp32()).quantize()


def test_product_never_imports_oracle():
  import os
  import re
  root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ai-edge-quantizer_b200")
  for d, _, files in os.walk(root):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        text = open(os.path.join(d, f)).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
        assert "refshim" not in text, f


def test_plugin_installs_into_reference_registry():
  """aeq_b200.plugin.install rebinds the reference's own materialisers to our arithmetic."""
  from oracle import refshim
  if not refshim.available():
    pytest.skip("reference tree not mounted")
  from aeq_b200 import plugin
  ram = refshim.ref("algorithm_manager")
  rq = ram.qtyping
  ref_had = ram.hadamard_rotation.get_tensor_quant_params
  ref_tu = refshim.ref("transformations.transformation_utils")
  ref_pack = ref_tu.pack_data
  bound = plugin.install(ram)
  try:
    assert {"min_max_uniform_quantize", "OCTAV", "MSE", "HADAMARD_ROTATION", "GPTQ", "OSCAR",
            "dequantized_weight_recovery"} <= set(bound)
    for key in ("min_max_uniform_quantize", "OCTAV", "MSE"):
      f = ram.get_quantization_func(key, rq.TFLOperationName.FULLY_CONNECTED,
                                    rq.QuantizeMode.MATERIALIZE)
      assert f.func.__name__ == "materialize_fc_conv"
      assert f.args[0].__module__.startswith("aeq_b200.")
    f = ram.get_quantization_func("min_max_uniform_quantize", rq.TFLOperationName.TRANSPOSE,
                                  rq.QuantizeMode.MATERIALIZE)
    assert f.func.__name__ == "materialize_transpose"
    assert ram.hadamard_rotation.get_tensor_quant_params is not ref_had
    # calibration functions (algorithm_manager_api.py:97-122) are the device twins ...
    FC, CAL = rq.TFLOperationName.FULLY_CONNECTED, rq.QuantizeMode.CALIBRATE
    for key, name in (("min_max_uniform_quantize", "min_max_calibrate"), ("OCTAV", "min_max_calibrate"),
                      ("HADAMARD_ROTATION", "min_max_calibrate"), ("GPTQ", "calibrate"),
                      ("OSCAR", "calibrate")):
      c = ram.get_quantization_func(key, FC, CAL)
      assert c.__module__.startswith("aeq_b200.") and c.__name__ == name, (key, c)
    assert ram.get_quantization_func("min_max_uniform_quantize", rq.TFLOperationName.TRANSPOSE, CAL
                                     ).__module__.startswith("aeq_b200.")
    # ... the QSV merges stay the reference's, and the pack step of QUANTIZE_TENSOR is rebound
    assert ram.get_update_qsv_func("GPTQ", FC).__module__.startswith("ai_edge_quantizer.")
    assert getattr(ref_tu.pack_data, "_aeqb200", False)
    assert np.array_equal(ref_tu.pack_data(8, np.arange(6, dtype=np.uint8)), np.arange(6, dtype=np.uint8))
  finally:
    plugin.uninstall()
  f = ram.get_quantization_func("OCTAV", rq.TFLOperationName.FULLY_CONNECTED,
                                rq.QuantizeMode.MATERIALIZE)
  assert f.args[0].__module__.startswith("ai_edge_quantizer.")
  assert ram.hadamard_rotation.get_tensor_quant_params is ref_had
  assert ref_tu.pack_data is ref_pack
  assert ram.get_quantization_func("GPTQ", rq.TFLOperationName.FULLY_CONNECTED,
                                   rq.QuantizeMode.CALIBRATE) is ram.gptq.calibrate
  # calibration / pack rebinding are opt-out
  plugin.install(ram, calibration=False, pack=False)
  try:
    assert ram.get_quantization_func("GPTQ", rq.TFLOperationName.FULLY_CONNECTED,
                                     rq.QuantizeMode.CALIBRATE) is ram.gptq.calibrate
    assert ref_tu.pack_data is ref_pack
  finally:
    plugin.uninstall()


def test_plugin_prefetch_walks_the_model_like_the_reference(monkeypatch):
  """plugin.prefetch mirrors params_generator's op walk using the reference's own modules; the
  arithmetic (GPU) is replaced by a recorder here."""
  import types
  from oracle import refshim
  if not refshim.available():
    pytest.skip("reference tree not mounted")
  from aeq_b200 import plugin, prefetch as pf
  rq = refshim.ref("qtyping")
  ram = refshim.ref("algorithm_manager")
  rfu = refshim.ref("utils.tfl_flatbuffer_utils")
  rcu = refshim.ref("algorithms.utils.common_utils")
  w = np.zeros((8, 16), np.float32)
  tensors = [types.SimpleNamespace(name=b"in", shape=[2, 16], buffer=0, type=0, quantization=None),
             types.SimpleNamespace(name=b"w", shape=[8, 16], buffer=1, type=0, quantization=None),
             types.SimpleNamespace(name=b"out", shape=[2, 8], buffer=0, type=0, quantization=None)]
  fc = rq.BuiltinOperator.FULLY_CONNECTED
  model = types.SimpleNamespace(
      operatorCodes=[types.SimpleNamespace(builtinCode=fc), types.SimpleNamespace(builtinCode="unknown")],
      buffers=[types.SimpleNamespace(data=None), types.SimpleNamespace(data=w.tobytes())],
      subgraphs=[types.SimpleNamespace(tensors=tensors, operators=[
          types.SimpleNamespace(opcodeIndex=0, inputs=[0, 1, -1], outputs=[2], builtinOptions=None),
          types.SimpleNamespace(opcodeIndex=1, inputs=[0], outputs=[2], builtinOptions=None)])])
  cfg = rq.OpQuantizationConfig(weight_tensor_config=rq.TensorQuantizationConfig(num_bits=8))

  class FakeRecipe:
    def get_quantization_configs(self, op_key, scope):
      assert op_key == rq.TFLOperationName.FULLY_CONNECTED and scope == "out;"
      return ram.AlgorithmName.MIN_MAX_UNIFORM_QUANT, cfg

  pgmod = types.ModuleType("fake_params_generator")
  pgmod.tfl_flatbuffer_utils, pgmod.algorithm_manager, pgmod.qtyping = rfu, ram, rq
  import sys
  monkeypatch.setitem(sys.modules, "fake_params_generator", pgmod)
  PG = type("ParamsGenerator", (), {"__module__": "fake_params_generator"})
  pg = PG()
  pg.float_model, pg._tensor_quant_params_cache = model, rcu.TensorQuantParamsCache()
  seen = {}

  def record(items, cache, make_params=None, get_tensor_data=None):
    seen["items"], seen["cache"] = list(items), cache
    seen["params"] = make_params(num_bits=8, quantized_dimension=0, scale=np.ones((8, 1), np.float32),
                                 zero_point=np.zeros((8, 1), np.int8), symmetric=True)
    seen["data"] = get_tensor_data(tensors[1], model.buffers)
    return {"quantized": len(seen["items"])}

  monkeypatch.setattr(pf, "prefetch_weights", record)
  assert plugin.prefetch(pg, FakeRecipe()) == {"quantized": 1}
  (info, graph), = seen["items"]
  assert info.op_name == rq.TFLOperationName.FULLY_CONNECTED and info.subgraph_op_index == 0
  assert graph.buffers is model.buffers and seen["cache"] is pg._tensor_quant_params_cache
  assert isinstance(seen["params"], rq.UniformQuantParams) and seen["data"].shape == (8, 16)


def test_histogram_calibration_host_logic():
  """Percentile range and QSV merge of the histogram calibration algorithm (host side only)."""
  from aeq_b200.algorithms.uniform_quantize import histogram_calibration as hc
  assert am.is_algorithm_registered(hc.ALGORITHM_KEY)
  assert am.get_update_qsv_func(hc.ALGORITHM_KEY, Op.FULLY_CONNECTED) is hc.histogram_update
  assert hc.ALGORITHM_KEY not in {a.value for a in am.AlgorithmName}  # the enum stays the reference's
  counts = np.zeros(100, np.int64)
  counts[10:90] = 10  # 800 values, uniform over bins 10..89
  ch = {"hist_counts": counts, "lower_bound": -5.0, "bin_width": 0.1, "min": np.array([-4.0]),
        "max": np.array([4.0])}
  hist = {"min": -4.0, "max": 4.0, "axis": None, "channels": [ch]}
  lo, hi = hc.percentile_range(hist, 100.0)
  assert lo == np.float32(-5.0 + 10 * 0.1) and hi == np.float32(-5.0 + 90 * 0.1)
  lo, hi = hc.percentile_range(hist, 95.0)  # 20 values trimmed per side = two bins
  assert lo == np.float32(-5.0 + 12 * 0.1) and hi == np.float32(-5.0 + 88 * 0.1)
  assert hc.percentile_range({}, 99.0) is None
  with pytest.raises(ValueError, match="percentile"):
    hc.set_percentile(40.0)
  a = {"min": np.array([[-1.0]], np.float32), "max": np.array([[2.0]], np.float32), hc.HISTOGRAM_KEY: hist}
  b = {"min": np.array([[-3.0]], np.float32), "max": np.array([[4.0]], np.float32), hc.HISTOGRAM_KEY: hist}
  assert hc.histogram_update({}, a)[hc.HISTOGRAM_KEY]["channels"][0]["hist_counts"].sum() == 800
  m = hc.histogram_update(a, b)
  assert m["min"] == np.float32(0.95) * np.float32(-1.0) + np.float32(1 - 0.95) * -3.0 or np.isclose(m["min"], -1.1)
  assert int(np.sum(m[hc.HISTOGRAM_KEY]["channels"][0]["hist_counts"])) == 1600
