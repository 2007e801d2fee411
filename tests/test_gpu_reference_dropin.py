"""The drop-in boundary proved on hardware against the UNMODIFIED reference.

`oracle/refshim` imports the reference's own modules — from /root/reference in the build
container, from the travelling byte-for-byte copy oracle/_ref (oracle/make_ref.py) on the GPU box.
Each test runs the reference's registered function first (pure NumPy: the expected value), then
`aeq_b200.plugin.install(reference algorithm_manager)` and the SAME registry entry again — now
the reference's materialiser / calibrator / transformation running our device arithmetic through
`plugin._adapt` — and compares.

Covers SURVEY.md §8(b)'s three Python callers: `get_tensor_quant_params` (through
common_quantize.materialize_fc_conv :519, materialize_embedding_lookup, the Hadamard module seam),
the calibration functions (algorithm_manager_api.py:97-122) and the pack step of
transformations/quantize_tensor.py:150-224.
"""
import types

import numpy as np
import pytest

from oracle import aeq_oracle as O
from oracle import refshim

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refshim.available(), reason="no reference tree (run oracle/make_ref.py)")]


def _tensor(name, shape, buffer=0, ttype=0):
  return types.SimpleNamespace(name=name.encode(), shape=list(shape), buffer=buffer, type=ttype,
                               quantization=None)


@pytest.fixture
def R(cuda):
  """The reference's modules, with our plug-in removed again after every test."""
  from aeq_b200 import plugin
  ns = types.SimpleNamespace()
  ns.am = refshim.ref("algorithm_manager")
  ns.q = ns.am.qtyping
  ns.cu = refshim.ref("algorithms.utils.common_utils")
  ns.qt = refshim.ref("transformations.quantize_tensor")
  ns.tu = refshim.ref("transformations.transformation_utils")
  ns.plugin = plugin
  yield ns
  plugin.uninstall()


def _graph(R, w, op_name="FULLY_CONNECTED", in_shape=None, out_shape=None, adj_y=None):
  w_ro = w.copy()
  w_ro.setflags(write=False)
  buffers = [types.SimpleNamespace(data=None), types.SimpleNamespace(data=w_ro.tobytes())]
  in_shape = in_shape or (4, w.shape[-1])
  out_shape = out_shape or (4, w.shape[0])
  tensors = [_tensor("input", in_shape), _tensor("weight", w.shape, 1), _tensor("output", out_shape)]
  opts = None if adj_y is None else types.SimpleNamespace(adjY=adj_y, adjX=False)
  inputs = [1, 0, -1][:2] if op_name == "EMBEDDING_LOOKUP" else [0, 1, -1]
  if op_name == "EMBEDDING_LOOKUP":
    tensors[0] = _tensor("input", (4,), 0, 2)
    inputs = [0, 1]
  op = types.SimpleNamespace(inputs=inputs, outputs=[2], builtinOptions=opts)
  return op, R.q.GraphInfo(subgraph_tensors=tensors, buffers=buffers)


def _materialize(R, alg, op_name, op, graph, cfg, qsvs=None, **op_cfg):
  info = R.q.OpInfo(op, R.q.TFLOperationName(op_name), 0,
                    R.q.OpQuantizationConfig(weight_tensor_config=cfg, **op_cfg))
  fn = R.am.get_quantization_func(alg, R.q.TFLOperationName(op_name), R.q.QuantizeMode.MATERIALIZE)
  with np.errstate(all="ignore"):
    out = fn(op_info=info, graph_info=graph, tensor_name_to_qsv=qsvs or {},
             tensor_quant_params_cache=R.cu.TensorQuantParamsCache())
  return {t.tensor_name: t for t in out}


def _cfg(R, bits, gran, sym=True, **params):
  return R.q.TensorQuantizationConfig(num_bits=bits, symmetric=sym,
                                      granularity=getattr(R.q.QuantGranularity, gran),
                                      algorithm_params=params)


def _same_params(a, b, exact=True, q_mismatch=0.0, scale_rtol=0.0):
  assert type(a) is type(b), (type(a), type(b))  # the REFERENCE's dataclass comes back
  assert a.num_bits == b.num_bits and a.symmetric == b.symmetric and a.block_size == b.block_size
  assert a.quantized_dimension == b.quantized_dimension
  assert a.scale.shape == b.scale.shape and a.scale.dtype == b.scale.dtype
  assert a.zero_point.shape == b.zero_point.shape
  assert a.quantized_data.shape == b.quantized_data.shape and a.quantized_data.dtype == b.quantized_data.dtype
  if exact:
    np.testing.assert_array_equal(a.scale, b.scale)
    np.testing.assert_array_equal(a.zero_point, b.zero_point)
    np.testing.assert_array_equal(a.quantized_data, b.quantized_data)
  else:
    np.testing.assert_allclose(a.scale, b.scale, rtol=scale_rtol)
    d = np.abs(a.quantized_data.astype(int) - b.quantized_data.astype(int))
    assert (d > 0).mean() <= q_mismatch and d.max() <= 1, ((d > 0).mean(), d.max())


CASES = [
    ("FULLY_CONNECTED", (96, 512), 8, "CHANNELWISE", True, dict(compute_precision="INTEGER")),
    ("FULLY_CONNECTED", (96, 512), 8, "CHANNELWISE", False, dict(compute_precision="INTEGER")),
    ("FULLY_CONNECTED", (64, 1024), 4, "BLOCKWISE_32", True, dict(compute_precision="INTEGER")),
    ("FULLY_CONNECTED", (64, 1024), 4, "CHANNELWISE", True, dict(explicit_dequantize=True)),  # weight-only
    ("FULLY_CONNECTED", (40, 256), 8, "TENSORWISE", True, dict(compute_precision="INTEGER")),
    ("EMBEDDING_LOOKUP", (300, 128), 4, "BLOCKWISE_64", True, dict(compute_precision="INTEGER")),
    ("CONV_2D", (16, 3, 3, 8), 8, "CHANNELWISE", True, dict(compute_precision="INTEGER")),
    ("DEPTHWISE_CONV_2D", (1, 3, 3, 64), 8, "CHANNELWISE", True, dict(compute_precision="INTEGER")),
]


@pytest.mark.parametrize("op_name,shape,bits,gran,sym,op_cfg", CASES)
def test_reference_materialiser_runs_device_min_max(R, op_name, shape, bits, gran, sym, op_cfg):
  """common_quantize.materialize_fc_conv / materialize_embedding_lookup of the reference, before
  and after plugin.install: identical transformations and bit-identical parameters."""
  from aeq_b200 import _lib
  w = O.synthetic_weight(int(np.prod(shape[:-1])), shape[-1], index=bits + len(shape)).reshape(shape)
  op_cfg = dict(op_cfg)
  if "compute_precision" in op_cfg:
    op_cfg["compute_precision"] = getattr(R.q.ComputePrecision, op_cfg["compute_precision"])
  alg = R.am.AlgorithmName.MIN_MAX_UNIFORM_QUANT
  in_shape = (1, 8, 8, shape[-1]) if "CONV" in op_name else None
  out_shape = (1, 8, 8, shape[0] if op_name == "CONV_2D" else shape[-1]) if "CONV" in op_name else None
  op, graph = _graph(R, w, op_name, in_shape, out_shape)
  want = _materialize(R, alg, op_name, op, graph, _cfg(R, bits, gran, sym), **op_cfg)
  assert "min_max_uniform_quantize" in R.plugin.install(R.am)
  before = _lib.load().aeqb_launch_count()
  got = _materialize(R, alg, op_name, op, graph, _cfg(R, bits, gran, sym), **op_cfg)
  assert _lib.load().aeqb_launch_count() > before, "the device path did not run"
  assert set(got) == set(want)
  for name in want:
    for side in ("consumers", "producer"):
      a, b = getattr(got[name], side), getattr(want[name], side)
      if b is None:
        assert a is None
        continue
      a, b = (a, b) if side == "consumers" else ([a], [b])
      assert [x.transformations for x in a] == [x.transformations for x in b]
  _same_params(got["weight"].consumers[0].parameters, want["weight"].consumers[0].parameters)


def test_batch_matmul_constant_weight_both_orientations(R):
  """BATCH_MATMUL quantises rank-1 (or rank-2 with adj_y): the swap-axes path."""
  alg = R.am.AlgorithmName.MIN_MAX_UNIFORM_QUANT
  for adj_y in (False, True):
    w = O.synthetic_weight(2 * 48, 80, index=9).reshape(2, 48, 80)
    op, graph = _graph(R, w, "BATCH_MATMUL", (2, 4, 48), (2, 4, 80), adj_y=adj_y)
    cfg = _cfg(R, 8, "CHANNELWISE")
    want = _materialize(R, alg, "BATCH_MATMUL", op, graph, cfg, compute_precision=R.q.ComputePrecision.INTEGER)
    R.plugin.install(R.am)
    got = _materialize(R, alg, "BATCH_MATMUL", op, graph, cfg, compute_precision=R.q.ComputePrecision.INTEGER)
    R.plugin.uninstall()
    wp, gp = want["weight"].consumers[0].parameters, got["weight"].consumers[0].parameters
    assert wp.quantized_dimension == (1 if adj_y else 2)
    _same_params(gp, wp)


@pytest.mark.parametrize("alg_name,gran,bits", [("OCTAV", "CHANNELWISE", 4), ("OCTAV", "BLOCKWISE_32", 4),
                                                ("MSE", "CHANNELWISE", 8)])
def test_reference_materialiser_runs_device_octav_mse(R, alg_name, gran, bits):
  w = O.synthetic_weight(64, 1024, index=31)
  op, graph = _graph(R, w)
  alg = R.am.AlgorithmName(alg_name)
  cfg = _cfg(R, bits, gran)
  want = _materialize(R, alg, "FULLY_CONNECTED", op, graph, cfg, compute_precision=R.q.ComputePrecision.INTEGER)
  assert alg_name in R.plugin.install(R.am)
  got = _materialize(R, alg, "FULLY_CONNECTED", op, graph, cfg, compute_precision=R.q.ComputePrecision.INTEGER)
  # order-dependent fp32 sums: scales <= 1e-6 relative (one bf16 ulp for blockwise), |dq| <= 1
  _same_params(got["weight"].consumers[0].parameters, want["weight"].consumers[0].parameters, exact=False,
               q_mismatch=2e-3, scale_rtol=1e-6 if gran == "CHANNELWISE" else 8e-3)


def test_hadamard_module_seam(R):
  """hadamard_rotation's materialisers call the module's own get_tensor_quant_params (:253):
  the seam is that attribute; the reference's custom-op materialiser keeps its graph edits."""
  w = O.synthetic_weight(32, 256, index=5)
  op, graph = _graph(R, w)
  alg = R.am.AlgorithmName.HADAMARD_ROTATION
  cfg = _cfg(R, 4, "CHANNELWISE")
  want = _materialize(R, alg, "FULLY_CONNECTED", op, graph, cfg, compute_precision=R.q.ComputePrecision.INTEGER)
  assert "HADAMARD_ROTATION" in R.plugin.install(R.am)
  got = _materialize(R, alg, "FULLY_CONNECTED", op, graph, cfg, compute_precision=R.q.ComputePrecision.INTEGER)
  assert ([t.transformations for t in got["input"].consumers]
          == [t.transformations for t in want["input"].consumers])  # INSERT_HADAMARD_ROTATION kept
  wp, gp = want["weight"].consumers[0].parameters, got["weight"].consumers[0].parameters
  assert type(gp.hadamard) is type(wp.hadamard) and gp.hadamard.hadamard_size == wp.hadamard.hadamard_size == 256
  np.testing.assert_array_equal(gp.hadamard.random_binary_vector, wp.hadamard.random_binary_vector)
  _same_params(gp, wp, exact=False, q_mismatch=5e-3, scale_rtol=2e-6)


def test_reference_calibration_functions_run_on_device(R):
  """The CALIBRATE registry entries after install are the device functions and return the
  reference's numbers: min / max with the (-3e38, 3e38) filter bit-exact, num_samples, and for
  GPTQ the float64 Hessian (gptq.py:100-106) within the sgemm-order tolerance."""
  from aeq_b200 import _lib
  w = O.synthetic_weight(32, 256, 1)
  op, graph = _graph(R, w, in_shape=(6, 40, 256), out_shape=(6, 40, 32))
  x = O.synthetic_activation((6, 40, 256), 3)
  x[0, 0, :4] = [-np.inf, 3.2e38, -3.3e38, 77.0]  # padding constants: filtered from min / max
  y = O.synthetic_activation((6, 40, 32), 4)
  content = {"input": x, "output": y}
  FC, CAL = R.q.TFLOperationName.FULLY_CONNECTED, R.q.QuantizeMode.CALIBRATE
  names = (R.am.AlgorithmName.MIN_MAX_UNIFORM_QUANT, R.am.AlgorithmName.OCTAV, R.am.AlgorithmName.GPTQ,
           R.am.AlgorithmName.HADAMARD_ROTATION)
  with np.errstate(all="ignore"):
    want = {a: R.am.get_quantization_func(a, FC, CAL)(op, graph, content) for a in names}
  theirs = {a: R.am.get_quantization_func(a, FC, CAL) for a in names}
  R.plugin.install(R.am)
  for a in names:
    fn = R.am.get_quantization_func(a, FC, CAL)
    assert fn is not theirs[a] and fn.__module__.startswith("aeq_b200."), (a, fn)
    before = _lib.load().aeqb_launch_count()
    got = fn(op, graph, content)
    assert _lib.load().aeqb_launch_count() > before
    assert set(got) == set(want[a]) == {"input", "output"}
    for name in got:
      for key in ("min", "max"):
        np.testing.assert_array_equal(got[name][key], want[a][name][key])
        assert got[name][key].shape == want[a][name][key].shape == (1, 1, 1)
      assert int(got[name]["num_samples"]) == int(want[a][name]["num_samples"]) == 6
    if a == R.am.AlgorithmName.GPTQ:
      h, hw = got["input"]["hessian"], want[a]["input"]["hessian"]
      assert h.dtype == hw.dtype == np.float64 and h.shape == hw.shape == (256, 256)
      # the 3.2e38 entry overflows fp32 products to inf / NaN in both (gptq_test.py:50-114)
      np.testing.assert_array_equal(np.isfinite(h), np.isfinite(hw))
      ok = np.isfinite(hw)
      np.testing.assert_allclose(h[ok], hw[ok], rtol=0, atol=4e-6 * np.abs(hw[ok]).max())
  # the QSV merge registered for GPTQ stays the reference's and accepts our QSVs
  upd = R.am.get_update_qsv_func(R.am.AlgorithmName.GPTQ, FC)
  x2 = O.synthetic_activation((2, 40, 256), 8)
  fn = R.am.get_quantization_func(R.am.AlgorithmName.GPTQ, FC, CAL)
  a, b = fn(op, graph, {"input": x2, "output": y[:2]}), fn(op, graph, {"input": x2 * 2, "output": y[:2]})
  merged = upd(a["input"], b["input"])
  assert int(merged["num_samples"]) == 4 and merged["hessian"].shape == (256, 256)
  R.plugin.uninstall()
  assert R.am.get_quantization_func(names[0], FC, CAL) is theirs[names[0]]


@pytest.mark.parametrize("bits,gran", [(4, "CHANNELWISE"), (4, "BLOCKWISE_32"), (2, "CHANNELWISE"), (8, "CHANNELWISE")])
def test_quantize_tensor_transformation_packs_on_device(R, bits, gran):
  """transformations/quantize_tensor.quantize_tensor of the reference (what TransformationPerformer
  registers for QUANTIZE_TENSOR): identical buffer bytes, tensor type and quantisation record
  whether `pack_data` is the reference's NumPy or the rebound device pack — first with integers
  the reference produced (aeqb_pack_bits runs), then with integers our kernels produced (the
  packed bytes of the fused pass are found, no second device trip)."""
  from aeq_b200 import _lib
  nmm = refshim.ref("algorithms.uniform_quantize.naive_min_max_quantize")
  w = O.synthetic_weight(48, 256, index=bits)
  cfg = _cfg(R, bits, gran)
  info = R.q.OpInfo(types.SimpleNamespace(inputs=[0, 1, -1], outputs=[2]), R.q.TFLOperationName.FULLY_CONNECTED,
                    0, R.q.OpQuantizationConfig(weight_tensor_config=cfg))

  def run(params):
    tensors = [_tensor("input", (4, 256)), _tensor("weight", w.shape, 1), _tensor("output", (4, 48))]
    model = types.SimpleNamespace(buffers=[types.SimpleNamespace(data=None, offset=0, size=0),
                                           types.SimpleNamespace(data=w.tobytes(), offset=0, size=0)])
    sub = types.SimpleNamespace(tensors=tensors)
    R.qt.quantize_tensor(R.tu.TransformationInput(1, model, sub, -1, [0], params, {}))
    return np.asarray(model.buffers[1].data).view(np.uint8).copy(), tensors[1], model, sub

  ref_params = nmm.get_tensor_quant_params(info, cfg, w, None)
  want_bytes, want_t, want_model, want_sub = run(ref_params)
  R.plugin.install(R.am)
  assert getattr(R.tu.pack_data, "_aeqb200", False)
  lib = _lib.load()
  before = lib.aeqb_launch_count()
  got_bytes, got_t, _, _ = run(ref_params)                 # reference integers -> device pack
  assert (lib.aeqb_launch_count() > before) == (bits < 8)
  np.testing.assert_array_equal(got_bytes, want_bytes)
  ours = R.am.naive_min_max_quantize  # the module is untouched; the registry entry is ours:
  fn = R.am.get_quantization_func(R.am.AlgorithmName.MIN_MAX_UNIFORM_QUANT, R.q.TFLOperationName.FULLY_CONNECTED,
                                  R.q.QuantizeMode.MATERIALIZE)
  op, graph = _graph(R, w)
  out = fn(op_info=R.q.OpInfo(op, R.q.TFLOperationName.FULLY_CONNECTED, 0,
                              R.q.OpQuantizationConfig(weight_tensor_config=cfg,
                                                       compute_precision=R.q.ComputePrecision.INTEGER)),
           graph_info=graph, tensor_name_to_qsv={}, tensor_quant_params_cache=R.cu.TensorQuantParamsCache())
  dev_params = [t for t in out if t.tensor_name == "weight"][0].consumers[0].parameters
  before = lib.aeqb_launch_count()
  got2_bytes, got2_t, got2_model, got2_sub = run(dev_params)  # device integers: packed bytes remembered
  if bits < 8:
    assert lib.aeqb_launch_count() == before, "the fused pass already packed these integers"
  np.testing.assert_array_equal(got2_bytes, want_bytes)
  assert got2_t.type == want_t.type == got_t.type
  np.testing.assert_array_equal(got2_t.quantization.scale if gran == "CHANNELWISE" else 0,
                                want_t.quantization.scale if gran == "CHANNELWISE" else 0)
  if gran != "CHANNELWISE":  # fp16 scale constant appended to the model (quantize_tensor.py:107-147)
    assert len(got2_model.buffers) == len(want_model.buffers) == 3
    np.testing.assert_array_equal(np.asarray(got2_model.buffers[2].data).view(np.uint8),
                                  np.asarray(want_model.buffers[2].data).view(np.uint8))
  del ours
  R.plugin.uninstall()
  assert not getattr(R.tu.pack_data, "_aeqb200", False)


def test_error_behaviour_through_the_reference_wrapper(R):
  """A non-divisible blockwise weight raises the reference's wrapped ValueError
  (common_utils.py:275-279) with the reference's message text, from our code."""
  w = O.synthetic_weight(8, 48, 1)  # 48 % 32 != 0
  op, graph = _graph(R, w)
  R.plugin.install(R.am)
  with pytest.raises(ValueError, match=r"Failed to get quantization parameters for tensor: weight\. Error: .*not"
                                        r" divisible by block size 32"):
    _materialize(R, R.am.AlgorithmName.MIN_MAX_UNIFORM_QUANT, "FULLY_CONNECTED", op, graph,
                 _cfg(R, 4, "BLOCKWISE_32"), compute_precision=R.q.ComputePrecision.INTEGER)


def test_histogram_calibration_registered_in_the_reference_registry(R):
  """plugin.install_histogram: the reference's registry resolves our key for every op its min-max
  algorithm covers; CALIBRATE returns the reference's min / max plus a histogram equal to the one the
  reference's DynamicHistogram builds, the update func merges, and MATERIALIZE (the reference's
  materialize_fc_conv over our adapted get_tensor_quant_params) quantises the weight exactly as the
  reference's min-max algorithm does."""
  ref_hu = refshim.ref("utils.histogram_utils")
  key = R.plugin.install_histogram(R.am)
  FC = R.q.TFLOperationName.FULLY_CONNECTED
  assert R.am.is_op_registered(key, FC)
  assert set(R.am.get_supported_ops(key)) == set(R.am.MIN_MAX_OP_NAME_MATERIALIZE_FUNC_DICT)
  w = O.synthetic_weight(32, 256, 5)
  op, graph = _graph(R, w, in_shape=(6, 40, 256), out_shape=(6, 40, 32))
  x, y = O.synthetic_activation((6, 40, 256), 3), O.synthetic_activation((6, 40, 32), 4)
  cal = R.am.get_quantization_func(key, FC, R.q.QuantizeMode.CALIBRATE)
  want_mm = R.am.get_quantization_func(R.am.AlgorithmName.MIN_MAX_UNIFORM_QUANT, FC, R.q.QuantizeMode.CALIBRATE)(
      op, graph, {"input": x, "output": y})
  got = cal(op, graph, {"input": x, "output": y})
  for name in ("input", "output"):
    np.testing.assert_array_equal(got[name]["min"], want_mm[name]["min"])
    np.testing.assert_array_equal(got[name]["max"], want_mm[name]["max"])
  h = ref_hu.DynamicHistogram(max_tensor_bins=2048)
  h.add(x)
  np.testing.assert_array_equal(got["input"]["histogram"]["channels"][0]["hist_counts"],
                                h.to_dict()["channels"][0]["hist_counts"])
  upd = R.am.get_update_qsv_func(key, FC)
  merged = upd(got["input"], cal(op, graph, {"input": x * 2, "output": y})["input"])
  h2 = ref_hu.DynamicHistogram(max_tensor_bins=2048)
  h2.add(x * 2)
  h.merge(h2)  # the reference's resampled merge (it may drop a few counts to rounding: so does ours)
  np.testing.assert_array_equal(merged["histogram"]["channels"][0]["hist_counts"],
                                h.to_dict()["channels"][0]["hist_counts"])
  # weight materialisation through the reference's own materialiser, against its min-max algorithm
  op_cfg = dict(compute_precision=R.q.ComputePrecision.INTEGER)
  want = _materialize(R, R.am.AlgorithmName.MIN_MAX_UNIFORM_QUANT, "FULLY_CONNECTED", op, graph,
                      _cfg(R, 8, "CHANNELWISE"), **op_cfg)
  got_p = _materialize(R, key, "FULLY_CONNECTED", op, graph, _cfg(R, 8, "CHANNELWISE"), **op_cfg)
  assert set(got_p) == set(want)
  _same_params(got_p["weight"].consumers[0].parameters, want["weight"].consumers[0].parameters)
