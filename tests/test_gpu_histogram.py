"""GPU parity of the dynamic-range histogram against fixtures generated from
the reference's utils/histogram_utils.py (tests/golden/make_golden.py:gen_histogram).  Counts are
integers and the range bookkeeping is the reference's fp32 arithmetic, so everything is exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
Z = os.path.join(os.path.dirname(__file__), "golden", "histogram.npz")


def _check(z, prefix, h):
  for c, impl in enumerate(h._impls):
    np.testing.assert_array_equal(impl.counts, z[f"{prefix}_c{c}_counts"], err_msg=f"{prefix} c{c}")
    assert np.float32(impl.lower_bound) == np.float32(z[f"{prefix}_c{c}_lb"])
    assert np.float32(impl.bin_width) == np.float32(z[f"{prefix}_c{c}_bw"])
    assert np.float32(impl.global_min) == np.float32(z[f"{prefix}_c{c}_min"])
    assert np.float32(impl.global_max) == np.float32(z[f"{prefix}_c{c}_max"])
    assert impl.counts.dtype == np.int64


def test_histogram_fixtures(cuda):
  import torch
  from aeq_b200.utils import histogram_utils as hu
  z = np.load(Z)
  batches = [z[f"b{j}"] for j in range(5)]
  h = hu.DynamicHistogram(max_tensor_bins=2048)
  for j, b in enumerate(batches):
    h.add(b if j % 2 else torch.from_numpy(b).to(cuda))  # NumPy and device-resident batches
    _check(z, f"t_after{j}", h)
  finite = sum(int(np.isfinite(b).sum()) for b in batches)
  assert int(h.counts.sum()) == finite
  h2 = hu.DynamicHistogram(max_tensor_bins=256, initial_bin_width=0.05)
  for b in batches[:3]:
    h2.add(b)
  _check(z, "w", h2)
  hc = hu.DynamicHistogram(max_tensor_bins=2048, axis=0)
  for b in batches[:4]:
    hc.add(b)
  _check(z, "ch", hc)
  with pytest.raises(AttributeError, match="not supported for per-channel"):
    _ = hc.counts
  a, bb = hu.DynamicHistogram(max_tensor_bins=512), hu.DynamicHistogram(max_tensor_bins=512)
  a.add(batches[0]); a.add(batches[2])
  bb.add(batches[3])
  a.merge(bb)
  _check(z, "merged", a)
  # dictionary round trip
  back = hu.DynamicHistogram.from_dict(h.to_dict())
  np.testing.assert_array_equal(back.counts, h.counts)
  assert hu.DynamicHistogram().to_dict() == {}


def test_hist_kernel_edges(cuda):
  """Unaligned views, clipping to the first / last bin, NaN / inf handling with and without filter."""
  import torch
  from aeq_b200 import device
  rng = np.random.default_rng(5)
  x = rng.standard_normal(100003).astype(np.float32) * 3
  x[:5] = [np.nan, np.inf, -np.inf, 1e30, -1e30]
  big = torch.from_numpy(np.concatenate([np.zeros(1, np.float32), x])).to(cuda)
  lb, bw, nb = np.float32(-4.0), np.float32(0.03125), 300
  for finite_only in (True, False):
    got = device.hist_accumulate(big[1:], lb, bw, nb, finite_only).cpu().numpy()
    data = x[np.isfinite(x)] if finite_only else x
    with np.errstate(all="ignore"):
      idx = np.clip(np.floor((data - lb) / bw).astype(np.int32), 0, nb - 1)
    np.testing.assert_array_equal(got, np.bincount(idx, minlength=nb))
  # accumulation into an existing counter
  c = device.hist_accumulate(big[1:], lb, bw, nb, True)
  c2 = device.hist_accumulate(big[1:], lb, bw, nb, True, counts=c.clone())
  np.testing.assert_array_equal(c2.cpu().numpy(), 2 * c.cpu().numpy())


def test_histogram_calibration_func_vs_reference_class(cuda):
  """`histogram_calibrate` + `histogram_update` through the registry: the
  QSV's histogram equals the one the REFERENCE's DynamicHistogram builds from the same batches
  (utils/histogram_utils.py:396-452, run unmodified through oracle/refshim), min / max follow the
  reference's moving average (qsv_utils.py:43-68), and the percentile range feeds the min-max
  materialiser."""
  from oracle import aeq_oracle as O
  from oracle import refshim
  if not refshim.available():
    pytest.skip("no reference tree (run oracle/make_ref.py)")
  from aeq_b200 import _lib, algorithm_manager as am, qtyping
  from aeq_b200.algorithms.uniform_quantize import histogram_calibration as hc
  from tests import synthetic_graph
  ref_hu = refshim.ref("utils.histogram_utils")
  ref_qsv = refshim.ref("utils.qsv_utils")
  op, graph = synthetic_graph.fc_graph(O.synthetic_weight(16, 128, 0), batch=8)
  FC = qtyping.TFLOperationName.FULLY_CONNECTED
  cal = am.get_quantization_func(hc.ALGORITHM_KEY, FC, qtyping.QuantizeMode.CALIBRATE)
  upd = am.get_update_qsv_func(hc.ALGORITHM_KEY, FC)
  assert cal is hc.histogram_calibrate and upd is hc.histogram_update
  qsv, want_h, want_mm = {}, ref_hu.DynamicHistogram(max_tensor_bins=2048), {}
  for j in range(4):
    x = O.synthetic_activation((8, 128), 10 + j) * (1.0 + j)  # the range grows: bins double / compact
    x[0, :3] = [np.inf, -np.inf, np.nan]
    y = O.synthetic_activation((8, 16), 20 + j)
    before = _lib.load().aeqb_launch_count()
    new = cal(op, graph, {"input": x, "output": y})
    assert _lib.load().aeqb_launch_count() >= before + 3  # min/max batch + finite min/max + bin count
    assert set(new) == {"input", "output"} and int(new["input"]["num_samples"]) == 8
    for name in new:
      qsv[name] = upd(qsv.get(name, {}), new[name])
    hj = ref_hu.DynamicHistogram(max_tensor_bins=2048)  # the reference, batch by batch: add, then merge
    hj.add(x)
    want_h.merge(hj)
    with np.errstate(all="ignore"):
      fin = x[(x > -3e38) & (x < 3e38)]
      one = {"min": np.reshape(fin.min(), (1, 1)), "max": np.reshape(fin.max(), (1, 1))}
    want_mm = ref_qsv.moving_average_update(want_mm, one) if want_mm else one
  got = qsv["input"][hc.HISTOGRAM_KEY]["channels"][0]
  want = want_h.to_dict()["channels"][0]
  np.testing.assert_array_equal(got["hist_counts"], want["hist_counts"])
  assert np.float32(got["lower_bound"]) == np.float32(want["lower_bound"])
  assert np.float32(got["bin_width"]) == np.float32(want["bin_width"])
  np.testing.assert_array_equal(qsv["input"]["min"], want_mm["min"])
  np.testing.assert_array_equal(qsv["input"]["max"], want_mm["max"])
  # materialisation: percentile 100 == min-max on the same QSV; a central 60 % of the mass is narrower
  # than the moving-average min / max and narrows the range
  cfg = qtyping.TensorQuantizationConfig(8, symmetric=False)
  info = synthetic_graph.op_info(op, None)
  from aeq_b200.algorithms.uniform_quantize import naive_min_max_quantize as nmm
  p100 = hc.get_tensor_quant_params(info, cfg, None, qsv["input"])
  base = nmm.get_tensor_quant_params(info, cfg, None, {k: qsv["input"][k] for k in ("min", "max")})
  np.testing.assert_array_equal(p100.scale, base.scale)
  np.testing.assert_array_equal(p100.zero_point, base.zero_point)
  hc.set_percentile(60.0)
  try:
    p99 = hc.get_tensor_quant_params(info, cfg, None, qsv["input"])
  finally:
    hc.set_percentile(100.0)
  assert float(p99.scale.ravel()[0]) < float(p100.scale.ravel()[0])
