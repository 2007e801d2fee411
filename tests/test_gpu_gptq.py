"""GPU parity of the GPTQ path (Hessian, damped inverse, OBS column loop) against the
reference-generated fixtures in tests/golden/gptq.npz, the reference's literal test vectors
(gptq_test.py:50-114, :214-299, :331-398) and the oracle.

Tolerances: X^T X accumulates in fp32 in a different order than sgemm; the
triangular inverse and L^-T L^-1 are fp32 with a different blocking than LAPACK / einsum.  A
quantisation decision that flips propagates through the OBS updates of its row, so integer
parity is stated as a mismatch fraction with |dq| <= 1 on the flipped entries' neighbours.
"""
import os

import numpy as np
import pytest

from oracle import aeq_oracle as O
from tests import synthetic_graph as sg

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "gptq.npz")


def _cfg(bits, gk, sym=True):
  from aeq_b200 import qtyping
  G = qtyping.QuantGranularity
  gran = {0: G.CHANNELWISE, -1: G.TENSORWISE, 32: G.BLOCKWISE_32, 64: G.BLOCKWISE_64}[gk]
  return qtyping.TensorQuantizationConfig(num_bits=bits, symmetric=sym, granularity=gran)


def _op_info(w, cfg):
  op, _ = sg.fc_graph(w)
  return sg.op_info(op, cfg)


def test_hessian_golden(cuda):
  import torch
  from aeq_b200 import device
  z = np.load(GOLD)
  for i in range(3):
    x, h = z[f"x{i}"], z[f"h{i}"]
    got = device.xtx(torch.from_numpy(x).to(cuda), 2.0 / x.shape[0])
    assert got.dtype == torch.float64
    got = got.cpu().numpy()
    np.testing.assert_allclose(got, h, rtol=0, atol=2e-6 * np.abs(np.diag(h)).max())
    np.testing.assert_array_equal(got, got.T)  # exactly symmetric


@pytest.mark.parametrize("tokens,k", [(1, 8), (33, 160), (4096, 128), (70000, 64), (513, 1000), (300, 11008 // 8)])
def test_hessian_vs_oracle(cuda, tokens, k):
  """Odd K (scalar loads), single tile with a token split, multi-tile."""
  import torch
  from aeq_b200 import device
  x = O.synthetic_activation((2, tokens, k), tokens % 97)
  want = O.gptq_hessian(x)
  got = device.xtx(torch.from_numpy(x).to(cuda), 2.0 / 2).cpu().numpy()
  np.testing.assert_allclose(got, want, rtol=0, atol=4e-6 * np.abs(np.diag(want)).max())


@pytest.fixture
def force_tensor_cores(monkeypatch):
  """Sends every eligible shape through the tcgen05 3xTF32 path (csrc/xtx_tc.cu); the default
  threshold (8 GFLOP) would keep oracle-sized cases on the SIMT kernel."""
  monkeypatch.setenv("AEQB_XTX_TC_MIN_GFLOP", "0")
  monkeypatch.delenv("AEQB_XTX_SIMT", raising=False)


@pytest.mark.parametrize("tokens,k", [
    (64, 128),       # one tile, two stages, one segment
    (1000, 256),     # ragged token count (zero-padded plane columns)
    (2048, 516),     # K % 128 != 0: TMA zero-fills the rows past K, partial float4 columns
    (777, 1028),     # three column blocks, tiles below the diagonal skipped
    (40000, 128),    # three token chunks (P accumulated across launches)
    (4096, 1536),
])
def test_hessian_tensor_core_vs_oracle(cuda, force_tensor_cores, tokens, k):
  """Same bar as the SIMT kernel: |dH| <= 4e-6 * max diag against the oracle's sgemm."""
  import torch
  from aeq_b200 import _lib, device
  x = O.synthetic_activation((2, tokens // 2, k), tokens % 89)
  x[..., 0] += 0.25  # an all-positive chain: the worst case for a truncating accumulator
  want = O.gptq_hessian(x)
  before = _lib.load().aeqb_launch_count()
  got = device.xtx(torch.from_numpy(x).to(cuda), 2.0 / 2).cpu().numpy()
  chunks = -(-tokens // 16384)
  assert _lib.load().aeqb_launch_count() - before == 2 * chunks + 2, "not on the tcgen05 path"
  np.testing.assert_allclose(got, want, rtol=0, atol=4e-6 * np.abs(np.diag(want)).max())
  np.testing.assert_array_equal(got, got.T)


def test_hessian_tensor_core_nonfinite_gate(cuda, force_tensor_cores):
  """inf / NaN / 1e38 inputs cannot be split into TF32 planes (inf * 0 = NaN): the device flag
  routes the whole product to the SIMT kernel, which propagates them like sgemm."""
  import torch
  from aeq_b200 import device
  x = O.synthetic_activation((1, 512, 256), 5)
  x[0, 3, 7] = np.inf
  x[0, 100, 9] = np.nan
  x[0, 200, 11] = 2e38
  with np.errstate(all="ignore"):
    want = O.gptq_hessian(x)
  got = device.xtx(torch.from_numpy(x).to(cuda), 2.0).cpu().numpy()
  np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
  np.testing.assert_array_equal(np.isinf(got), np.isinf(want))
  ok = np.isfinite(want)
  np.testing.assert_allclose(got[ok], want[ok], rtol=0, atol=4e-6 * np.abs(want[ok]).max())


def test_hessian_inverse_tensor_core(cuda, force_tensor_cores):
  """H^-1 = L^-T L^-1 through the tcgen05 contraction (K = T = 768)."""
  import torch
  from aeq_b200 import device
  k = 768
  x = O.synthetic_activation((4, 2 * k, k), 3)
  h = O.gptq_hessian(x)
  got = device.hessian_inverse(torch.from_numpy(h.copy()).to(cuda), 0.01).cpu().numpy().astype(np.float64)
  want = O.gptq_hessian_inverse(h.copy())
  np.testing.assert_allclose(got, want, rtol=0, atol=1e-4 * np.abs(want).max())
  d = O.gptq_damped_diagonal(h)
  hd = h.copy()
  np.fill_diagonal(hd, d)
  np.testing.assert_allclose(hd @ got, np.eye(k), rtol=0, atol=2e-3)


def test_calibrate_mirror_and_merge(cuda):
  """gptq.calibrate QSVs (gptq_test.py:50-114: +-1e39 filtered from min/max, not from H) and the
  sample-weighted Hessian merge (qsv_utils_test.py:111-180)."""
  import types
  import torch
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import gptq
  from aeq_b200.utils import qsv_utils
  x = O.synthetic_activation((3, 40, 96), 5)
  op = types.SimpleNamespace(inputs=[0, 1, -1], outputs=[2])
  tensors = [sg.tensor("in", x.shape, 0), sg.tensor("w", (8, 96), 1), sg.tensor("out", (3, 40, 8), 0)]
  graph = qtyping.GraphInfo(subgraph_tensors=tensors,
                            buffers=[types.SimpleNamespace(data=None),
                                     types.SimpleNamespace(data=np.zeros((8, 96), np.float32).tobytes())])
  y = O.synthetic_activation((3, 40, 8), 6)
  qs = gptq.calibrate(op, graph, {"in": x, "out": y})
  assert set(qs) == {"in", "out"}
  ref = O.activation_qsv(x)
  np.testing.assert_array_equal(qs["in"]["min"], ref["min"])
  np.testing.assert_array_equal(qs["in"]["max"], ref["max"])
  assert int(qs["in"]["num_samples"]) == 3
  h = O.gptq_hessian(x)
  assert qs["in"]["hessian"].dtype == np.float64
  np.testing.assert_allclose(qs["in"]["hessian"], h, rtol=0, atol=4e-6 * np.diag(h).max())
  # merge: host arrays and device tensors give the same numbers
  x2 = O.synthetic_activation((5, 40, 96), 7)
  qs2 = gptq.calibrate(op, graph, {"in": x2, "out": y})
  merged = qsv_utils.gptq_and_moving_average_update(qs["in"], qs2["in"])
  want = O.gptq_update({**ref, "hessian": qs["in"]["hessian"]},
                       {**O.activation_qsv(x2), "hessian": qs2["in"]["hessian"]})
  np.testing.assert_allclose(merged["hessian"], want["hessian"], rtol=1e-15)
  assert int(merged["num_samples"]) == 8
  np.testing.assert_array_equal(merged["min"], want["min"])
  d1 = gptq.calibrate(op, graph, {"in": x, "out": y}, keep_on_device=True)
  d2 = gptq.calibrate(op, graph, {"in": x2, "out": y}, keep_on_device=True)
  md = qsv_utils.gptq_and_moving_average_update(d1["in"], d2["in"])
  assert isinstance(md["hessian"], torch.Tensor) and md["hessian"].is_cuda
  np.testing.assert_array_equal(md["hessian"].cpu().numpy(), merged["hessian"])


def test_hessian_inverse_golden(cuda):
  import torch
  from aeq_b200 import device
  z = np.load(GOLD)
  for i in range(3):
    h, want = z[f"h{i}"], z[f"hinv{i}"]
    hd = torch.from_numpy(h.copy()).to(cuda)
    got = device.hessian_inverse(hd, 0.01, keep_damped_diagonal=True)
    assert got.dtype == torch.float32
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=0, atol=2e-5 * np.abs(want).max())
    # the reference leaves the damped diagonal in the caller's matrix
    np.testing.assert_allclose(torch.diagonal(hd).cpu().numpy(), z[f"hdiag_after{i}"], rtol=1e-15)
    hd2 = torch.from_numpy(h.copy()).to(cuda)
    device.hessian_inverse(hd2, 0.01, keep_damped_diagonal=False)
    np.testing.assert_array_equal(hd2.cpu().numpy(), h)


@pytest.mark.parametrize("k", [3, 31, 64, 65, 127, 128, 129, 200, 257, 383, 1024, 1161])
def test_hessian_inverse_property(cuda, k):
  """H_damped @ Hinv == I for block-edge orders (32 / 64 / 128 block boundaries, padding, odd K:
  the 8-byte cp.async path of chol_dmma.cu)."""
  import torch
  from aeq_b200 import device
  x = O.synthetic_activation((4, max(2 * k, 64), k), k)
  h = O.gptq_hessian(x)
  if k == 31:  # a dead input feature: zero row / column, diagonal 0 -> replaced by 1
    h[0, :] = 0.0
    h[:, 0] = 0.0
  hd = torch.from_numpy(h.copy()).to(cuda)
  got = device.hessian_inverse(hd, 0.01, keep_damped_diagonal=True).cpu().numpy().astype(np.float64)
  damped = hd.cpu().numpy()
  want = O.gptq_hessian_inverse(h.copy())
  np.testing.assert_allclose(got, want, rtol=0, atol=1e-4 * np.abs(want).max())
  np.testing.assert_allclose(damped @ got, np.eye(k), rtol=0, atol=2e-3)


@pytest.mark.parametrize("k", [31, 64, 200, 1024])
@pytest.mark.parametrize("min_k", ["0", "1000000", "dmma", "dmma-serial"])
def test_hessian_inverse_both_cholesky_variants(cuda, monkeypatch, k, min_k):
  """Every Cholesky variant meets the same bars: the DMMA / lookahead one (chol_dmma.cu, the
  default; also with the lookahead stream switched off), the two-level SIMT one (rank-128 trailing
  updates, left-looking panels) and the single-level one (AEQB_CHOL_TWO_LEVEL_MIN_K above K)."""
  import torch
  from aeq_b200 import device
  if min_k.startswith("dmma"):
    monkeypatch.setenv("AEQB_CHOL_DMMA_MIN_K", "0")
    if min_k == "dmma-serial":
      monkeypatch.setenv("AEQB_CHOL_NO_LOOKAHEAD", "1")
  else:
    monkeypatch.setenv("AEQB_CHOL_DMMA_MIN_K", "1000000")
    monkeypatch.setenv("AEQB_CHOL_TWO_LEVEL_MIN_K", min_k)
  x = O.synthetic_activation((4, max(2 * k, 64), k), k + 1)
  h = O.gptq_hessian(x)
  if k == 31:
    h[0, :] = 0.0
    h[:, 0] = 0.0
  got = device.hessian_inverse(torch.from_numpy(h.copy()).to(cuda), 0.01).cpu().numpy().astype(np.float64)
  want = O.gptq_hessian_inverse(h.copy())
  np.testing.assert_allclose(got, want, rtol=0, atol=1e-4 * np.abs(want).max())
  hd = h.copy()
  np.fill_diagonal(hd, O.gptq_damped_diagonal(h))
  np.testing.assert_allclose(hd @ got, np.eye(k), rtol=0, atol=2e-3)
  bad = np.eye(k)
  bad[k // 2, k // 2] = -5.0
  with pytest.raises(np.linalg.LinAlgError):
    device.hessian_inverse(torch.from_numpy(bad).to(cuda), 0.0)


def test_hessian_inverse_not_positive_definite(cuda):
  import torch
  from aeq_b200 import device
  h = -np.eye(8)
  with pytest.raises(np.linalg.LinAlgError):
    device.hessian_inverse(torch.from_numpy(h).to(cuda))


def test_column_loop_same_hinv_vs_oracle(cuda):
  """Same Hinv on both sides isolates the OBS loop: one 64-column block is operation-for-operation
  the reference (bit-exact); several blocks differ only by the inter-block dot-product order."""
  import torch
  from aeq_b200 import device
  z = np.load(GOLD)
  for i, sym, bits, gk in ((0, True, 4, 0), (1, True, 4, 0), (1, True, 8, 32), (2, False, 4, 0), (2, True, 4, 32)):
    w, hinv = z[f"w{i}"], np.ascontiguousarray(z[f"hinv{i}"])  # strtri hands back Fortran order
    mn, mx = O.weight_minmax(w, max(gk, 0), True)
    zp, scale = O.scale_zp(mn, mx, bits, sym, gk > 0)
    with np.errstate(all="ignore"):
      want = O.gptq_quantize(w, scale, zp, hinv, bits, sym, max(gk, 0))
    got = device.gptq_quantize(torch.from_numpy(w).to(cuda), torch.from_numpy(hinv).to(cuda),
                               torch.from_numpy(scale.reshape(-1)).to(cuda),
                               torch.from_numpy(zp.astype(np.int32).reshape(-1)).to(cuda),
                               max(gk, 0), bits, sym).cpu().numpy()
    d = np.abs(got.astype(int) - want.astype(int))
    assert (d > 0).mean() <= 5e-3, (i, sym, bits, gk, (d > 0).mean())
    np.testing.assert_array_equal(got[:, :64], want[:, :64])  # first block: no inter-block term yet
  # K <= 64: a single block, bit-exact end to end given the same Hinv
  w = O.synthetic_weight(40, 48, 3)
  x = O.synthetic_activation((2, 100, 48), 3)
  hinv = O.gptq_hessian_inverse(O.gptq_hessian(x))
  for sym in (True, False):
    mn, mx = O.weight_minmax(w, 0, True)
    zp, scale = O.scale_zp(mn, mx, 4, sym, False)
    want = O.gptq_quantize(w, scale, zp, hinv, 4, sym)
    got = device.gptq_quantize(torch.from_numpy(w).to(cuda), torch.from_numpy(hinv).to(cuda),
                               torch.from_numpy(scale.reshape(-1)).to(cuda),
                               torch.from_numpy(zp.astype(np.int32).reshape(-1)).to(cuda), 0, 4, sym)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_column_loop_divide_window_fallback(cuda):
  """The column kernel divides through hoisted reciprocals only inside the window where that is
  provably the IEEE quotient; rows outside it (an all-zero row: scale 1e-9 / 7 is below 2^-30, a
  row of 1e30s: errors above 2^90) are redone with IEEE divides.  Single block: bit-exact."""
  import torch
  from aeq_b200 import device
  w = O.synthetic_weight(24, 64, 11)
  w[1, :] = 0.0
  w[2, :] = 1e30
  w[3, :8] = 1e-38
  x = O.synthetic_activation((2, 200, 64), 11)
  hinv = O.gptq_hessian_inverse(O.gptq_hessian(x))
  for sym in (True, False):
    mn, mx = O.weight_minmax(w, 0, True)
    zp, scale = O.scale_zp(mn, mx, 4, sym, False)
    with np.errstate(all="ignore"):
      want = O.gptq_quantize(w, scale, zp, hinv, 4, sym)
    got = device.gptq_quantize(torch.from_numpy(w).to(cuda), torch.from_numpy(hinv).to(cuda),
                               torch.from_numpy(scale.reshape(-1)).to(cuda),
                               torch.from_numpy(zp.astype(np.int32).reshape(-1)).to(cuda), 0, 4, sym)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_gptq_golden_end_to_end(cuda):
  from aeq_b200.algorithms.uniform_quantize import gptq
  z = np.load(GOLD)
  for key in [str(c) for c in z["cases"]]:
    wname, b, g = key.split("_")
    i, bits, gk = int(wname[1:]), int(b[1:]), int(g[1:])
    w, h, x = z[wname], z[f"h{i}"], z[f"x{i}"]
    cfg = _cfg(bits, gk)
    qsv = {"activation_tensor_qsv": {"hessian": h.copy(), "num_samples": x.shape[0]}}
    r = gptq.get_tensor_quant_params(_op_info(w, cfg), cfg, w, qsv)
    np.testing.assert_array_equal(r.scale, z[key + "_scale"])  # min-max scales: bit-exact
    assert r.quantized_data.dtype == np.int8
    d = np.abs(r.quantized_data.astype(int) - z[key + "_q"].astype(int))
    assert (d > 0).mean() <= 1e-2, (key, (d > 0).mean())
    assert d.max() <= 2, (key, d.max())
    # the caller's Hessian keeps the damped diagonal, like the reference
    np.testing.assert_allclose(np.diag(qsv["activation_tensor_qsv"]["hessian"]), z[f"hdiag_after{i}"],
                               rtol=1e-15)


def test_reference_literals(cuda):
  """gptq_test.py:214-256, :258-299."""
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import gptq
  cfg = _cfg(8, -1)
  w = np.array([[1.1, 2.1914, 0.6], [-1.1, 0.1, -0.6]], dtype=np.float32)
  hess = np.array([[15.0, 0.5, 0.1], [0.5, 1.0, 0.2], [0.1, 0.2, 1.0]], dtype=np.float32)
  info = qtyping.OpInfo(op=sg.fc_graph(w)[0], op_name=qtyping.TFLOperationName.FULLY_CONNECTED,
                        subgraph_op_index=-1, op_quant_config=qtyping.OpQuantizationConfig())
  qsv = {"min": np.array([[-1.1]]), "max": np.array([[2.2]]),
         "activation_tensor_qsv": {"hessian": hess.copy(), "num_samples": 1}}
  r = gptq.get_tensor_quant_params(info, cfg, w, qsv)
  np.testing.assert_allclose(r.scale, np.array([[2.2 / 127]]), rtol=1e-6)
  np.testing.assert_array_equal(r.zero_point, np.array([[0]]))
  np.testing.assert_array_equal(r.quantized_data, np.array([[64, 126, 35], [-64, 6, -35]], np.int8))
  info2 = qtyping.OpInfo(op=sg.fc_graph(w)[0], op_name=qtyping.TFLOperationName.FULLY_CONNECTED,
                         subgraph_op_index=-1,
                         op_quant_config=qtyping.OpQuantizationConfig(weight_tensor_config=cfg))
  r = gptq.get_tensor_quant_params(
      info2, cfg, w, {"activation_tensor_qsv": {"hessian": hess.copy(), "num_samples": 1}})
  np.testing.assert_allclose(r.scale, np.array([[2.1914 / 127]]), rtol=1e-6)
  np.testing.assert_array_equal(r.quantized_data, np.array([[64, 127, 35], [-64, 6, -35]], np.int8))
  # no Hessian: parameters only, no quantized data (gptq.py:290-294)
  r = gptq.get_tensor_quant_params(info2, cfg, w, None)
  assert r.quantized_data is None
  with pytest.raises(ValueError, match="not found in tensor_name_to_qsv"):
    gptq.get_tensor_quant_params(info2, cfg, None, None)


def test_gptq_larger_proxy_loss(cuda):
  """[256, 512] weight: integers close to the oracle's, and the proxy loss tr(E H E^T) no worse
  than the oracle's by more than 0.1 % (and far better than plain rounding)."""
  from aeq_b200.algorithms.uniform_quantize import gptq
  w = O.synthetic_weight(256, 512, 77)
  x = O.synthetic_activation((4, 512, 512), 77)
  x *= (1.0 + np.arange(512, dtype=np.float32) % 7)  # anisotropic activations
  h = O.gptq_hessian(x)
  cfg = _cfg(4, 0)
  r = gptq.get_tensor_quant_params(_op_info(w, cfg), cfg, w,
                                   {"activation_tensor_qsv": {"hessian": h.copy(), "num_samples": 4}})
  ref = O.gptq_requant(w, h.copy(), 4)
  plain = O.minmax_requant(w, 4)

  def loss(q):
    e = (w - q.astype(np.float32) * ref["scale"]).astype(np.float64)
    return float(np.einsum("ri,ij,rj->", e, h, e))

  d = np.abs(r.quantized_data.astype(int) - ref["q"].astype(int))
  assert (d > 0).mean() <= 2e-2, (d > 0).mean()
  assert loss(r.quantized_data) <= loss(ref["q"]) * 1.001
  assert loss(r.quantized_data) < 0.97 * loss(plain["q"])


def test_hadamard_rotated_hessian(cuda):
  """R^T H R on the device (two passes of the rotation kernel around a transpose) against the
  oracle's float64 einsum, and against the definition: the Hessian of the rotated activations."""
  import torch
  from aeq_b200.algorithms.uniform_quantize import hadamard_gptq
  for k, n in ((256, 256), (768, 256), (512, 64), (96, 32), (48, 16)):
    x = O.synthetic_activation((3, 200, k), k)
    x *= (1.0 + np.arange(k, dtype=np.float32) % 5)
    h = O.gptq_hessian(x)
    want = O.hadamard_rotate_hessian(h, n)
    got = hadamard_gptq.rotate_hessian_device(torch.from_numpy(h).to(cuda), n).cpu().numpy()
    assert got.dtype == np.float64
    np.testing.assert_allclose(got, want, rtol=0, atol=4e-6 * np.abs(np.diag(want)).max())
    xr, n2 = O.hadamard_rotate(x, n)
    assert n2 == n
    np.testing.assert_allclose(O.gptq_hessian(xr), want, rtol=0, atol=2e-5 * np.abs(np.diag(want)).max())


@pytest.mark.parametrize("rows,k,max_size", [(64, 256, None), (96, 768, None), (40, 512, 64)])
def test_hadamard_gptq_matches_composed_oracle(cuda, rows, k, max_size):
  """BASELINE.json configs[4] as one algorithm: rotated weight, rotated Hessian, GPTQ.  Scales are
  min/max of the rotated weight (<= 1e-6 rel.: the rotation's summation order), integers within
  GPTQ's bars (a flipped rounding decision propagates along its row), proxy loss within 0.1 %."""
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import hadamard_gptq
  w = O.synthetic_weight(rows, k, rows + k)
  x = O.synthetic_activation((4, 2 * k, k), k)
  x = x + 0.5 * np.roll(x, 1, axis=-1)  # correlated input features: GPTQ has something to do
  h = O.gptq_hessian(x)
  cfg = qtyping.TensorQuantizationConfig(4, True, qtyping.QuantGranularity.CHANNELWISE,
                                         algorithm_params={"max_hadamard_size": max_size} if max_size else {})
  r = hadamard_gptq.get_tensor_quant_params(_op_info(w, cfg), cfg, w,
                                            {"activation_tensor_qsv": {"hessian": h.copy(), "num_samples": 4}})
  ref = O.hadamard_gptq_requant(w, h.copy(), 4, max_size)
  assert r.hadamard.hadamard_size == ref["hadamard_size"]
  np.testing.assert_array_equal(r.hadamard.random_binary_vector, ref["random_binary_vector"])
  assert r.quantized_dimension == 0 and r.quantized_data.dtype == np.int8 and r.scale.shape == (rows, 1)
  np.testing.assert_allclose(r.scale, ref["scale"], rtol=2e-6)
  d = np.abs(r.quantized_data.astype(int) - ref["q"].astype(int))
  assert (d > 0).mean() <= 2e-2 and d.max() <= 2, ((d > 0).mean(), d.max())

  def loss(q, scale):
    e = (ref["rotated"] - q.astype(np.float32) * scale).astype(np.float64)
    return float(np.einsum("ri,ij,rj->", e, ref["hessian_rotated"], e))

  assert abs(loss(r.quantized_data, r.scale) - loss(ref["q"], ref["scale"])) <= 1e-3 * loss(ref["q"], ref["scale"])
  plain = O.minmax_requant(ref["rotated"], 4)
  assert loss(r.quantized_data, r.scale) < loss(plain["q"], plain["scale"])
  # without a Hessian the weight keeps GPTQ's behaviour: parameters only
  r0 = hadamard_gptq.get_tensor_quant_params(_op_info(w, cfg), cfg, w, None)
  assert r0.quantized_data is None


@pytest.mark.parametrize("rows,k,bits,sym,gk", [(256, 1024, 4, True, 0), (100, 1152, 4, False, 0),
                                                (160, 1024, 8, True, 32), (33, 2048, 4, True, 0)])
def test_left_looking_tensor_core_path_vs_oracle(cuda, rows, k, bits, sym, gk):
  """K >= 1024 (a multiple of 64) takes the left-looking schedule: lane-per-row column kernel +
  one 3xTF32 tcgen05 product per block over everything quantised so far.  Same H^-1 on both sides:
  the first block has no inter-block term (bit-exact), the rest differs from the reference's
  right-looking loop only by where the update's dot products are rounded."""
  import torch
  from aeq_b200 import _lib, device
  w = O.synthetic_weight(rows, k, rows + k)
  x = O.synthetic_activation((4, 2 * k, k), k + 1)
  x = x + 0.5 * np.roll(x, 1, axis=-1)
  hinv = np.ascontiguousarray(O.gptq_hessian_inverse(O.gptq_hessian(x)))
  mn, mx = O.weight_minmax(w, max(gk, 0), True)
  zp, scale = O.scale_zp(mn, mx, bits, sym, gk > 0)
  with np.errstate(all="ignore"):
    want = O.gptq_quantize(w, scale, zp, hinv, bits, sym, max(gk, 0))
  before = _lib.load().aeqb_launch_count()
  got = device.gptq_quantize(torch.from_numpy(w).to(cuda), torch.from_numpy(hinv).to(cuda),
                             torch.from_numpy(scale.reshape(-1)).to(cuda),
                             torch.from_numpy(zp.astype(np.int32).reshape(-1)).to(cuda),
                             max(gk, 0), bits, sym).cpu().numpy()
  nblocks = k // 64
  # plane split + one column kernel per block + the tensor-core product of every block from the third on
  # (lookahead: a block's predecessor contributes through the column kernel itself)
  assert _lib.load().aeqb_launch_count() - before == 1 + nblocks + (nblocks - 2), "not on the left-looking path"
  np.testing.assert_array_equal(got[:, :64], want[:, :64])
  d = np.abs(got.astype(int) - want.astype(int))
  assert (d > 0).mean() <= 5e-3 and d.max() <= 2, ((d > 0).mean(), d.max())

  def loss(qq):
    s = scale if gk == 0 else np.repeat(scale, gk, axis=1)
    z = zp if gk == 0 else np.repeat(zp, gk, axis=1)
    e = (w - (qq.astype(np.float32) - z) * s).astype(np.float64)
    h = np.linalg.inv(hinv.astype(np.float64))
    return float(np.einsum("ri,ij,rj->", e, h, e))

  assert abs(loss(got) - loss(want)) <= 1e-3 * loss(want)


@pytest.mark.parametrize("rows,k", [(40, 200), (70, 512)])
def test_lane_per_row_column_kernel_equals_the_row_per_warp_one(cuda, monkeypatch, rows, k):
  """Shapes below the tensor-core threshold keep the reference's right-looking order; the new
  lane-per-row column kernel and the round-1 row-per-warp kernel (AEQB_GPTQ_OLD_COLS=1, read once
  per process, so compared against the oracle instead) give the reference's integers given the
  same H^-1: single-block shapes bit-exact, multi-block within the update's rounding order."""
  import torch
  from aeq_b200 import device
  w = O.synthetic_weight(rows, k, 7)
  w[1, :] = 0.0      # scale 1e-9 / 7: outside the exact-divide window -> IEEE redo of the warp
  w[2, :] = 1e30
  x = O.synthetic_activation((2, 2 * k, k), 9)
  hinv = np.ascontiguousarray(O.gptq_hessian_inverse(O.gptq_hessian(x)))
  for sym in (True, False):
    mn, mx = O.weight_minmax(w, 0, True)
    zp, scale = O.scale_zp(mn, mx, 4, sym, False)
    with np.errstate(all="ignore"):
      want = O.gptq_quantize(w, scale, zp, hinv, 4, sym)
    got = device.gptq_quantize(torch.from_numpy(w).to(cuda), torch.from_numpy(hinv).to(cuda),
                               torch.from_numpy(scale.reshape(-1)).to(cuda),
                               torch.from_numpy(zp.astype(np.int32).reshape(-1)).to(cuda), 0, 4, sym).cpu().numpy()
    np.testing.assert_array_equal(got[:, :64], want[:, :64])
    assert (got != want).mean() <= 5e-3


def test_layer_level_concurrent_gptq_equals_one_weight_at_a_time(cuda):
  """hadamard_gptq.quantize_layer_device (every Hessian inverse and OBS loop on its own stream, the
  lookahead Cholesky keeping one side stream per caller stream) returns exactly what
  quantize_device returns for each weight alone: same kernels, same launches per problem."""
  import torch
  from aeq_b200.algorithms.uniform_quantize import hadamard_gptq
  shapes = [(96, 512), (64, 512), (160, 512), (64, 384), (200, 384)]
  feeds = ["a", "a", "a", "b", "b"]
  ws = [torch.from_numpy(O.synthetic_weight(r, k, 40 + i)).to(cuda) for i, (r, k) in enumerate(shapes)]
  hs = {"a": torch.from_numpy(O.gptq_hessian(O.synthetic_activation((4, 300, 512), 7))).to(cuda),
        "b": torch.from_numpy(O.gptq_hessian(O.synthetic_activation((4, 300, 384), 8))).to(cuda)}
  for concurrent in (True, False):
    got = hadamard_gptq.quantize_layer_device(ws, feeds, hs, 4, True, 128, concurrent=concurrent)
    torch.cuda.synchronize()
    for w, f, (q, scale, zp, n) in zip(ws, feeds, got):
      wq, wscale, wzp, wn = hadamard_gptq.quantize_device(w, hs[f], 4, True, 128)
      assert n == wn == 128
      assert torch.equal(q, wq) and torch.equal(scale, wscale) and torch.equal(zp, wzp)
  bad = {"a": -torch.eye(512, dtype=torch.float64, device=cuda), "b": hs["b"]}
  with pytest.raises(np.linalg.LinAlgError):
    hadamard_gptq.quantize_layer_device(ws, feeds, bad, 4, True, 128)
