"""GPU parity of OCTAV, MSE and Hadamard rotation against the reference-generated
golden fixtures (tests/golden/*.npz) and against the oracle on larger seeded inputs.

Tolerances: these algorithms contain order-dependent fp32 sums, so
 * clipping constants / per-channel scales: <= 1e-6 relative,
 * blockwise scales are rounded fp32 -> bf16 -> fp16, so a 1e-7 difference before the
   rounding may move a scale by one bf16 ulp (2^-8 relative): <= 0.2 % of them may do so,
 * integers: |dq| <= 1 and a mismatch fraction <= 1e-3 (exactly 0 wherever the scale is
   bit-identical and the path is the fused kernel),
 * rotated weights: |d| <= 2e-6 * max|rotated| (sgemm vs butterfly summation order).
"""
import numpy as np
import pytest

from oracle import aeq_oracle as O
from tests import synthetic_graph as sg

pytestmark = pytest.mark.gpu

GOLD = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")


def _cfg(bits, gk, **params):
  from aeq_b200 import qtyping
  G = qtyping.QuantGranularity
  gran = {0: G.CHANNELWISE, -1: G.TENSORWISE, 32: G.BLOCKWISE_32, 64: G.BLOCKWISE_64,
          128: G.BLOCKWISE_128, 256: G.BLOCKWISE_256}[gk]
  return qtyping.TensorQuantizationConfig(num_bits=bits, symmetric=True, granularity=gran,
                                          algorithm_params=params)


def _run(mod, w, cfg):
  op, _ = sg.fc_graph(w)
  return mod.get_tensor_quant_params(sg.op_info(op, cfg), cfg, w, None)


def _assert_scales(got, want, blockwise):
  got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
  assert got.shape == want.shape
  if not blockwise:
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=0)
    return
  rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
  assert rel.max() <= 2.0 ** -7, rel.max()           # at most one bf16 ulp
  assert (rel > 0).mean() <= 2e-3, (rel > 0).mean()  # and only rarely


def _assert_ints(got, want, frac=1e-3):
  d = np.abs(got.astype(np.int32) - want.astype(np.int32))
  assert d.max() <= 1, d.max()
  assert (d > 0).mean() <= frac, (d > 0).mean()


def _cases(name):
  z = np.load(f"{GOLD}/{name}.npz")
  return z, [str(c) for c in z["cases"]]


def test_octav_golden(cuda):
  from aeq_b200.algorithms.uniform_quantize import octav
  z, cases = _cases("octav")
  for key in cases:
    wname, b, g = key.split("_")
    bits, gk = int(b[1:]), int(g[1:])
    w = z[wname]
    r = _run(octav, w, _cfg(bits, gk))
    _assert_scales(r.scale, z[key + "_scale"], gk > 0)
    _assert_ints(r.quantized_data, z[key + "_q"])
    assert r.quantized_data.dtype == np.int8 and r.zero_point.dtype == np.int8
    assert not r.zero_point.any()
    # the clipping constants themselves, through the NumPy-facing mirror
    if gk > 0:
      clip = octav.guess_clipping_with_octav(w.reshape(w.shape[0], -1, gk), bits, 2)
    elif gk == 0:
      clip = octav.guess_clipping_with_octav(w, bits, (1,))
    else:
      clip = octav.guess_clipping_with_octav(w, bits, (0, 1))
    np.testing.assert_allclose(clip, z[key + "_clip"], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("shape", [(64, 4096), (24, 11008), (5, 16384), (9, 260), (3, 33), (40, 1024),
                                   (33, 2048), (17, 3072), (11, 3584), (13, 1536), (19, 640), (1, 4096)])
@pytest.mark.parametrize("bits", [4, 8])
def test_octav_rows_vs_oracle(cuda, shape, bits):
  """Row-resident kernel (every NV class), the generic kernel (odd shapes) and the trace /
  early-stop selection against the oracle's iteration-by-iteration trace."""
  from aeq_b200 import device
  import torch
  w = O.synthetic_weight(*shape, index=17 + bits)
  w[0, :] = 0.0                      # converges to 0: zeros are counted twice from iteration 2 on
  if shape[0] > 2:
    w[1, :] = 2.5                    # everything above the initial guess
    w[2, ::3] = 0.0
  x = torch.from_numpy(w).to(cuda)
  with np.errstate(all="ignore"):
    want, trace = O.octav_clip(w, bits, (1,), return_trace=True)
  got = device.octav_clip_rows(x, bits).cpu().numpy()
  np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-12)
  # without early stop the constant is the 10th iterate
  with np.errstate(all="ignore"):
    want10 = O.octav_clip(w, bits, (1,), early_stop=False)
  got10 = device.octav_clip_rows(x, bits, early_stop=False).cpu().numpy()
  np.testing.assert_allclose(got10, want10, rtol=1e-6, atol=1e-12)
  assert len(trace) <= 10


@pytest.mark.parametrize("shape", [(3700, 2048), (2000, 4096), (1700, 4096), (4300, 4096)])
def test_octav_rows_warp_pipeline(cuda, shape):
  """More rows than resident warps: every warp of the one-warp-per-row kernel refills its
  shared-memory row buffer (bulk copy + mbarrier phase flip) at least once; rows with NaN / inf
  / zero elements sit in both rounds.  [4300, 4096] is three rounds of the 1776 resident warps of a
  B200; special values sit at both ends of the rows."""
  from aeq_b200 import device
  import torch
  w = O.synthetic_weight(*shape, index=91)
  for r in (0, shape[0] - 1):
    w[r, 5] = np.nan
    w[r, 77] = 0.0
  w[1, :] = 0.0
  w[shape[0] - 2, ::2] = 0.0
  w[shape[0] - 3, 9] = np.inf
  if shape[1] > 3000:
    w[7, 3000] = np.nan
    w[shape[0] - 5, 4095] = np.nan
    w[8, 2560:] = 0.0
    w[shape[0] - 6, 2600] = np.inf
    w[9, 2560:] = 3.0
    w[shape[0] - 7, :2560] = 0.0
  with np.errstate(all="ignore"):
    want = O.octav_clip(w, 4, (1,))
  got = device.octav_clip_rows(torch.from_numpy(w).to(cuda), 4).cpu().numpy()
  np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize("shape,block", [((64, 4096), 32), ((7, 11008), 32), ((16, 1024), 64),
                                         ((16, 1024), 128), ((9, 2048), 256), ((3, 32), 32)])
@pytest.mark.parametrize("bits", [4, 8])
def test_octav_blocks_vs_oracle(cuda, shape, block, bits):
  from aeq_b200 import device
  import torch
  w = O.synthetic_weight(*shape, index=23 + bits)
  w[0, :block] = 0.0
  x = torch.from_numpy(w).to(cuda)
  r, c = shape
  with np.errstate(all="ignore"):
    want = O.octav_clip(w.reshape(r, c // block, block), bits, 2).reshape(r, c // block)
  got = device.octav_clip_blocks(x, block, bits).cpu().numpy()
  np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-12)


def test_octav_full_requant_vs_oracle(cuda):
  from aeq_b200.algorithms.uniform_quantize import octav
  w = O.synthetic_weight(96, 4096, 31)
  for bits, gk in ((4, 0), (8, 0), (4, 32), (4, -1)):
    with np.errstate(all="ignore"):
      ref = O.octav_requant(w, bits, block=max(gk, 0), per_channel=(gk == 0))
    r = _run(octav, w, _cfg(bits, gk))
    _assert_scales(r.scale, ref["scale"], gk > 0)
    _assert_ints(r.quantized_data, ref["q"])
    same = np.broadcast_to(r.scale == ref["scale"], r.scale.shape)
    if gk == 0:  # rows whose scale is bit-identical must have bit-identical integers
      np.testing.assert_array_equal(r.quantized_data[same[:, 0]], ref["q"][same[:, 0]])


def test_octav_errors(cuda):
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import octav
  w = O.synthetic_weight(4, 64, 1)
  cfg = qtyping.TensorQuantizationConfig(num_bits=4, symmetric=False,
                                         granularity=qtyping.QuantGranularity.CHANNELWISE)
  op, _ = sg.fc_graph(w)
  with pytest.raises(ValueError, match="OCTAV supports symmetric quantization only"):
    octav.get_tensor_quant_params(sg.op_info(op, cfg), cfg, w, None)
  with pytest.raises(ValueError, match="not divisible by block size"):
    _run(octav, O.synthetic_weight(4, 48, 1), _cfg(4, 32))


def test_mse_golden(cuda):
  from aeq_b200.algorithms.uniform_quantize import mse
  z, cases = _cases("mse")
  for key in cases:
    wname, b, g = key.split("_")
    bits, gk = int(b[1:]), int(g[1:])
    r = _run(mse, z[wname], _cfg(bits, gk))
    _assert_scales(r.scale, z[key + "_scale"], False)
    _assert_ints(r.quantized_data, z[key + "_q"])
    assert r.zero_point.dtype == np.int32 and r.zero_point.shape == r.scale.shape


def test_mse_vs_oracle_and_errors(cuda):
  from aeq_b200.algorithms.uniform_quantize import mse
  for shape in ((64, 4096), (5, 11008), (7, 33), (2, 70000)):
    w = O.synthetic_weight(*shape, index=41)
    for bits in (4, 8):
      ref = O.mse_requant(w, bits)
      r = _run(mse, w, _cfg(bits, 0))
      _assert_scales(r.scale, ref["scale"], False)
      _assert_ints(r.quantized_data, ref["q"])
  with pytest.raises(ValueError, match="Blockwise quantization is not supported for MSE"):
    _run(mse, O.synthetic_weight(4, 64, 1), _cfg(4, 32))
  with pytest.raises(KeyError):
    _run(mse, O.synthetic_weight(4, 64, 1), _cfg(2, 0))


@pytest.mark.parametrize("shape", [(96, 4096), (9, 11008), (40, 1536), (3, 16384), (130, 256)])
def test_mse_fused_equals_unfused(cuda, shape):
  """aeqb_requant_mse_rows_f32 (one pass) against aeqb_mse_scale_rows_f32 + aeqb_quantize_f32 and the
  oracle: same fp64 sum of fp32 squares up to association, so scales are equal to 1 ulp at worst
  and bit-equal here; degenerate rows (all zero -> scale 0 -> q 0, NaN -> q 0) follow the reference."""
  import torch
  from aeq_b200 import device
  w = O.synthetic_weight(*shape, index=shape[1] % 23)
  w[1, :] = 0.0
  if shape[0] > 4:
    w[3, 5] = np.nan
  x = torch.from_numpy(w).to(cuda)
  for bits, k in ((8, 0.05408), (4, 0.37755)):
    fused = device.requant_mse_rows(x, bits, k, want_packed=(bits == 4))
    scale = device.mse_scale_rows(x, k)
    q = device.quantize(x, scale.reshape(-1), None, bits, True, shape[0], shape[1])
    np.testing.assert_array_equal(fused.scale.cpu().numpy().ravel(), scale.cpu().numpy().ravel())
    np.testing.assert_array_equal(fused.q.cpu().numpy(), q.cpu().numpy())
    assert not fused.zero_point.any()
    if bits == 4:
      np.testing.assert_array_equal(fused.packed.cpu().numpy(), O.pack_bits(4, fused.q.cpu().numpy()))
    with np.errstate(all="ignore"):
      ref = O.mse_requant(w, bits)
    ok = np.isfinite(ref["scale"]).ravel()
    np.testing.assert_allclose(fused.scale.cpu().numpy().ravel()[ok], ref["scale"].ravel()[ok], rtol=1e-6)
    _assert_ints(fused.q.cpu().numpy(), ref["q"])
    assert not fused.q.cpu().numpy()[1].any()


def test_hadamard_golden(cuda):
  from aeq_b200.algorithms.uniform_quantize import hadamard_rotation as had
  z, cases = _cases("hadamard")
  for i in range(4):
    w, want = z[f"w{i}"], z[f"w{i}_rot"]
    cap = int(z[f"w{i}_cap"])
    rot, n, vec = had._rotate_with_diagonal_hadamard(w, 1, None if cap < 0 else cap)
    assert n == int(z[f"w{i}_b4_hsize"])
    assert vec.dtype == np.int8 and vec.shape == (n,) and vec.all()
    np.testing.assert_allclose(rot, want, rtol=0, atol=2e-6 * np.abs(want).max())
  for key in cases:
    wname, b = key.split("_")
    i, bits = int(wname[1:]), int(b[1:])
    cap = int(z[f"w{i}_cap"])
    params = {} if cap < 0 else {"max_hadamard_size": cap}
    r = _run(had, z[wname], _cfg(bits, 0, **params))
    assert r.hadamard.hadamard_size == int(z[key + "_hsize"])
    np.testing.assert_allclose(r.scale, z[key + "_scale"], rtol=2e-6)
    # a rotated value moved by 1e-7 relative can cross a rounding boundary
    _assert_ints(r.quantized_data, z[key + "_q"], frac=2e-3)


def test_hadamard_reference_literals(cuda):
  """hadamard_rotation_test.py:274-351: the three integer goldens."""
  from aeq_b200.algorithms.uniform_quantize import hadamard_rotation as had
  r = _run(had, np.ones((6, 6), np.float32), _cfg(8, 0))
  np.testing.assert_array_equal(r.quantized_data, np.tile([127, 0], (6, 3)))
  assert r.hadamard.hadamard_size == 2
  w = np.tile(np.array([[1, 2], [3, 4]], np.float32), (3, 3))
  r = _run(had, w, _cfg(8, 0))
  np.testing.assert_array_equal(r.quantized_data, np.tile([[127, -42], [127, -18]], (3, 3)))


@pytest.mark.parametrize("shape,n", [((64, 4096), 4096), ((16, 4096), 128), ((8, 11008), 256),
                                     ((4, 16384), 16384), ((3, 8192), 8192), ((5, 2048), 2),
                                     ((2, 32768), 32768), ((6, 24), 8),
                                     # 16 KiB tile kernel (numel % 4096 == 0, n = 256..4096), the
                                     # last one with more tiles than resident warps
                                     ((16, 11008), 256), ((24, 1024), 512), ((8, 4096), 1024),
                                     ((4, 3072), 1024), ((32, 2048), 2048), ((1200, 4096), 4096)])
def test_hadamard_rows_vs_oracle(cuda, shape, n):
  """Every radix path (even / odd log2 n), tiled segments, the 16 KiB tile kernel and the
  > 64 KiB global fallback."""
  from aeq_b200 import device
  import torch
  w = O.synthetic_weight(*shape, index=n % 89)
  want = np.matmul(w.reshape(-1, n), O.hadamard_matrix(n)).reshape(shape)
  got = device.hadamard_rows(torch.from_numpy(w).to(cuda), n).cpu().numpy()
  np.testing.assert_allclose(got, want, rtol=0, atol=2e-6 * np.abs(want).max())
  # orthogonality: rotating twice returns the input (H/sqrt(n) is an involution)
  back = device.hadamard_rows(torch.from_numpy(got).to(cuda), n).cpu().numpy()
  np.testing.assert_allclose(back, w, rtol=0, atol=4e-6 * np.abs(w).max())


def test_registry_reaches_new_algorithms(cuda):
  """OCTAV / MSE / HADAMARD_ROTATION through algorithm_manager like params_generator does."""
  from aeq_b200 import algorithm_manager as am
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.utils import common_utils
  w = O.synthetic_weight(32, 256, 3)
  for key, ref in (("OCTAV", O.octav_requant(w, 4)), ("MSE", O.mse_requant(w, 4)),
                   ("HADAMARD_ROTATION", O.hadamard_requant(w, 4))):
    cfg = _cfg(4, 0)
    op, graph = sg.fc_graph(w)
    info = qtyping.OpInfo(op, qtyping.TFLOperationName.FULLY_CONNECTED, 0,
                          qtyping.OpQuantizationConfig(
                              weight_tensor_config=cfg,
                              compute_precision=qtyping.ComputePrecision.INTEGER))
    fn = am.get_quantization_func(key, qtyping.TFLOperationName.FULLY_CONNECTED,
                                  qtyping.QuantizeMode.MATERIALIZE)
    out = fn(info, graph, {}, common_utils.TensorQuantParamsCache())
    params = [t for t in out if t.tensor_name == "weight"][0].consumers[0].parameters
    np.testing.assert_allclose(params.scale, ref["scale"], rtol=2e-6)
    _assert_ints(params.quantized_data, ref["q"], frac=2e-3)


def test_tensorwise_large_groups(cuda):
  """TENSORWISE on tensors past the single-CTA cut-over (grid-wide reductions per iteration):
  OCTAV and MSE constants vs the oracle, odd sizes and an unaligned view included."""
  import torch
  from aeq_b200 import device
  from aeq_b200.algorithms.uniform_quantize import mse, octav
  for shape, idx in (((256, 4096), 1), ((37, 4099), 2)):
    w = O.synthetic_weight(*shape, index=idx)
    w[0, :7] = 0.0
    for bits in (4, 8):
      with np.errstate(all="ignore"):
        want = O.octav_clip(w, bits, (0, 1))
      got = device.octav_clip_rows(torch.from_numpy(w).to(cuda).reshape(1, -1), bits).cpu().numpy()
      np.testing.assert_allclose(got.reshape(-1), want.reshape(-1), rtol=1e-6)
      r = _run(octav, w, _cfg(bits, -1))
      ref = O.octav_requant(w, bits, per_channel=False)
      np.testing.assert_allclose(r.scale, ref["scale"], rtol=1e-6)
      _assert_ints(r.quantized_data, ref["q"])
      rm = _run(mse, w, _cfg(bits, -1))
      refm = O.mse_requant(w, bits, per_channel=False)
      np.testing.assert_allclose(rm.scale, refm["scale"], rtol=1e-6)
      _assert_ints(rm.quantized_data, refm["q"])
  flat = torch.from_numpy(np.concatenate([np.zeros(1, np.float32), O.synthetic_weight(64, 2048, 9).reshape(-1)])).to(cuda)
  view = flat[1:].reshape(1, -1)  # 4-byte aligned only
  want = O.octav_clip(flat[1:].cpu().numpy().reshape(64, 2048), 4, (0, 1))
  np.testing.assert_allclose(device.octav_clip_rows(view, 4).cpu().numpy().reshape(-1), want.reshape(-1), rtol=1e-6)
