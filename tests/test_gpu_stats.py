"""GPU parity of the statistics and unfused kernels vs the oracle / reference fixtures."""
import os

import numpy as np
import pytest

from oracle import aeq_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _dev(x, cuda):
  import torch
  return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


def test_activation_minmax_fixtures_and_ema(cuda):
  from aeq_b200.algorithms.uniform_quantize import common_quantize as cq
  from aeq_b200.utils import qsv_utils
  g = np.load(os.path.join(GOLD, "calibration.npz"))
  q = {}
  for j in range(6):
    mm = cq.get_activation_min_max(g[f"a{j}"], -3e38, 3e38)
    np.testing.assert_array_equal(mm["min"], g[f"a{j}_min"])
    np.testing.assert_array_equal(mm["max"], g[f"a{j}_max"])
    assert mm["min"].shape == (1, 1, 1) and mm["min"].dtype == np.float32
    q = qsv_utils.moving_average_update(q, mm)
  np.testing.assert_array_equal(q["min"], g["ema_min"])
  np.testing.assert_array_equal(q["max"], g["ema_max"])


@pytest.mark.parametrize("n", [1, 3, 5, 1023, 4096, 1 << 20, (1 << 22) + 7])
def test_minmax_tensor_sizes_and_alignment(cuda, n):
  import torch
  from aeq_b200 import device
  x = np.random.default_rng(n).standard_normal(n + 3).astype(np.float32)
  t = _dev(x, cuda)
  for off in (0, 1, 3):
    got = device.minmax_tensor(t[off:off + n]).cpu().numpy()
    assert got[0] == x[off:off + n].min() and got[1] == x[off:off + n].max()


def test_minmax_tensor_nan_inf_semantics(cuda):
  from aeq_b200 import device
  x = np.array([1.0, np.nan, -2.0, np.inf, -np.inf, 5.0], np.float32)
  got = device.minmax_tensor(_dev(x, cuda), -3e38, 3e38).cpu().numpy()
  mn, mx = O.activation_minmax(x)
  assert got[0] == mn.item() and got[1] == mx.item()  # NaN fails both comparisons, inf filtered
  raw = device.minmax_tensor(_dev(x, cuda)).cpu().numpy()
  assert np.isnan(raw[0]) and np.isnan(raw[1])  # np.min / np.max propagate NaN
  empty = device.minmax_tensor(_dev(np.array([3.2e38, 3.3e38], np.float32), cuda), -3e38, 3e38).cpu().numpy()
  assert empty[1] == np.float32(3.3e38)  # nothing passes -> raw fallback


@pytest.mark.parametrize("shape", [(5, 7), (16, 256), (9, 4096), (3, 11008)])
def test_row_and_block_minmax(cuda, shape):
  from aeq_b200 import device
  w = O.synthetic_weight(*shape, index=3)
  mn, mx, ss = device.row_stats(_dev(w, cuda), want_sumsq=True)
  np.testing.assert_array_equal(mn.cpu().numpy(), w.min(axis=1, keepdims=True))
  np.testing.assert_array_equal(mx.cpu().numpy(), w.max(axis=1, keepdims=True))
  np.testing.assert_allclose(ss.cpu().numpy(), (w.astype(np.float64) ** 2).sum(axis=1, keepdims=True), rtol=2e-6)
  for block in (32, 64, 128, 256):
    if shape[1] % block:
      continue
    bmn, bmx = device.minmax_blocks(_dev(w, cuda), block)
    rmn, rmx = O.weight_minmax(w, block)
    np.testing.assert_array_equal(bmn.cpu().numpy(), rmn)
    np.testing.assert_array_equal(bmx.cpu().numpy(), rmx)


def test_init_tensor_min_max_granularities(cuda):
  from aeq_b200 import qtyping
  from aeq_b200.algorithms.uniform_quantize import common_quantize as cq
  from tests import synthetic_graph as sg
  w = O.synthetic_weight(12, 256, 8)
  op, _ = sg.fc_graph(w)
  G = qtyping.QuantGranularity
  for gran, block, per in ((G.CHANNELWISE, 0, True), (G.TENSORWISE, 0, False), (G.BLOCKWISE_64, 64, True)):
    mm = cq.init_tensor_min_max(w, sg.op_info(op, qtyping.TensorQuantizationConfig(4, True, gran)))
    rmn, rmx = O.weight_minmax(w, block, per)
    np.testing.assert_array_equal(mm["min"], rmn)
    np.testing.assert_array_equal(mm["max"], rmx)


def test_pack_fixtures(cuda):
  from aeq_b200 import device
  g = np.load(os.path.join(GOLD, "pack.npz"))
  for bits, n in ((4, 15), (4, 4096), (2, 13), (2, 1024)):
    got = device.pack_bits(_dev(g[f"b{bits}_n{n}_in"], cuda), bits).cpu().numpy()
    np.testing.assert_array_equal(got, g[f"b{bits}_n{n}_out"])


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_quantize_dequantize_any_axis(cuda, axis):
  from aeq_b200 import device
  x = np.random.default_rng(axis).standard_normal((6, 10, 14)).astype(np.float32)
  shape = [1, 1, 1]
  shape[axis] = x.shape[axis]
  scale = (np.abs(x).max(axis=tuple(a for a in range(3) if a != axis), keepdims=True) / 127).astype(np.float32)
  zp = np.arange(x.shape[axis], dtype=np.int8).reshape(shape) - 3
  ref = O.quantize(x, scale, zp, 8, False)
  inner = int(np.prod(x.shape[axis + 1:]))
  q = device.quantize(_dev(x, cuda), _dev(scale.reshape(-1), cuda), _dev(zp.reshape(-1).astype(np.int32), cuda),
                      8, False, x.shape[axis], inner)
  np.testing.assert_array_equal(q.cpu().numpy(), ref)
  dq = device.dequantize(q, _dev(scale.reshape(-1), cuda), _dev(zp.reshape(-1).astype(np.int32), cuda),
                         x.shape[axis], inner, wrap8=True)
  np.testing.assert_array_equal(dq.cpu().numpy(), O.dequantize(ref, scale, zp).astype(np.float32))


def test_scale_zp_kernel_matches_oracle(cuda):
  from aeq_b200 import device
  rng = np.random.default_rng(5)
  mn = -np.abs(rng.standard_normal(1000)).astype(np.float32)
  mx = np.abs(rng.standard_normal(1000)).astype(np.float32)
  mn[:3], mx[:3] = 0.0, 0.0
  mn[3], mx[4] = -1e-12, 7e4
  clip = (np.abs(rng.standard_normal(1000)) * 0.5).astype(np.float32)
  for bits in (2, 4, 8):
    for sym in (True, False):
      for blockwise in (False, True):
        if blockwise and not sym:
          continue
        for c in (None, clip):
          if blockwise and c is not None:
            continue  # float64 intermediate in the reference; covered by OCTAV tests with tolerance
          zp, sc, _ = device.scale_zp_from_minmax(_dev(mn, cuda), _dev(mx, cuda), bits, sym, blockwise,
                                                  None if c is None else _dev(c, cuda))
          with np.errstate(all="ignore"):
            ozp, osc = O.scale_zp(mn, mx, bits, sym, blockwise, c)
          np.testing.assert_array_equal(sc.cpu().numpy(), osc)
          np.testing.assert_array_equal(zp.cpu().numpy().astype(ozp.dtype), ozp)  # uqt:585 cast wraps


def test_minmax_tensors_batched(cuda):
  """70 ragged tensors (> 64: two launches), unaligned views, NaN / inf / all-filtered cases,
  twice in a row (the workspace resets itself)."""
  import torch
  from aeq_b200 import device
  rng = np.random.default_rng(11)
  sizes = [1, 2, 3, 5, 1023, 1024, 1025, 4096, 4097, 70000, 262144 + 3] * 6 + [8, 16, 100, 9]
  arrs = [rng.standard_normal(n).astype(np.float32) * (1 + i % 5) for i, n in enumerate(sizes)]
  arrs[3][1] = np.nan
  arrs[9][5] = np.inf
  arrs[10][7] = -np.inf
  arrs[12][:] = 3.2e38            # everything filtered -> raw fallback
  arrs[20][0] = 3.39e38
  big = torch.from_numpy(np.concatenate([np.zeros(1, np.float32)] + arrs)).to(cuda)
  xs, off = [], 1                 # views at odd offsets: not 16-byte aligned
  for a in arrs:
    xs.append(big[off:off + a.size])
    off += a.size
  for _ in range(2):
    got = device.minmax_tensors(xs, -3e38, 3e38).cpu().numpy()
    for a, g in zip(arrs, got):
      with np.errstate(all="ignore"):
        mn, mx = O.activation_minmax(a)
      np.testing.assert_array_equal(g, np.array([mn.item(), mx.item()], np.float32))
  raw = device.minmax_tensors(xs).cpu().numpy()
  for a, g in zip(arrs, raw):
    with np.errstate(all="ignore"):
      np.testing.assert_array_equal(g, np.array([a.min(), a.max()], np.float32))


def test_ema_sequence_matches_the_calibrator_fold(cuda):
  """aeqb_ema_sequence_f32 == qsv_utils.moving_average_update folded in batch order, bit-exact
  (first batch verbatim, fp32 weak-scalar arithmetic, no FMA contraction)."""
  import torch
  from aeq_b200 import device
  rng = np.random.default_rng(12)
  for n in (1, 2, 7, 512):
    mins = (-rng.random(n, dtype=np.float32) * 5 - 0.1).astype(np.float32)
    maxs = (rng.random(n, dtype=np.float32) * 7 + 0.1).astype(np.float32)
    want = O.ema_sequence([np.full((1, 1, 1), m, np.float32) for m in mins],
                          [np.full((1, 1, 1), m, np.float32) for m in maxs])
    pairs = torch.from_numpy(np.stack([mins, maxs], axis=1)).to(cuda)
    got = device.ema_sequence(pairs).cpu().numpy()
    assert got[0] == want[0].item() and got[1] == want[1].item(), (n, got, want)
  # end to end: per-batch filtered min / max of activation batches, then the fold
  acts = [O.synthetic_activation((4, 64, 256), 40 + i) * (1 + i) for i in range(9)]
  acts[3][0, 0, 0] = -3.3e38  # filtered by the (-3e38, 3e38) window
  mm = device.minmax_tensors([torch.from_numpy(a).to(cuda) for a in acts], -3e38, 3e38)
  got = device.ema_sequence(mm).cpu().numpy()
  q = {}
  for a in acts:
    q = O.ema_update(q, O.activation_qsv(a))
  assert got[0] == q["min"].item() and got[1] == q["max"].item()
