"""GPU parity at the sizes bench.py times (BASELINE.json configs[1], [2], [4]).

The small-shape tests compare every degenerate case with the oracle; these close the gap the
round-1 verdict named: the configurations that are TIMED were only property-checked.
  * [4096, 4096] per-channel INT8 and block-32 packed INT4: bit-exact against the oracle.
  * the 64-tensor persistent launches bench.py times (every CTA's shared-memory ring wraps
    >= 7 times, the producer walks a 64-entry job table): sampled tensors bit-exact.
  * GPTQ at K = 4096 and K = 11008 (two-level Cholesky with rank-128 trailing updates,
    6 - 8 levels of recursive doubling, the 272-tile tcgen05 grid): H_damped @ Hinv = I in
    float64 on the device, Hessian against a float64 product, proxy loss of a [4096, 4096] OBS
    run against a float64 right-looking replay of gptq.py:131-216 on the same H^-1.
"""
import numpy as np
import pytest

from oracle import aeq_oracle as O

pytestmark = pytest.mark.gpu


def _device_weights(cuda, n, seed, rows=4096, cols=4096):
  """bench.py's generator: N(0, 0.02) with one x20 outlier per 1024 elements."""
  import torch
  g = torch.Generator(device=cuda).manual_seed(seed)
  ws = []
  for _ in range(n):
    w = torch.randn(rows, cols, device=cuda, generator=g) * 0.02
    w.view(-1)[::1024] *= 20.0
    ws.append(w)
  return ws


def test_fc4096_int8_per_channel_bit_exact(cuda):
  import torch
  from aeq_b200 import device
  w = O.synthetic_weight(4096, 4096, index=1)
  w[17, :] = 0.0  # a dead output channel: scale 1e-9 / 127, all integers 0
  ref = O.minmax_requant(w, 8, True)
  out = device.requant_rows(torch.from_numpy(w).to(cuda), 8, True)
  np.testing.assert_array_equal(out.scale.cpu().numpy(), ref["scale"])
  np.testing.assert_array_equal(out.zero_point.cpu().numpy(), ref["zero_point"].astype(np.int32))
  np.testing.assert_array_equal(out.q.cpu().numpy(), ref["q"])


def test_fc4096_int4_block32_packed_bit_exact(cuda):
  import torch
  from aeq_b200 import device
  w = O.synthetic_weight(4096, 4096, index=2)
  w[5, 64:128] = 0.0  # two dead blocks: the bf16 -> fp16 rounding flushes their scale to 0
  with np.errstate(all="ignore"):
    ref = O.minmax_requant(w, 4, True, block=32)
  out = device.requant_blocks(torch.from_numpy(w).to(cuda), 32, 4, want_packed=True)
  np.testing.assert_array_equal(out.scale.cpu().numpy(), ref["scale"])
  np.testing.assert_array_equal(out.scale_f16.cpu().numpy(), O.blockwise_scale_fp16(ref["scale"]))
  np.testing.assert_array_equal(out.q.cpu().numpy(), ref["q"])
  np.testing.assert_array_equal(out.packed.cpu().numpy(), O.pack_bits(4, ref["q"]))


def test_bench_stack_64_tensors_rows_sampled_vs_oracle(cuda):
  """The launch bench.py times: 64 x [4096, 4096] in ONE persistent launch."""
  from aeq_b200 import _lib, device
  ws = _device_weights(cuda, 64, seed=1000)
  before = _lib.load().aeqb_launch_count()
  outs = device.requant_rows_batch(ws, 8, True)
  assert _lib.load().aeqb_launch_count() - before == 1, "the 64-tensor stack must be one launch"
  for i in (0, 31, 63):
    ref = O.minmax_requant(ws[i].cpu().numpy(), 8, True)
    np.testing.assert_array_equal(outs[i].scale.cpu().numpy(), ref["scale"])
    np.testing.assert_array_equal(outs[i].q.cpu().numpy(), ref["q"])
  # INT4 packed per-channel through the same launch shape (cfg 5's weight side)
  outs = device.requant_rows_batch(ws, 4, True, want_q=False, want_packed=True)
  for i in (7, 62):
    ref = O.minmax_requant(ws[i].cpu().numpy(), 4, True)
    np.testing.assert_array_equal(outs[i].packed.cpu().numpy(), O.pack_bits(4, ref["q"]))
    np.testing.assert_array_equal(outs[i].scale.cpu().numpy(), ref["scale"])


@pytest.mark.parametrize("bits,packed", [(8, False), (4, True)])
def test_large_batch_of_llama_shapes_takes_96k_tiles_bit_exact(cuda, bits, packed):
  """A large batch (>= 96 MiB; this one is 548 MiB) takes the 96 KiB-tile class whatever it writes
  (rows_job_class with batch_bytes): rows of 11008 / 5120 / 3584 / 2560 floats leave partial tiles whose row maxima are
  merged run-wise, 1536-float rows are shorter than the 16 consumer warps' chunks (slot-wise path),
  6144- and 2048-float rows split evenly over the warps (folded in registers, spr = 3 and 1).  Every
  tensor of the batch bit-exact against the oracle, INT8 and packed INT4."""
  import torch
  from aeq_b200 import device
  shapes = [(4096, 11008), (4096, 11008), (4096, 5120), (4096, 3584), (900, 2560), (2050, 1536), (1030, 6144),
            (3000, 2048)]
  assert sum(r * c * 4 for r, c in shapes) >= 512 << 20
  ws = [O.synthetic_weight(r, c, index=300 + i) for i, (r, c) in enumerate(shapes)]
  ws[2][5, :] = 0.0
  ws[3][0, 7] = np.inf
  outs = device.requant_rows_batch([torch.from_numpy(w).to(cuda) for w in ws], bits, True, want_q=not packed,
                                   want_packed=packed)
  for w, o in zip(ws, outs):
    with np.errstate(all="ignore"):
      ref = O.minmax_requant(w, bits, True)
    np.testing.assert_array_equal(o.scale.cpu().numpy(), ref["scale"])
    if packed:
      np.testing.assert_array_equal(o.packed.cpu().numpy(), O.pack_bits(bits, ref["q"]))
    else:
      np.testing.assert_array_equal(o.q.cpu().numpy(), ref["q"])


def test_bench_stack_64_tensors_blocks_sampled_vs_oracle(cuda):
  from aeq_b200 import _lib, device
  ws = _device_weights(cuda, 64, seed=1001)
  before = _lib.load().aeqb_launch_count()
  outs = device.requant_blocks_batch(ws, 32, 4, want_q=False, want_packed=True, want_scale=True)
  assert _lib.load().aeqb_launch_count() - before == 1
  for i in (0, 40, 63):
    ref = O.minmax_requant(ws[i].cpu().numpy(), 4, True, block=32)
    np.testing.assert_array_equal(outs[i].packed.cpu().numpy(), O.pack_bits(4, ref["q"]))
    np.testing.assert_array_equal(outs[i].scale.cpu().numpy(), ref["scale"])
    np.testing.assert_array_equal(outs[i].scale_f16.cpu().numpy(), O.blockwise_scale_fp16(ref["scale"]))


def test_gemma2b_layer_blocks_batch_vs_oracle(cuda):
  """configs[2]'s ragged tensor set (one decoder layer: q, o, k, v, gate, up, down), repeated so
  that the batch exceeds one 64-job launch: every tensor of the first and last layer bit-exact."""
  from aeq_b200 import device
  shapes = [(2048, 2048), (2048, 2048), (256, 2048), (256, 2048), (16384, 2048), (16384, 2048),
            (2048, 16384)]
  ws = []
  for layer in range(10):
    for r, c in shapes:
      ws.extend(_device_weights(cuda, 1, seed=50 + len(ws), rows=r, cols=c))
  outs = device.requant_blocks_batch(ws, 32, 4, want_q=False, want_packed=True)
  for i in list(range(7)) + list(range(63, 70)):
    ref = O.minmax_requant(ws[i].cpu().numpy(), 4, True, block=32)
    np.testing.assert_array_equal(outs[i].packed.cpu().numpy(), O.pack_bits(4, ref["q"]))
    np.testing.assert_array_equal(outs[i].scale_f16.cpu().numpy(), O.blockwise_scale_fp16(ref["scale"]))


# ---------------------------------------------------------------------------------- GPTQ at size
def _hessian_f64(x, num_samples):
  """(2 / num_samples) * X^T X in float64 on the device (gptq.py:100-106 evaluated exactly)."""
  xd = x.reshape(-1, x.shape[-1]).double()
  return (2.0 / num_samples) * (xd.T @ xd)


@pytest.mark.parametrize("k,tokens", [(4096, 16384), (11008, 16384)])
@pytest.mark.parametrize("two_level", ["dmma", "0", "1000000"])
def test_hessian_inverse_at_layer_size(cuda, monkeypatch, k, tokens, two_level):
  """Every Cholesky variant at the Llama-7B orders (DMMA + lookahead: the default and the one the
  bench times; two-level SIMT; single-level): the float32 inverse against the damped float64
  Hessian it was computed from."""
  import torch
  from aeq_b200 import device
  if k == 11008 and two_level == "1000000":
    pytest.skip("the single-level Cholesky is the small-K variant; K = 4096 covers it")
  if two_level == "dmma":
    monkeypatch.setenv("AEQB_CHOL_DMMA_MIN_K", "0")
  else:
    monkeypatch.setenv("AEQB_CHOL_DMMA_MIN_K", "1000000")
    monkeypatch.setenv("AEQB_CHOL_TWO_LEVEL_MIN_K", two_level)
  g = torch.Generator(device=cuda).manual_seed(k)
  x = torch.randn(tokens, k, device=cuda, generator=g)
  x *= 1.0 + (torch.arange(k, device=cuda) % 7).float()  # anisotropic input features
  h = device.xtx(x, 2.0 / 8)
  want_h = _hessian_f64(x, 8)
  assert float((h - want_h).abs().max()) <= 4e-6 * float(torch.diagonal(want_h).max())
  assert bool((h == h.T).all())
  del x, want_h
  hd = h.clone()
  hinv = device.hessian_inverse(hd, 0.01, keep_damped_diagonal=True).double()
  # the caller's matrix now holds the damped diagonal (gptq.py:114,123) and nothing else changed
  d = torch.diagonal(h)
  want_diag = torch.where(d == 0, torch.ones_like(d), d)
  want_diag = want_diag + 0.01 * want_diag.mean()
  assert float((torch.diagonal(hd) - want_diag).abs().max()) <= 1e-12 * float(want_diag.max())
  resid = hd @ hinv
  resid.diagonal().sub_(1.0)
  assert float(resid.abs().max()) <= 2e-3, float(resid.abs().max())
  assert float((hinv - hinv.T).abs().max()) <= 1e-5 * float(hinv.abs().max())


def _replay_f64(w, hinv, scale, bits):
  """gptq._apply_gptq (gptq.py:131-216) restated in float64 torch ops, symmetric per-channel."""
  import torch
  qmin, qmax = -(2 ** (bits - 1)), 2 ** (bits - 1) - 1
  if bits >= 8:
    qmin += 1
  W = w.double().clone()
  H = hinv.double()
  s = scale.double().reshape(-1)
  rows, k = W.shape
  q_all = torch.empty((rows, k), dtype=torch.int8, device=w.device)
  for b0 in range(0, k, 64):
    b1 = min(b0 + 64, k)
    wb = W[:, b0:b1].clone()
    err = torch.zeros_like(wb)
    for i in range(b1 - b0):
      col = b0 + i
      qc = torch.clamp(torch.round(wb[:, i] / s), qmin, qmax)
      q_all[:, col] = qc.to(torch.int8)
      e = (wb[:, i] - qc * s) / H[col, col]
      err[:, i] = e
      if i < b1 - b0 - 1:
        wb[:, i + 1:] -= torch.outer(e, H[col, col + 1:b1])
    W[:, b1:] -= err @ H[b0:b1, b1:]
  return q_all


@pytest.mark.parametrize("rows,k", [(4096, 4096), (1024, 11008)])
def test_obs_loop_at_layer_size_vs_f64_replay(cuda, rows, k):
  """The OBS loop on a full layer: integers against a float64 right-looking replay that uses the
  same H^-1, proxy loss tr(E H E^T) within 0.1 %, and far below plain rounding."""
  import torch
  from aeq_b200 import device
  g = torch.Generator(device=cuda).manual_seed(rows + k)
  x = torch.randn(max(8192, 2 * k), k, device=cuda, generator=g)
  # correlated input features, so that the OBS updates matter (plain rounding is clearly worse)
  x = x + 0.5 * x.roll(1, dims=1) + 0.25 * x.roll(2, dims=1)
  h = device.xtx(x, 2.0 / 8)
  del x
  hinv = device.hessian_inverse(h.clone(), 0.01)
  w = _device_weights(cuda, 1, seed=k, rows=rows, cols=k)[0]
  scale = w.abs().amax(dim=1) / 7.0
  q = device.gptq_quantize(w, hinv, scale, None, 0, 4, True)
  q_ref = _replay_f64(w, hinv, scale, 4)
  q_rtn = torch.clamp(torch.round(w / scale[:, None]), -8, 7)

  def loss(qq):
    e = w.double() - qq.double() * scale.double()[:, None]
    return float(((e @ h) * e).sum())

  mism = float((q != q_ref).double().mean())
  assert mism <= 2e-2, mism
  assert int((q.int() - q_ref.int()).abs().max()) <= 2
  l, l_ref, l_rtn = loss(q), loss(q_ref), loss(q_rtn)
  assert abs(l - l_ref) <= 1e-3 * l_ref, (l, l_ref)
  assert l < 0.97 * l_rtn, (l, l_rtn)


def test_more_than_64_tensors_travel_as_one_launch(cuda, monkeypatch):
  """A model's weight set above the 64 inline kernel-parameter jobs goes out as ONE persistent
  launch whose job table sits in device memory (uploaded stream-ordered from a pinned ring):
  every tensor bit-exact, rows and blocks kernels, with and without mirrored scales; the
  chained-launch path (AEQB_NO_JOB_TABLE) is covered by the 72-job test in test_gpu_requant."""
  import ctypes
  import types
  import torch
  from aeq_b200 import _lib, device
  lib = _lib.load()
  shapes = [(40 + (i % 5), 4096) for i in range(130)]
  ws = [O.synthetic_weight(r, c, index=500 + i) for i, (r, c) in enumerate(shapes)]
  xs = [torch.from_numpy(w).to(cuda) for w in ws]
  before = lib.aeqb_launch_count()
  outs = device.requant_rows_batch(xs, 8, True)
  assert lib.aeqb_launch_count() - before == 1
  for w, o in zip(ws, outs):
    ref = O.minmax_requant(w, 8, True)
    np.testing.assert_array_equal(o.q.cpu().numpy(), ref["q"])
    np.testing.assert_array_equal(o.scale.cpu().numpy(), ref["scale"])
  # the same call again reuses the cached host-side job list (steady-state callers) and the ring
  for _ in range(12):
    outs2 = device.requant_rows_batch(xs, 8, True, outs=outs)
  np.testing.assert_array_equal(outs2[129].q.cpu().numpy(), O.minmax_requant(ws[129], 8, True)["q"])
  # mirrored scales through the table path
  slots = sum(r for r, _ in shapes)
  buf = torch.full((2, slots), -1.0, dtype=torch.float32, device=cuda)
  deltas = (ctypes.c_int64 * 1)(buf[1].data_ptr() - buf[0].data_ptr())
  mirror = types.SimpleNamespace(deltas_ptr=ctypes.cast(deltas, ctypes.c_void_p), n_peers=1)
  mo, off = [], 0
  for r, c in shapes:
    mo.append(device.Requantized(torch.empty((r, c), dtype=torch.int8, device=cuda), None,
                                 buf[0, off:off + r].view(r, 1), torch.empty((r, 1), dtype=torch.int32, device=cuda)))
    off += r
  before = lib.aeqb_launch_count()
  device.requant_rows_batch(xs, 8, True, outs=mo, mirror=mirror)
  assert lib.aeqb_launch_count() - before == 2  # the requantisation + the 16-byte-store mirror of the scales
  want = np.concatenate([O.minmax_requant(w, 8, True)["scale"].reshape(-1) for w in ws])
  got = buf.cpu().numpy()
  np.testing.assert_array_equal(got[0], want)
  np.testing.assert_array_equal(got[1], want)
  # blocks kernel
  before = lib.aeqb_launch_count()
  bo = device.requant_blocks_batch(xs, 32, 4, want_q=False, want_packed=True, want_scale=True)
  assert lib.aeqb_launch_count() - before == 1
  for w, o in zip(ws[::7], bo[::7]):
    ref = O.minmax_requant(w, 4, True, block=32)
    np.testing.assert_array_equal(o.packed.cpu().numpy(), O.pack_bits(4, ref["q"]))
    np.testing.assert_array_equal(o.scale.cpu().numpy(), ref["scale"])
    np.testing.assert_array_equal(o.scale_f16.cpu().numpy(), O.blockwise_scale_fp16(ref["scale"]))
