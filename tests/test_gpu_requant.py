"""GPU parity: fused requantisation kernels vs the NumPy oracle, through the C ABI.

Bit-exact bar for integers, packed bytes, scales and zero points (min-max path
has order-free reductions and IEEE divides only).
"""
import numpy as np
import pytest

from oracle import aeq_oracle as O

pytestmark = pytest.mark.gpu


def _dev(x, cuda):
  import torch
  return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


def _special(w):
  """Injects the degenerate rows/blocks the reference handles (Appendix A)."""
  w = w.copy()
  if w.shape[0] >= 4:
    w[1, :] = 0.0                      # all-zero row -> scale 1e-9/qmax
    w[2, : w.shape[1] // 2] = 1e-8     # tiny values: blockwise scale flushes to 0
    w[3, 0] = 3.0e38                   # huge outlier
  return w


ROW_SHAPES = [(16, 8), (7, 33), (64, 128), (130, 256), (33, 1024), (96, 4096),
              (5, 8192), (9, 11008), (3, 16384), (2, 16512), (1, 4)]


@pytest.mark.parametrize("shape", ROW_SHAPES)
@pytest.mark.parametrize("bits,symmetric", [(8, True), (8, False), (4, True), (4, False), (2, True)])
def test_rows_matches_oracle(cuda, shape, bits, symmetric):
  from aeq_b200 import device
  w = _special(O.synthetic_weight(*shape, index=shape[1] % 97))
  ref = O.minmax_requant(w, bits, symmetric)
  want_packed = bits < 8 and (shape[1] * bits) % 8 == 0
  out = device.requant_rows(_dev(w, cuda), bits, symmetric, want_packed=want_packed)
  np.testing.assert_array_equal(out.scale.cpu().numpy(), ref["scale"])
  np.testing.assert_array_equal(out.zero_point.cpu().numpy(), ref["zero_point"].astype(np.int32))
  np.testing.assert_array_equal(out.q.cpu().numpy(), ref["q"])
  if want_packed:
    np.testing.assert_array_equal(out.packed.cpu().numpy(), O.pack_bits(bits, ref["q"]))


@pytest.mark.parametrize("shape", [(4, 32), (16, 64), (64, 256), (33, 2048), (130, 4096), (8, 11008), (3, 16384)])
@pytest.mark.parametrize("block", [32, 64, 128, 256])
@pytest.mark.parametrize("bits", [4, 8])
def test_blocks_matches_oracle(cuda, shape, block, bits):
  from aeq_b200 import device
  if shape[1] % block:
    pytest.skip("cols not divisible by block")
  w = _special(O.synthetic_weight(*shape, index=block + bits))
  ref = O.minmax_requant(w, bits, True, block=block)
  out = device.requant_blocks(_dev(w, cuda), block, bits, want_packed=(bits == 4))
  np.testing.assert_array_equal(out.scale.cpu().numpy(), ref["scale"])
  np.testing.assert_array_equal(out.scale_f16.cpu().numpy(), O.blockwise_scale_fp16(ref["scale"]))
  np.testing.assert_array_equal(out.q.cpu().numpy(), ref["q"])
  if bits == 4:
    np.testing.assert_array_equal(out.packed.cpu().numpy(), O.pack_bits(4, ref["q"]))


def test_nan_and_inf_rows(cuda):
  from aeq_b200 import device
  w = O.synthetic_weight(8, 256, 5)
  w[0, 3] = np.nan
  w[1, 7] = np.inf
  w[2, 9] = -np.inf
  with np.errstate(all="ignore"):
    ref = O.minmax_requant(w, 8, True)
  out = device.requant_rows(_dev(w, cuda), 8, True)
  np.testing.assert_array_equal(out.scale.cpu().numpy(), ref["scale"])
  np.testing.assert_array_equal(out.q.cpu().numpy(), ref["q"])


def test_unaligned_views(cuda):
  """Row/column-offset device views take the generic kernels."""
  import torch
  from aeq_b200 import device
  w = O.synthetic_weight(9, 260, 3)
  flat = torch.empty(w.size + 1, dtype=torch.float32, device=cuda)
  flat[1:] = _dev(w.reshape(-1), cuda)
  view = flat[1:].view(9, 260)
  out = device.requant_rows(view, 8, True)
  ref = O.minmax_requant(w, 8, True)
  np.testing.assert_array_equal(out.q.cpu().numpy(), ref["q"])
  w2 = O.synthetic_weight(5, 256, 4)
  flat = torch.empty(w2.size + 1, dtype=torch.float32, device=cuda)
  flat[1:] = _dev(w2.reshape(-1), cuda)
  out = device.requant_blocks(flat[1:].view(5, 256), 32, 4, want_packed=True)
  ref = O.minmax_requant(w2, 4, True, block=32)
  np.testing.assert_array_equal(out.q.cpu().numpy(), ref["q"])
  np.testing.assert_array_equal(out.packed.cpu().numpy(), O.pack_bits(4, ref["q"]))


def test_batched_entry_points_match_singles(cuda):
  """Ragged batch (stream classes 1 and 2, generic) == per-tensor calls == oracle."""
  from aeq_b200 import device
  shapes = [(64, 4096), (7, 33), (5, 11008), (130, 256), (3, 16384), (16, 8)] * 12  # 72 > 64 jobs
  ws = [_special(O.synthetic_weight(r, c, i)) for i, (r, c) in enumerate(shapes)]
  xs = [_dev(w, cuda) for w in ws]
  outs = device.requant_rows_batch(xs, 8, True)
  for w, o in zip(ws, outs):
    ref = O.minmax_requant(w, 8, True)
    np.testing.assert_array_equal(o.q.cpu().numpy(), ref["q"])
    np.testing.assert_array_equal(o.scale.cpu().numpy(), ref["scale"])
  bshapes = [(64, 4096), (5, 11008), (130, 256), (3, 16384), (4, 32)] * 14
  ws = [_special(O.synthetic_weight(r, c, 100 + i)) for i, (r, c) in enumerate(bshapes)]
  xs = [_dev(w, cuda) for w in ws]
  outs = device.requant_blocks_batch(xs, 32, 4, want_q=True, want_packed=True, want_scale=True)
  for w, o in zip(ws, outs):
    ref = O.minmax_requant(w, 4, True, block=32)
    np.testing.assert_array_equal(o.q.cpu().numpy(), ref["q"])
    np.testing.assert_array_equal(o.packed.cpu().numpy(), O.pack_bits(4, ref["q"]))
    np.testing.assert_array_equal(o.scale.cpu().numpy(), ref["scale"])
    np.testing.assert_array_equal(o.scale_f16.cpu().numpy(), O.blockwise_scale_fp16(ref["scale"]))


def test_full_size_properties(cuda):
  """BASELINE-size tensor (4096 x 4096): size-independent properties instead of an oracle run.
  Dequantised error <= scale/2 everywhere, per-row extreme hits +-qmax, pack/unpack round trip,
  idempotence of requantising the dequantised tensor."""
  import torch
  from aeq_b200 import device
  g = torch.Generator(device=cuda).manual_seed(7)
  x = torch.randn(4096, 4096, device=cuda, generator=g) * 0.02
  r = device.requant_rows(x, 8, True)
  dq = r.q.float() * r.scale
  assert bool(((dq - x).abs() <= r.scale * 0.50005).all())
  assert bool((r.q.abs().amax(dim=1) == 127).all())
  r2 = device.requant_rows(dq, 8, True)
  assert bool((r2.q == r.q).all())
  b = device.requant_blocks(x, 32, 4, want_packed=True)
  lo = (b.packed & 0xF).to(torch.int8)
  hi = (b.packed >> 4).to(torch.int8)
  unpacked = torch.stack([lo, hi], dim=1).reshape(4096, 4096)
  unpacked = torch.where(unpacked > 7, unpacked - 16, unpacked)
  assert bool((unpacked == b.q).all())
  sc = b.scale.repeat_interleave(32, dim=1)
  assert bool(((b.q.float() * sc - x).abs() <= sc * 0.50005).all())
  assert bool((b.scale_f16.float() == b.scale).all())


@pytest.mark.gpu
def test_scales_mirrored_into_peer_copies(cuda):
  """aeqb_requant_rows_batch_mirror_f32: the kernel stores every row's scale at the same offset
  of each peer mapping of the gathered buffer.  Here the "peers" are two more rows of one local
  buffer (the kernel only sees byte offsets); the cross-process mapping itself is exercised by
  bench.py at N > 1, which compares it with NCCL's all-gather of the same scales."""
  import ctypes
  import types
  import torch
  from aeq_b200 import _lib, device
  shapes = [(64, 4096), (24, 11008), (40, 1024), (48, 2560)]   # every tile-stream class
  ws = [O.synthetic_weight(r, c, index=60 + i) for i, (r, c) in enumerate(shapes)]
  xs = [_dev(w, cuda) for w in ws]
  slots = sum(r for r, _ in shapes)
  buf = torch.full((3, slots), -1.0, dtype=torch.float32, device=cuda)
  deltas = (ctypes.c_int64 * 2)(buf[1].data_ptr() - buf[0].data_ptr(), buf[2].data_ptr() - buf[0].data_ptr())
  mirror = types.SimpleNamespace(deltas_ptr=ctypes.cast(deltas, ctypes.c_void_p), n_peers=2)
  outs, off = [], 0
  for r, c in shapes:
    outs.append(device.Requantized(torch.empty((r, c), dtype=torch.int8, device=cuda), None,
                                   buf[0, off:off + r].view(r, 1),
                                   torch.empty((r, 1), dtype=torch.int32, device=cuda)))
    off += r
  device.requant_rows_batch(xs, 8, True, outs=outs, mirror=mirror)
  torch.cuda.synchronize()
  want = np.concatenate([O.minmax_requant(w, 8, True)["scale"].reshape(-1) for w in ws])
  got = buf.cpu().numpy()
  for k in range(3):
    np.testing.assert_array_equal(got[k], want)
  for w, o in zip(ws, outs):
    np.testing.assert_array_equal(o.q.cpu().numpy(), O.minmax_requant(w, 8, True)["q"])
  # the mirror is a second launch over every job's scale vector (16-byte stores where the alignment
  # allows, scalar otherwise), so a tensor that falls to the generic kernel is mirrored too, at an
  # unaligned offset of the gathered buffer
  odd_w = O.synthetic_weight(5, 33, index=3)
  odd = _dev(odd_w, cuda)
  buf.fill_(-1.0)
  o = device.Requantized(torch.empty((5, 33), dtype=torch.int8, device=cuda), None,
                         buf[0, 3:8].view(5, 1), torch.empty((5, 1), dtype=torch.int32, device=cuda))
  device.requant_rows_batch([odd], 8, True, outs=[o], mirror=mirror)
  torch.cuda.synchronize()
  got = buf.cpu().numpy()
  for k in range(3):
    np.testing.assert_array_equal(got[k, 3:8], O.minmax_requant(odd_w, 8, True)["scale"].reshape(-1))
    assert (got[k, :3] == -1.0).all() and (got[k, 8:] == -1.0).all()


@pytest.mark.gpu
def test_peer_buffer_alloc_free(cuda):
  """aeqb_peer_alloc hands out a zeroed device buffer and a 64-byte IPC handle."""
  import ctypes
  from aeq_b200 import _lib
  ptr, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
  _lib.call("aeqb_peer_alloc", 1 << 20, ctypes.byref(ptr), ctypes.cast(handle, ctypes.c_void_p))
  assert ptr.value and any(bytes(handle))
  _lib.call("aeqb_peer_free", ptr)
  with pytest.raises(_lib.AeqbError):
    _lib.call("aeqb_peer_alloc", 0, ctypes.byref(ptr), ctypes.cast(handle, ctypes.c_void_p))


@pytest.mark.gpu
@pytest.mark.parametrize("block", [32, 64, 128, 256])
def test_block_scales_mirrored_into_peer_copies(cuda, block):
  """aeqb_requant_blocks_batch_mirror_f32: every block's fp16 scale lands at the same offset of
  each peer mapping of the gathered buffer (here two more rows of one local buffer; the kernel only
  sees byte offsets).  Ragged tensors: full tiles, a partial last tile, a partial last slice, an
  odd number of scales per warp strip."""
  import ctypes
  import types
  import torch
  from aeq_b200 import _lib, device
  shapes = [(64, 4096), (24, 11008 // 256 * 256), (3, 2048), (5, 1280), (1, 256), (130, 256)]
  ws = [_special(O.synthetic_weight(r, c, index=80 + i)) for i, (r, c) in enumerate(shapes)]
  xs = [_dev(w, cuda) for w in ws]
  counts = [w.size // block for w in ws]
  offs = np.concatenate([[0], np.cumsum([(n + 7) // 8 * 8 for n in counts])])  # 16-byte aligned starts
  buf = torch.full((3, int(offs[-1])), -1.0, dtype=torch.float16, device=cuda)
  deltas = (ctypes.c_int64 * 2)(buf[1].data_ptr() - buf[0].data_ptr(), buf[2].data_ptr() - buf[0].data_ptr())
  mirror = types.SimpleNamespace(deltas_ptr=ctypes.cast(deltas, ctypes.c_void_p), n_peers=2)
  outs = []
  for w, n, o in zip(ws, counts, offs[:-1]):
    outs.append(device.Requantized(None, torch.empty(w.size // 2, dtype=torch.uint8, device=cuda), None, None,
                                   buf[0, int(o):int(o) + n].view(w.shape[0], -1)))
  device.requant_blocks_batch(xs, block, 4, outs=outs, mirror=mirror)
  torch.cuda.synchronize()
  got = buf.cpu().numpy()
  for w, n, o, out in zip(ws, counts, offs[:-1], outs):
    ref = O.minmax_requant(w, 4, True, block=block)
    want = O.blockwise_scale_fp16(ref["scale"]).reshape(-1)
    for k in range(3):
      np.testing.assert_array_equal(got[k, int(o):int(o) + n].view(np.uint16), want.view(np.uint16))
    np.testing.assert_array_equal(out.packed.cpu().numpy(), O.pack_bits(4, ref["q"]))
  # padding between tensors was not touched on any copy
  pad = np.ones(int(offs[-1]), bool)
  for n, o in zip(counts, offs[:-1]):
    pad[int(o):int(o) + n] = False
  assert (got[:, pad] == -1.0).all()
  # a job that also wants the one-value-per-byte integers cannot be mirrored: loud error
  bad = [device.Requantized(torch.empty(ws[0].shape, dtype=torch.int8, device=cuda), outs[0].packed, None, None,
                            outs[0].scale_f16)]
  with pytest.raises(_lib.AeqbError, match="cannot mirror"):
    device.requant_blocks_batch(xs[:1], block, 4, outs=bad, mirror=mirror)
