"""GPU parity of the host-buffer C-ABI entry points (NumPy in -> NumPy out).

aeqb_host_requant_{rows,blocks}_batch_f32 chunk tensors into ~32 MiB row ranges and
pipeline H2D -> kernel -> D2H; chunking must not change one bit of the result.
"""
import numpy as np
import pytest

from oracle import aeq_oracle as O

pytestmark = pytest.mark.gpu


def _check_rows(ws, outs, bits, symmetric, packed):
  for w, o in zip(ws, outs):
    ref = O.minmax_requant(w, bits, symmetric)
    np.testing.assert_array_equal(o[0], ref["q"])
    np.testing.assert_array_equal(o[2], ref["scale"])
    np.testing.assert_array_equal(o[3], ref["zero_point"].astype(np.int32))
    if packed:
      np.testing.assert_array_equal(o[1], O.pack_bits(bits, ref["q"]))


@pytest.mark.parametrize("bits,symmetric", [(8, True), (8, False), (4, True)])
def test_host_rows_pageable_ragged(cuda, bits, symmetric):
  """Pageable NumPy arrays of ragged shapes (stream classes and the generic kernel)."""
  from aeq_b200 import host
  shapes = [(64, 4096), (7, 33), (5, 11008), (130, 256), (3, 16384), (16, 8), (1, 4), (2100, 4096)]
  ws = [O.synthetic_weight(r, c, i) for i, (r, c) in enumerate(shapes)]
  packed = bits == 4
  outs = host.requant_rows(ws, bits, symmetric, want_packed=False)
  _check_rows(ws, outs, bits, symmetric, False)
  if packed:
    even = [w for w in ws if w.shape[1] % 2 == 0]
    outs = host.requant_rows(even, bits, symmetric, want_packed=True)
    _check_rows(even, outs, bits, symmetric, True)


def test_host_rows_pinned_and_chunked(cuda):
  """A tensor larger than one 32 MiB chunk, page-locked in and out (DMA in place)."""
  from aeq_b200 import host
  w = O.synthetic_weight(4200, 4096, 11)  # 65.6 MiB -> 3 chunks
  h = host.pinned_empty(w.shape, np.float32)
  h[...] = w
  outs = host.requant_rows([h], 8, True, alloc=host.pinned_empty)
  _check_rows([w], outs, 8, True, False)
  # reusing the outputs (the benchmark's steady state) gives the same bytes
  again = host.requant_rows([h], 8, True, outs=outs)
  _check_rows([w], again, 8, True, False)


def test_host_unaligned_input(cuda):
  """mmap'd flatbuffer views carry no alignment guarantee (tfl_flatbuffer_utils.py:254-263)."""
  from aeq_b200 import host
  w = O.synthetic_weight(33, 1024, 5)
  raw = np.empty(w.nbytes + 4, dtype=np.uint8)
  view = raw[1:1 + w.nbytes].view(np.uint8)
  view[...] = w.view(np.uint8).reshape(-1)
  # 1-byte offset: not even 4-byte aligned; NumPy exposes it as an unaligned float32 array
  unaligned = np.frombuffer(raw, dtype=np.uint8, count=w.nbytes, offset=1).view(np.float32).reshape(w.shape)
  assert not unaligned.flags.aligned
  outs = host.requant_rows([unaligned], 8, True)
  _check_rows([w], outs, 8, True, False)


@pytest.mark.parametrize("block", [32, 128])
def test_host_blocks(cuda, block):
  from aeq_b200 import host
  shapes = [(64, 4096), (5, 11008 if block == 32 else 11008 // 128 * 128), (130, 256), (2100, 4096), (4, block)]
  ws = [O.synthetic_weight(r, c, 50 + i) for i, (r, c) in enumerate(shapes)]
  outs = host.requant_blocks(ws, block, 4, want_q=True, want_packed=True, want_scale=True,
                             want_scale_f16=True)
  for w, o in zip(ws, outs):
    ref = O.minmax_requant(w, 4, True, block=block)
    np.testing.assert_array_equal(o[0], ref["q"])
    np.testing.assert_array_equal(o[1], O.pack_bits(4, ref["q"]))
    np.testing.assert_array_equal(o[2], ref["scale"])
    np.testing.assert_array_equal(o[3], O.blockwise_scale_fp16(ref["scale"]))


def test_host_errors(cuda):
  from aeq_b200 import host
  with pytest.raises(ValueError, match="not divisible by block size"):
    host.requant_blocks([np.zeros((4, 48), np.float32)], 32, 4)
  with pytest.raises(ValueError, match="float32"):
    host.requant_rows([np.zeros((4, 48), np.float64)], 8)


def test_host_rows_read_only_memmap_views(cuda, tmp_path):
  """The boundary the reference presents (SURVEY.md §8b "Ownership"): read-only NumPy views on an
  mmap'd file in, fresh NumPy arrays out; many chunks in flight (the slot ring wraps several times,
  staging and copy-out run on the worker threads)."""
  from aeq_b200 import _lib, host
  ws = [O.synthetic_weight(2100, 4096, 70 + i) for i in range(4)] + [O.synthetic_weight(9, 4096, 75)]
  path = tmp_path / "weights.bin"
  with open(path, "wb") as f:
    f.write(b"\x00" * 3)  # payloads start at an odd offset: no alignment guarantee inside a flatbuffer
    for w in ws:
      f.write(w.tobytes())
  mm = np.memmap(path, dtype=np.uint8, mode="r")
  views, off = [], 3
  for w in ws:
    views.append(mm[off:off + w.nbytes].view(np.float32).reshape(w.shape))
    off += w.nbytes
  assert not views[0].flags.writeable and not views[0].flags.aligned
  outs = host.requant_rows(views, 8, True)
  _check_rows(ws, outs, 8, True, False)
  assert _lib.load().aeqb_host_worker_threads() >= 1
  outs4 = host.requant_blocks(views, 32, 4, want_q=True, want_packed=True, want_scale=True, want_scale_f16=True)
  for w, o in zip(ws, outs4):
    ref = O.minmax_requant(w, 4, True, block=32)
    np.testing.assert_array_equal(o[1], O.pack_bits(4, ref["q"]))
    np.testing.assert_array_equal(o[2], ref["scale"])


def test_host_error_leaves_no_dangling_state(cuda):
  """A job that fails validation makes the call fail before anything is enqueued; a later call
  works and earlier outputs are untouched (the round-1 pipeline returned with chunks in flight
  and wrote stale staging into freed arrays on the next call)."""
  from aeq_b200 import _lib, host
  good = O.synthetic_weight(2100, 4096, 90)
  canary = np.full((2100, 4096), 77, np.int8)
  outs = [(canary, None, np.empty((2100, 1), np.float32), np.empty((2100, 1), np.int32)),
          (np.empty((3, 5), np.int8), np.empty(8, np.uint8), np.empty((3, 1), np.float32),
           np.empty((3, 1), np.int32))]
  with pytest.raises(_lib.AeqbError, match="straddle rows"):
    host.requant_rows([good, np.ones((3, 5), np.float32)], 4, True, outs=outs)  # 5 % 2 != 0
  assert (canary == 77).all(), "nothing may have been written when the call failed"
  with pytest.raises(_lib.AeqbError, match="rows longer than"):
    host.requant_rows([np.zeros((1, (32 << 20) // 4 + 128), np.float32)], 8, True)
  _check_rows([good], host.requant_rows([good], 8, True), 8, True, False)


def test_host_fan_out_over_devices(cuda):
  """aeqb_host_set_devices: chunks dealt round-robin over every visible GPU, same bytes out."""
  import torch
  from aeq_b200 import _lib, host
  ws = [O.synthetic_weight(2100, 4096, 80 + i) for i in range(3)]
  try:
    devs = host.set_devices("all")
    assert devs == list(range(torch.cuda.device_count()))
    _check_rows(ws, host.requant_rows(ws, 8, True), 8, True, False)
    with pytest.raises(_lib.AeqbError, match="not visible"):
      host.set_devices([torch.cuda.device_count() + 3])
  finally:
    host.set_devices(None)
  assert torch.cuda.current_device() == 0
  _check_rows(ws[:1], host.requant_rows(ws[:1], 8, True), 8, True, False)


def test_staged_copies_round_trip(cuda):
  """hostio.to_device / to_host above 1 MiB go through aeqb_host_copy_in / _out (pinned ring +
  worker threads): odd sizes, read-only and unaligned sources, several dtypes."""
  import torch
  from aeq_b200 import hostio
  rng = np.random.default_rng(3)
  a = rng.standard_normal(9_000_001, dtype=np.float32)          # 36 MB, not a multiple of the piece
  a.setflags(write=False)
  d = hostio.to_device(a)
  assert d.is_cuda and d.dtype == torch.float32
  np.testing.assert_array_equal(d.cpu().numpy(), a)
  np.testing.assert_array_equal(hostio.to_host(d), a)
  raw = np.empty(a.nbytes + 1, np.uint8)
  raw[1:] = a.view(np.uint8)
  un = raw[1:].view(np.float32)
  np.testing.assert_array_equal(hostio.to_host(hostio.to_device(un)), a)
  q = rng.integers(-128, 127, size=(3000, 1001), dtype=np.int8)
  np.testing.assert_array_equal(hostio.to_host(hostio.to_device(q) + 0), q)
  h = rng.standard_normal((600, 600))                             # float64 Hessian-sized
  np.testing.assert_array_equal(hostio.to_host(hostio.to_device(h)), h)
  # a producer still running on the current stream is waited for
  big = torch.ones(64 << 20, device=cuda)
  for _ in range(5):
    big = big * 1.0001
  got = hostio.to_host(big)
  assert got.shape == (64 << 20,) and np.all(got == got[0]) and got[0] > 1.0
