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
