"""The stage schedule of `hadamard_tiles<LOG2N>` (csrc/hadamard.cu), replayed in NumPy fp32.

A warp owns a 4096-float tile.  In the load map (float4 slot j*32 + lane) the registers of a lane
hold bits e[1:0] and e[11:7] of the element index, so the kernel runs the butterfly stages of bits
0, 1 and 7..log2(n)-1 there; after the one exchange (float4 slot lane*32 + k) the registers hold
e[6:0] and it runs bits 2..6, then multiplies by fl(1/sqrt(n)).  This test checks that the two
stage sets cover every bit of every segment length the kernel is launched for (n = 256..4096),
that stages may be applied in that order (Sylvester H_n is a Kronecker product of 2x2 blocks: the
bit stages commute), and that the fp32 result stays within the tolerance the GPU test uses against
the reference's dense product (hadamard_rotation.py:93-134).
"""
import numpy as np
import pytest

from oracle import aeq_oracle as O

TILE = 4096


def _stage(x, bit):
  h = 1 << bit
  v = x.reshape(-1, 2, h)                     # [..., pair member, offset inside half]
  a, b = v[:, 0, :].copy(), v[:, 1, :].copy()
  v[:, 0, :] = a + b
  v[:, 1, :] = a - b


def tile_schedule(tile, log2n):
  x = tile.astype(np.float32).copy()
  map1 = [0, 1] + list(range(7, log2n))       # registers hold e[1:0], e[11:7]
  map2 = [2, 3, 4, 5, 6]                      # after the exchange: e[6:0]
  assert sorted(map1 + map2) == list(range(log2n)), "every bit of the segment index has a stage"
  assert all(b in (0, 1) or 7 <= b <= 11 for b in map1) and all(2 <= b <= 6 for b in map2)
  for bit in map1 + map2:
    _stage(x, bit)
  return x * np.float32(1.0 / np.sqrt(np.float32(1 << log2n)))


@pytest.mark.parametrize("log2n", [8, 9, 10, 11, 12])
def test_schedule_matches_dense_product(log2n):
  n = 1 << log2n
  w = O.synthetic_weight(2, TILE, index=log2n).reshape(-1)         # two tiles
  got = np.concatenate([tile_schedule(w[i:i + TILE], log2n) for i in range(0, w.size, TILE)])
  want = np.matmul(w.reshape(-1, n), O.hadamard_matrix(n)).reshape(-1)
  assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
  # an involution: rotating twice returns the input
  back = np.concatenate([tile_schedule(got[i:i + TILE], log2n) for i in range(0, w.size, TILE)])
  assert np.abs(back - w).max() <= 4e-6 * np.abs(w).max()
