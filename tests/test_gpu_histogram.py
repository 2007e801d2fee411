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
