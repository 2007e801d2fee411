"""Dynamic-range histogram calibration with the per-element work on the device.

Mirror of ai_edge_quantizer/utils/histogram_utils.py (`_DynamicHistogram1D` :24-271,
`DynamicHistogram` :274-480), same attributes, dictionaries and merge rules.  The two passes
over the data in `add` — finite min / max and the bin count — are the device kernels
`aeqb_minmax_tensors_f32` (filter (-inf, +inf) = np.isfinite) and `aeqb_hist_accumulate_f32`;
everything that touches only the <= max_bins counters (range growth, bin-width doubling and
compaction, resampled merges) is bookkeeping on a few KiB and stays on the host, exactly as
the reference orders it.  Data may be NumPy arrays or CUDA tensors.
"""
from __future__ import annotations

import math
from typing import Any, Mapping, Optional, Sequence

import numpy as np

from .. import hostio


def _finite_min_max_device(x_dev):
  """(min, max) over the finite elements of a device tensor as np.float32, or None if none."""
  from .. import device
  mm = hostio.to_host(device.minmax_tensors([x_dev.reshape(-1)], -math.inf, math.inf))[0]
  if not (np.isfinite(mm[0]) and np.isfinite(mm[1])):
    return None
  return np.float32(mm[0]), np.float32(mm[1])


class _DynamicHistogram1D:
  """A 1-D histogram whose range grows with the data while the bin count stays <= max_bins."""

  def __init__(self, max_bins: int = 2048, initial_bin_width: Optional[float] = None):
    self.max_bins = max_bins
    self.bin_width = initial_bin_width
    self.counts = np.zeros(1, dtype=np.int64)
    self.lower_bound = 0.0
    self.initialized = False
    self.global_min = float("inf")
    self.global_max = float("-inf")

  # ------------------------------------------------------------------ (de)serialisation
  @classmethod
  def from_dict(cls, d: Mapping[str, Any], max_bins: int = 2048) -> "_DynamicHistogram1D":
    obj = cls(max_bins=max_bins)
    if "hist_counts" in d:
      obj.counts = np.array(d["hist_counts"], dtype=np.int64)
      obj.lower_bound, obj.bin_width, obj.initialized = d["lower_bound"], d["bin_width"], True
      first = lambda v: v[0] if isinstance(v, (np.ndarray, list, tuple)) else v
      obj.global_min, obj.global_max = first(d["min"]), first(d["max"])
    return obj

  def to_dict(self) -> dict[str, Any]:
    if not self.initialized:
      return {}
    return {"hist_counts": self.counts, "bin_edges": self.bin_edges, "bin_width": self.bin_width,
            "lower_bound": self.lower_bound, "min": np.array([self.global_min]),
            "max": np.array([self.global_max])}

  @property
  def bin_edges(self) -> np.ndarray:
    if not self.initialized:
      return np.array([self.lower_bound])
    return self.lower_bound + np.arange(len(self.counts) + 1) * self.bin_width

  # ------------------------------------------------------------------ range bookkeeping (host)
  def _initialize(self, d_min, d_max) -> None:
    if self.bin_width is None:
      span = d_max - d_min
      pad = span * 0.1 if span > 0 else 1e-4  # 10 % margin, or a token width for constant data
      lo, hi = d_min - pad, d_max + pad
      self.lower_bound = lo
      self.bin_width = max((hi - lo) / self.max_bins, 1e-5)
      nbins = self.max_bins
    else:
      self.lower_bound = d_min
      nbins = max(int(np.ceil((d_max - d_min) / self.bin_width)), 1)
    self.counts = np.zeros(nbins, dtype=np.int64)
    self.initialized = True

  def _double_bin_width_and_compact(self) -> None:
    if len(self.counts) % 2:
      self.counts = np.pad(self.counts, (0, 1), "constant", constant_values=0)
    self.counts = self.counts.reshape(-1, 2).sum(axis=1)
    self.bin_width *= 2.0

  def _expand_to_fit(self, d_min, d_max) -> None:
    if d_min < self.lower_bound:
      need = lambda: int(np.ceil((self.lower_bound - d_min) / self.bin_width))
      extra = need()
      while len(self.counts) + extra > self.max_bins:
        self._double_bin_width_and_compact()
        extra = need()
      self.counts = np.pad(self.counts, (extra, 0), "constant", constant_values=0)
      self.lower_bound -= extra * self.bin_width
    upper = self.lower_bound + len(self.counts) * self.bin_width
    if d_max > upper:
      def need():
        top = self.lower_bound + len(self.counts) * self.bin_width
        return int(np.ceil((d_max - top) / self.bin_width))
      extra = need()
      while len(self.counts) + extra > self.max_bins:
        self._double_bin_width_and_compact()
        extra = need()
      self.counts = np.pad(self.counts, (0, extra), "constant", constant_values=0)

  # ------------------------------------------------------------------ data path (device)
  def add(self, data) -> None:
    """Adds the finite elements of `data` (NumPy array or CUDA tensor)."""
    if (data.numel() if hasattr(data, "numel") else data.size) == 0:
      return
    x = hostio.to_device(data, np.float32).reshape(-1)
    mm = _finite_min_max_device(x)
    if mm is None:
      return
    self._add_device(x, mm[0], mm[1])

  def _add_device(self, x_dev, d_min, d_max) -> None:
    from .. import device
    self.global_min = min(self.global_min, d_min)
    self.global_max = max(self.global_max, d_max)
    if not self.initialized:
      self._initialize(d_min, d_max)
    self._expand_to_fit(d_min, d_max)
    delta = device.hist_accumulate(x_dev, float(np.float32(self.lower_bound)),
                                   float(np.float32(self.bin_width)), len(self.counts))
    self.counts = self.counts + hostio.to_host(delta)

  # ------------------------------------------------------------------ merge (host, <= max_bins)
  def _accumulate_resampled(self, other: "_DynamicHistogram1D") -> None:
    acc = self.counts.astype(np.float64)
    for i in np.flatnonzero(other.counts):
      c = other.counts[i]
      left = other.lower_bound + i * other.bin_width
      right = left + other.bin_width
      first = max(0, int(np.floor((left - self.lower_bound) / self.bin_width)))
      last = min(len(self.counts), int(np.ceil((right - self.lower_bound) / self.bin_width)))
      for j in range(first, last):
        lo = self.lower_bound + j * self.bin_width
        a, b = max(left, lo), min(right, lo + self.bin_width)
        if a < b:
          acc[j] += c * ((b - a) / other.bin_width)
    self.counts = np.round(acc).astype(np.int64)

  def merge(self, other: "_DynamicHistogram1D") -> None:
    self.global_min = min(self.global_min, other.global_min)
    self.global_max = max(self.global_max, other.global_max)
    if not other.initialized:
      return
    if not self.initialized:
      self.bin_width, self.counts = other.bin_width, np.copy(other.counts)
      self.lower_bound, self.initialized = other.lower_bound, True
      return
    while self.bin_width < other.bin_width:
      self._double_bin_width_and_compact()
    self._expand_to_fit(other.lower_bound, other.lower_bound + len(other.counts) * other.bin_width)
    self._accumulate_resampled(other)


class DynamicHistogram:
  """Per-tensor (axis=None) or per-channel (axis=k) set of `_DynamicHistogram1D`."""

  def __init__(self, max_tensor_bins: int = 2048, initial_bin_width: Optional[float] = None,
               axis: Optional[int] = None):
    self.initial_bin_width = initial_bin_width
    self.max_tensor_bins = max_tensor_bins
    self.axis = axis
    self._impls: Optional[Sequence[_DynamicHistogram1D]] = None

  @property
  def initialized(self) -> bool:
    if self._impls is None:
      return False
    return self._impls[0].initialized if self.axis is None else True

  @property
  def global_min(self) -> np.ndarray:
    if self._impls is None:
      return np.array([float("inf")]) if self.axis is None else np.array([])
    return np.array([h.global_min for h in self._impls])

  @property
  def global_max(self) -> np.ndarray:
    if self._impls is None:
      return np.array([float("-inf")]) if self.axis is None else np.array([])
    return np.array([h.global_max for h in self._impls])

  def _per_tensor(self, what: str, default):
    if self.axis is not None:
      raise AttributeError(
          f"{what} is not supported for per-channel histogram, use _impls[i].{what}")
    return default if self._impls is None else getattr(self._impls[0], what)

  @property
  def counts(self) -> np.ndarray:
    return self._per_tensor("counts", np.zeros(1, dtype=np.int64))

  @property
  def bin_width(self):
    return self._per_tensor("bin_width", None)

  @property
  def lower_bound(self) -> float:
    return self._per_tensor("lower_bound", 0.0)

  def _make_impls(self, n: int, bins: int):
    self._impls = [_DynamicHistogram1D(max_bins=bins, initial_bin_width=self.initial_bin_width)
                   for _ in range(n)]

  def add(self, data) -> None:
    size = data.numel() if hasattr(data, "numel") else data.size
    if size == 0:
      return
    if self._impls is None:
      if self.axis is None:
        self._make_impls(1, self.max_tensor_bins)
      else:
        ch = data.shape[self.axis]
        self._make_impls(ch, max(self.max_tensor_bins // ch, 1))
    x = hostio.to_device(data, np.float32)
    if self.axis is None:
      self._impls[0].add(x)
      return
    from .. import device
    chans = x.movedim(self.axis, 0).contiguous().reshape(len(self._impls), -1)
    # one batched launch for every channel's finite min / max, then one bin-count per channel
    mm = hostio.to_host(device.minmax_tensors(list(chans), -math.inf, math.inf))
    for i, h in enumerate(self._impls):
      if np.isfinite(mm[i, 0]) and np.isfinite(mm[i, 1]):
        h._add_device(chans[i], np.float32(mm[i, 0]), np.float32(mm[i, 1]))  # pylint: disable=protected-access

  def merge(self, other: "DynamicHistogram") -> None:
    if self.axis != other.axis:
      raise ValueError(
          f"Cannot merge histograms with different axis: {self.axis} vs {other.axis}")
    if self._impls is None and other._impls is not None:
      self._make_impls(len(other._impls), other._impls[0].max_bins)
    if self._impls is None or other._impls is None:
      return
    if len(self._impls) != len(other._impls):
      raise ValueError(
          "Cannot merge: different number of channels:"
          f" {len(self._impls)} vs {len(other._impls)}")
    for mine, theirs in zip(self._impls, other._impls):
      mine.merge(theirs)

  def to_dict(self) -> dict[str, Any]:
    if not self.initialized:
      return {}
    return {"min": self.global_min, "max": self.global_max, "axis": self.axis,
            "channels": [h.to_dict() for h in self._impls]}

  @classmethod
  def from_dict(cls, d: Mapping[str, Any], max_tensor_bins: int = 2048) -> "DynamicHistogram":
    if not d:
      return cls(max_tensor_bins=max_tensor_bins)
    if "channels" not in d:
      raise ValueError(f"Invalid dictionary format for DynamicHistogram: {d}")
    obj = cls(max_tensor_bins=max_tensor_bins, axis=d["axis"])
    bins = max(max_tensor_bins // len(d["channels"]), 1)
    obj._impls = [_DynamicHistogram1D.from_dict(h, max_bins=bins) for h in d["channels"]]
    return obj
