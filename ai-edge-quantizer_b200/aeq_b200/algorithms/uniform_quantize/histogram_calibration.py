"""`histogram_min_max_uniform_quantize`: device histogram calibration behind the plug-in API.

The reference ships the histogram class (ai_edge_quantizer/utils/histogram_utils.py:274-480,
`DynamicHistogram.add` :396-416, `.merge` :418-452) but binds it to no algorithm; SURVEY.md §8(f)
row 4 asks for it to be wired into a calibration func.  This module is that binding, in the
reference's own three-callable shape (algorithm_manager_api.py:51-122):

  calibration_func  `histogram_calibrate`   min / max exactly as `min_max_calibrate`
                    (naive_min_max_quantize.py:181-226: one batched `aeqb_minmax_tensors_f32`
                    launch per op) plus one `DynamicHistogram.add` per runtime tensor — the bin
                    count is `aeqb_hist_accumulate_f32` on the same device copy, so a batch is
                    uploaded once;
  update_qsv_func   `histogram_update`      histograms merge (resampling rule of the reference's
                    `_DynamicHistogram1D.merge`), min / max follow the moving average of
                    `qsv_utils.moving_average_update` (:43-68) so that with `percentile = 100`
                    the algorithm is indistinguishable from min_max_uniform_quantize;
  materialize       `get_tensor_quant_params`  weights: the fused min-max kernels; activations:
                    the (lo, hi) range that keeps `percentile` % of the histogram mass on each
                    side-trimmed tail, then `tensor_zp_scale_from_min_max` as min-max does.

Counts are integers and the range bookkeeping is the reference's arithmetic, so the histogram a
calibration run ends with equals the one the reference class builds from the same batches
(tests/test_gpu_histogram.py).
"""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from ... import hostio
from ... import qtyping
from ...utils import histogram_utils
from ...utils import qsv_utils
from . import common_quantize
from . import naive_min_max_quantize

ALGORITHM_KEY = "histogram_min_max_uniform_quantize"
HISTOGRAM_KEY = "histogram"
DEFAULT_MAX_BINS = 2048
# Fraction (in %) of the calibration mass the quantised range must cover; 100 keeps the
# calibrated min / max untouched.  Module-level like the reference's algorithm constants
# (octav.py:55-64, mse.py:30-33); `set_percentile` changes it for a run.
_PERCENTILE = 100.0


def set_percentile(p: float) -> None:
  global _PERCENTILE
  if not 50.0 < p <= 100.0:
    raise ValueError(f"percentile must be in (50, 100], got {p}")
  _PERCENTILE = float(p)


def get_percentile() -> float:
  return _PERCENTILE


def histogram_calibrate(tfl_op, graph_info: qtyping.GraphInfo, tensor_content_map,
                        inputs_to_ignore=None, outputs_to_ignore=None,
                        valid_range: tuple[float, float] = (-3e38, 3e38),
                        max_tensor_bins: int = DEFAULT_MAX_BINS, **kwargs) -> dict:
  """{tensor name: {min, max, num_samples, histogram}} for every runtime tensor of the op."""
  del kwargs
  ids = common_quantize.get_tensor_indices_requiring_calibration(
      tfl_op, graph_info, inputs_to_ignore, outputs_to_ignore)
  # every float batch is uploaded ONCE: the min / max launch and the bin count read the same copy
  resident = _ResidentMap(tensor_content_map)
  out = {}
  for name, content, qsv in common_quantize.collect_activation_statistics_batch(
      ids, graph_info, resident, valid_range[0], valid_range[1]):
    if isinstance(content, np.ndarray):
      out[name] = qsv  # integer index tensors: plain min / max, nothing to bin
      continue
    h = histogram_utils.DynamicHistogram(max_tensor_bins=max_tensor_bins)
    h.add(content)
    qsv[HISTOGRAM_KEY] = h.to_dict()
    out[name] = qsv
  return out


class _ResidentMap(dict):
  """tensor_content_map view whose float arrays are device tensors (uploaded on first use)."""

  def __init__(self, source):
    super().__init__()
    self._source = source

  def __missing__(self, name):
    c = self._source[name]
    if not (isinstance(c, np.ndarray) and np.issubdtype(c.dtype, np.integer)):
      c = hostio.to_device(c, np.float32)
    self[name] = c
    return c


def histogram_update(qsv: qtyping.QSV, new_qsv: qtyping.QSV, smoothing_factor: float = 0.95,
                     max_tensor_bins: int = DEFAULT_MAX_BINS) -> qtyping.QSV:
  """Moving-average min / max (qsv_utils.py:43-68) + merged histograms."""
  merged = qsv_utils.moving_average_update(
      {k: v for k, v in qsv.items() if k != HISTOGRAM_KEY},
      {k: v for k, v in new_qsv.items() if k != HISTOGRAM_KEY}, smoothing_factor)
  old_h, new_h = qsv.get(HISTOGRAM_KEY), new_qsv.get(HISTOGRAM_KEY)
  if old_h or new_h:
    h = histogram_utils.DynamicHistogram.from_dict(old_h or {}, max_tensor_bins)
    h.merge(histogram_utils.DynamicHistogram.from_dict(new_h or {}, max_tensor_bins))
    merged[HISTOGRAM_KEY] = h.to_dict()
  return merged


def percentile_range(hist: dict, percentile: float) -> Optional[tuple[np.float32, np.float32]]:
  """(lo, hi) bin edges that leave (100 - percentile) / 2 % of the mass outside on each side;
  None for an empty histogram.  Edges, not bin centres: the kept range never cuts into a bin
  that holds kept mass, so percentile = 100 returns a range that contains min and max."""
  if not hist or "channels" not in hist or not hist["channels"] or not hist["channels"][0]:
    return None
  ch = hist["channels"][0]
  counts = np.asarray(ch["hist_counts"], dtype=np.int64)
  total = int(counts.sum())
  if total == 0:
    return None
  lb, bw = float(ch["lower_bound"]), float(ch["bin_width"])
  tail = total * (100.0 - percentile) / 200.0
  cum = np.cumsum(counts)
  lo_bin = int(np.searchsorted(cum, tail, side="right"))          # first bin whose cum count > tail
  rcum = np.cumsum(counts[::-1])
  hi_bin = len(counts) - 1 - int(np.searchsorted(rcum, tail, side="right"))
  lo_bin = min(lo_bin, len(counts) - 1)
  hi_bin = max(hi_bin, lo_bin)
  return np.float32(lb + lo_bin * bw), np.float32(lb + (hi_bin + 1) * bw)


def get_tensor_quant_params(op_info: qtyping.OpInfo,
                            tensor_quant_config: qtyping.TensorQuantizationConfig,
                            tensor_content: Optional[np.ndarray] = None,
                            tensor_qsv: Optional[dict[str, Any]] = None) -> qtyping.UniformQuantParams:
  """Min-max materialisation on the percentile range of the calibrated histogram."""
  if tensor_qsv is not None and HISTOGRAM_KEY in tensor_qsv and "min" in tensor_qsv:
    qsv = {k: v for k, v in tensor_qsv.items() if k != HISTOGRAM_KEY}
    if _PERCENTILE < 100.0:
      r = percentile_range(tensor_qsv[HISTOGRAM_KEY], _PERCENTILE)
      if r is not None:
        # never wider than what calibration saw (the histogram's outer bins are padded)
        qsv["min"] = np.maximum(np.asarray(qsv["min"], np.float32), r[0])
        qsv["max"] = np.minimum(np.asarray(qsv["max"], np.float32), r[1])
    tensor_qsv = qsv
  return naive_min_max_quantize.get_tensor_quant_params(op_info, tensor_quant_config, tensor_content,
                                                        tensor_qsv)
