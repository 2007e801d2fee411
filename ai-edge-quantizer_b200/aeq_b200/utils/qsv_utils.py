"""QSV (quantisation-statistics) merge rules across calibration batches.

Mirror of ai_edge_quantizer/utils/qsv_utils.py:43-122.  These are O(1)-sized
host recurrences (a `(1,)*ndim` min/max pair per tensor per batch) that must be
evaluated in sample order in fp32 exactly like NumPy does, so they stay NumPy
expressions on purpose; the per-batch reductions that feed them are the device
kernels in `aeq_b200.calibration`.  The K x K Hessian merge of GPTQ is
device-side (`aeq_b200.device.hessian_merge`) when handed device tensors.
"""
from __future__ import annotations

import numpy as np

from .. import qtyping


def _ema(smoothing_factor, old, new):
  return smoothing_factor * old + (1.0 - smoothing_factor) * new


def moving_average_update(qsv: qtyping.QSV, new_qsv: qtyping.QSV,
                          smoothing_factor: float = 0.95) -> qtyping.QSV:
  """0.95 * old + 0.05 * new on min and max; first observation kept verbatim."""
  if not qsv:
    return new_qsv
  return {k: _ema(smoothing_factor, qsv[k], new_qsv[k]) for k in ("min", "max")}


def min_max_update(qsv: qtyping.QSV, new_qsv: qtyping.QSV) -> qtyping.QSV:
  """Union of ranges: elementwise min of mins, max of maxes."""
  if not qsv:
    return new_qsv
  return {"min": np.minimum(qsv["min"], new_qsv["min"]),
          "max": np.maximum(qsv["max"], new_qsv["max"])}


def gptq_and_moving_average_update(qsv: qtyping.QSV, new_qsv: qtyping.QSV) -> qtyping.QSV:
  """EMA on min/max plus the sample-weighted running mean of the Hessian."""
  if not qsv:
    return new_qsv
  out = moving_average_update(qsv, new_qsv)
  n_old, n_new = qsv["num_samples"], new_qsv["num_samples"]
  total = n_old + n_new
  if total == 0:
    out["hessian"], out["num_samples"] = new_qsv["hessian"], 0
  else:
    out["hessian"] = _merge_hessian(qsv["hessian"], n_old, new_qsv["hessian"], n_new, total)
    out["num_samples"] = total
  return out


def _oscar_merge_mu2(qsv1: qtyping.QSV, qsv2: qtyping.QSV):
  """Sample-weighted mean of the per-channel second moments (qsv_utils.py:125-157); O(channels)."""
  if "mu2" not in qsv1 and "mu2" not in qsv2:
    return None, 0
  if "mu2" not in qsv1:
    return qsv2.get("mu2"), qsv2.get("num_samples", 0)
  if "mu2" not in qsv2:
    return qsv1.get("mu2"), qsv1.get("num_samples", 0)
  n1, n2 = qsv1.get("num_samples", 0), qsv2.get("num_samples", 0)
  total = n1 + n2
  if total == 0:
    return qsv2["mu2"], 0
  return (qsv1["mu2"] * n1 + qsv2["mu2"] * n2) / total, total


def oscar_and_moving_average_update(qsv: qtyping.QSV, new_qsv: qtyping.QSV) -> qtyping.QSV:
  """EMA on min/max plus the merged mu2 (qsv_utils.py:160-171)."""
  if not qsv:
    return new_qsv
  out = moving_average_update(qsv, new_qsv)
  out["mu2"], out["num_samples"] = _oscar_merge_mu2(qsv, new_qsv)
  return out


def _merge_hessian(h_old, n_old, h_new, n_new, total):
  """(H_old * n_old + H_new * n_new) / total in float64 on the device (aeqb_hessian_merge_f64).

  NumPy in -> NumPy out; if either side is a device tensor the result stays on the device."""
  del total
  import torch
  from .. import device, hostio
  keep = isinstance(h_old, torch.Tensor) or isinstance(h_new, torch.Tensor)
  a = h_old if isinstance(h_old, torch.Tensor) else hostio.to_device(np.asarray(h_old, np.float64))
  b = h_new if isinstance(h_new, torch.Tensor) else hostio.to_device(np.asarray(h_new, np.float64))
  out = device.hessian_merge(a.double(), float(n_old), b.double(), float(n_new))
  return out if keep else hostio.to_host(out)
