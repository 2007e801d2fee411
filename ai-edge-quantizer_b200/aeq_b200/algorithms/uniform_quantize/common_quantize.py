"""Numeric tail of the reference's common_quantize.py, computed on the device.

Mirror of ai_edge_quantizer/algorithms/uniform_quantize/common_quantize.py:
`init_tensor_min_max` :1311-1359, `get_activation_min_max` :1362-1413,
`collect_activation_tensor_statistics` :1416-1456,
`get_tensor_indices_requiring_calibration` :1459-1494.  The ~50 per-op
`materialize_*` bookkeeping functions above line 1300 of that file are the
callers of this path, not part of it.
"""
from __future__ import annotations

from typing import MutableMapping, Optional, Sequence

import numpy as np

from ... import hostio
from ... import qtyping
from ...utils import tfl_flatbuffer_utils
from ..utils import common_utils
from . import uniform_quantize_tensor as uqt

_Gran = qtyping.QuantGranularity


def init_tensor_min_max(tensor_data: Optional[np.ndarray], op_info: qtyping.OpInfo) -> qtyping.QSV:
  """{min, max} of a weight at the configured granularity (empty dict if nothing to do)."""
  cfg = op_info.op_quant_config.weight_tensor_config
  if tensor_data is None or cfg is None:
    return {}
  from ... import device
  gran = cfg.granularity
  shape = tensor_data.shape
  if uqt.is_blockwise(gran):
    block = uqt.extract_block_size_from_granularity(gran)
    qdim = tfl_flatbuffer_utils.TFL_OP_TO_BLOCKWISE_WEIGHT_QUANTIZED_DIM[op_info.op_name]
    uqt._blockwise_shape(shape, qdim, block)
    if qdim != tensor_data.ndim - 1:
      raise ValueError("blockwise quantisation cuts the last axis")
    x = hostio.to_device(tensor_data.reshape(-1, shape[-1]), np.float32)
    mn, mx = device.minmax_blocks(x, block)
    out_shape = (*shape[:-1], shape[-1] // block)  # keepdims=False
  elif gran == _Gran.CHANNELWISE and (
      qdim := common_utils.get_weight_quantized_dim(op_info, tensor_data, gran)) is not None:
    x = device.channel_rows(hostio.to_device(tensor_data, np.float32), shape, qdim)
    mn, mx, _ = device.row_stats(x)
    out_shape = [1] * tensor_data.ndim
    out_shape[qdim] = shape[qdim]
  elif gran in (_Gran.TENSORWISE, _Gran.CHANNELWISE):
    mm = device.minmax_tensor(hostio.to_device(tensor_data.reshape(-1), np.float32))
    mn, mx = mm[0:1], mm[1:2]
    out_shape = [1] * tensor_data.ndim
  else:
    raise ValueError(f"Unsupported granularity: {gran}")
  return {"min": hostio.to_host(mn).reshape(out_shape),
          "max": hostio.to_host(mx).reshape(out_shape)}


def get_activation_min_max(tensor_content, valid_float_range_min: float | None = None,
                           valid_float_range_max: float | None = None) -> dict:
  """Scalar min over x > lo and max over x < hi (raw fallback), shaped (1,)*ndim.

  `tensor_content` may be a NumPy array or a device tensor (device-resident
  calibration batches skip the host round trip).
  """
  from ... import device
  ndim = tensor_content.ndim
  if isinstance(tensor_content, np.ndarray) and np.issubdtype(tensor_content.dtype, np.integer):
    # Integer activations (indices etc.): the reference takes plain min / max; they are
    # exactly representable in fp32 only below 2^24, so reduce them as int64 -> two
    # values on the host side of the boundary is not an option either: refuse loudly.
    if tensor_content.size and max(abs(int(tensor_content.min())), abs(int(tensor_content.max()))) >= 2**24:
      raise ValueError("integer activation values beyond 2^24 are not supported")
    x = hostio.to_device(tensor_content.astype(np.float32).reshape(-1))
    mm = hostio.to_host(device.minmax_tensor(x)).astype(tensor_content.dtype)
  else:
    x = hostio.to_device(tensor_content, np.float32).reshape(-1)
    mm = hostio.to_host(device.minmax_tensor(x, valid_float_range_min, valid_float_range_max))
  shape = (1,) * ndim
  return {"min": np.reshape(mm[0], shape), "max": np.reshape(mm[1], shape)}


def get_activation_min_max_batch(contents: Sequence, valid_float_range_min: float | None = None,
                                 valid_float_range_max: float | None = None) -> list[dict]:
  """`get_activation_min_max` for all float tensors of a calibration step in ONE device launch
  (per 64 tensors) and one download; integer tensors take the single-tensor path."""
  from ... import device
  out: list = [None] * len(contents)
  idx, dev_tensors = [], []
  for i, c in enumerate(contents):
    if isinstance(c, np.ndarray) and np.issubdtype(c.dtype, np.integer):
      out[i] = get_activation_min_max(c, valid_float_range_min, valid_float_range_max)
    else:
      idx.append(i)
      dev_tensors.append(hostio.to_device(c, np.float32).reshape(-1))
  if idx:
    mm = hostio.to_host(device.minmax_tensors(dev_tensors, valid_float_range_min,
                                              valid_float_range_max))
    for k, i in enumerate(idx):
      shape = (1,) * contents[i].ndim
      out[i] = {"min": np.reshape(mm[k, 0], shape), "max": np.reshape(mm[k, 1], shape)}
  return out


def collect_activation_statistics_batch(tensor_indices: Sequence[int], graph_info: qtyping.GraphInfo,
                                        tensor_content_map: MutableMapping,
                                        valid_float_range_min: float | None = None,
                                        valid_float_range_max: float | None = None) -> list:
  """[(name, content, {min, max, num_samples})] for the runtime tensors among `tensor_indices`
  (constants are skipped), reduced in one batched launch."""
  names, contents = [], []
  for tid in tensor_indices:
    tensor = graph_info.subgraph_tensors[tid]
    if tfl_flatbuffer_utils.get_tensor_data(tensor, graph_info.buffers) is not None:
      continue
    name = tfl_flatbuffer_utils.get_tensor_name(tensor)
    names.append(name)
    contents.append(tensor_content_map[name])
  qsvs = get_activation_min_max_batch(contents, valid_float_range_min, valid_float_range_max)
  for c, q in zip(contents, qsvs):
    q["num_samples"] = np.array(c.shape[0] if c.ndim > 0 else 1)
  return list(zip(names, contents, qsvs))


def collect_activation_tensor_statistics(tensor_idx: int, graph_info: qtyping.GraphInfo,
                                         tensor_content_map: MutableMapping,
                                         valid_float_range_min: float | None = None,
                                         valid_float_range_max: float | None = None):
  """(name, content, {min, max, num_samples}) of a runtime tensor; None for constants."""
  tensor = graph_info.subgraph_tensors[tensor_idx]
  if tfl_flatbuffer_utils.get_tensor_data(tensor, graph_info.buffers) is not None:
    return None
  name = tfl_flatbuffer_utils.get_tensor_name(tensor)
  content = tensor_content_map[name]
  qsv = get_activation_min_max(content, valid_float_range_min, valid_float_range_max)
  qsv["num_samples"] = np.array(content.shape[0] if content.ndim > 0 else 1)
  return name, content, qsv


def check_if_quantized(tensor) -> bool:
  q = getattr(tensor, "quantization", None)
  return q is not None and getattr(q, "scale", None) is not None


def get_tensor_indices_requiring_calibration(tfl_op, graph_info: qtyping.GraphInfo,
                                             inputs_to_ignore: Sequence[int] | None = None,
                                             outputs_to_ignore: Sequence[int] | None = None) -> list[int]:
  """Tensor ids of the op's runtime inputs / outputs that still need statistics."""
  skip_in = set(inputs_to_ignore or [])
  skip_in.update(k for k, tid in enumerate(tfl_op.inputs)
                 if check_if_quantized(graph_info.subgraph_tensors[tid]))
  skip_out = set(outputs_to_ignore or [])
  return ([tid for k, tid in enumerate(tfl_op.inputs) if k not in skip_in and tid != -1]
          + [tid for k, tid in enumerate(tfl_op.outputs) if k not in skip_out and tid != -1])
