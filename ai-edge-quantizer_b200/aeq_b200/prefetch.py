"""Batched weight driver: requantise every constant weight of a model in a few launches and
pre-populate the `(buffer, config)` cache, so that the per-op `materialize` calls of
`params_generator.generate_quantization_parameters` (params_generator.py:110-183) become cache
hits (`common_utils.py:260-264`) instead of one device round trip per op.

SURVEY.md §8(f) row 1.  The reference walks ops sequentially and quantises each weight when it
meets it; nothing in that walk depends on the ORDER in which weights are quantised (the only
shared state is the cache), so all min-max weights of one (granularity, bits, symmetry) group go
through ONE `aeqb_host_requant_{rows,blocks}_batch_f32` call: chunked, pipelined H2D -> fused
kernel -> D2H over all of them.  Works with this package's data classes or, through
`make_params`, with the reference's.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import numpy as np

from . import host
from . import qtyping
from .algorithms.uniform_quantize import uniform_quantize_tensor as uqt
from .transformations import quantize_tensor
from .utils import tfl_flatbuffer_utils

# op name -> quantised dimension of its weight when CHANNELWISE (dim 0 is the only layout the
# fused row kernel covers; DEPTHWISE_CONV_2D quantises dim 3 and stays on the per-op path).
_DIM0_OPS = ("FULLY_CONNECTED", "CONV_2D", "EMBEDDING_LOOKUP", "CONV_2D_TRANSPOSE")
_BLOCKWISE_OPS = ("FULLY_CONNECTED", "EMBEDDING_LOOKUP")
_IGNORED_INPUTS = {"CONV_2D_TRANSPOSE": (0,), "EMBEDDING_LOOKUP": (0,)}


def _name(op_name) -> str:
  return getattr(op_name, "value", str(op_name))


def _default_params(**kw):
  return qtyping.UniformQuantParams(**kw)


MSE_MULTIPLIERS = {8: 0.05408, 4: 0.37755}  # mse.py:30-33


def prefetch_weights(items: Iterable, cache, make_params: Optional[Callable] = None,
                     get_tensor_data: Optional[Callable] = None,
                     algorithm: str = "min_max_uniform_quantize") -> dict:
  """Quantises the constant float32 weights of `items` and inserts them into `cache`.

  items:   iterable of (op_info, graph_info) for ops whose algorithm is `algorithm`: min-max uniform
           quantisation (default) or "MSE" (per-channel symmetric INT8 / INT4 with 128-multiple rows;
           everything else is left to the per-op path)
           (`op_info.op_quant_config.weight_tensor_config` is the tensor config).
  cache:   a TensorQuantParamsCache (ours or the reference's: `lookup` / `insert`).
  make_params: keyword factory for the UniformQuantParams class to emit (default: ours).
  Returns counters: tensors quantised, cache hits skipped, tensors left to the per-op path.
  """
  make_params = make_params or _default_params
  get_tensor_data = get_tensor_data or tfl_flatbuffer_utils.get_tensor_data
  groups: dict = {}
  stats = {"quantized": 0, "already_cached": 0, "left_to_per_op_path": 0, "batched_calls": 0}
  seen = set()
  for op_info, graph_info in items:
    cfg = op_info.op_quant_config.weight_tensor_config
    op = _name(op_info.op_name)
    if cfg is None or op not in _DIM0_OPS:
      continue
    min_elems = getattr(op_info.op_quant_config, "min_weight_elements", 0)
    for pos, tid in enumerate(op_info.op.inputs):
      if tid == -1 or pos in _IGNORED_INPUTS.get(op, ()):
        continue
      tensor = graph_info.subgraph_tensors[tid]
      data = get_tensor_data(tensor, graph_info.buffers)
      if data is None or data.dtype != np.float32 or data.size < min_elems or data.ndim < 2:
        continue  # runtime tensor, bias (1-D: quantised from the scales later) or tiny weight
      key = (tensor.buffer, cfg)
      if key in seen:
        continue
      seen.add(key)
      if cache.lookup(tensor.buffer, cfg) is not None:
        stats["already_cached"] += 1
        continue
      block = uqt.extract_block_size_from_granularity(_gran(cfg))
      gran = _name(cfg.granularity)
      if algorithm == "MSE":
        cols = int(np.prod(data.shape[1:]))
        if (block or gran != "CHANNELWISE" or not cfg.symmetric or cfg.num_bits not in MSE_MULTIPLIERS
            or cols % 128 or cols * 4 > 98304):
          stats["left_to_per_op_path"] += 1  # the per-op call handles (or rejects) these
          continue
        groups.setdefault(("mse", cfg.num_bits), []).append((tensor.buffer, cfg, data))
        continue
      if block:
        if op not in _BLOCKWISE_OPS or not cfg.symmetric or data.shape[-1] % block:
          stats["left_to_per_op_path"] += 1  # the per-op call raises the reference's error
          continue
        gkey = ("blocks", block, cfg.num_bits)
      elif gran == "CHANNELWISE" and cfg.num_bits in (2, 4, 8):
        gkey = ("rows", bool(cfg.symmetric), cfg.num_bits)
      else:
        stats["left_to_per_op_path"] += 1  # TENSORWISE etc.: single-tensor kernels
        continue
      groups.setdefault(gkey, []).append((tensor.buffer, cfg, data))
  for gkey, members in groups.items():
    if gkey[0] == "mse":
      bits = gkey[1]
      ws = [d.reshape(d.shape[0], -1) for _, _, d in members]
      outs = host.requant_mse_rows(ws, bits, MSE_MULTIPLIERS[bits])
      for (buf, cfg, d), (q, _, scale, _zp) in zip(members, outs):
        pshape = [d.shape[0]] + [1] * (d.ndim - 1)
        scale = scale.reshape(pshape)
        cache.insert(buf, cfg, make_params(
            num_bits=bits, quantized_dimension=0, scale=scale,
            zero_point=np.zeros_like(scale, dtype=np.int32), symmetric=True,  # mse.py:109
            quantized_data=q.reshape(d.shape), block_size=0))
    elif gkey[0] == "rows":
      _, sym, bits = gkey
      ws = [d.reshape(d.shape[0], -1) for _, _, d in members]
      fuse_pack = bits in (2, 4) and all(w.shape[1] % (8 // bits) == 0 for w in ws)
      outs = host.requant_rows(ws, bits, sym, want_packed=fuse_pack)
      for (buf, cfg, d), (q, packed, scale, zp) in zip(members, outs):
        if packed is not None:
          quantize_tensor.remember_packed(q, bits, packed)
        pshape = [d.shape[0]] + [1] * (d.ndim - 1)
        cache.insert(buf, cfg, make_params(
            num_bits=bits, quantized_dimension=0, scale=scale.reshape(pshape),
            zero_point=zp.reshape(pshape).astype(uqt.numpy_dtype_for(bits)), symmetric=sym,
            quantized_data=q.reshape(d.shape), block_size=0))
    else:
      _, block, bits = gkey
      ws = [d.reshape(-1, d.shape[-1]) for _, _, d in members]
      outs = host.requant_blocks(ws, block, bits, want_packed=(bits == 4))
      for (buf, cfg, d), (q, packed, scale, _) in zip(members, outs):
        if packed is not None:
          quantize_tensor.remember_packed(q, bits, packed)
        sshape = (*d.shape[:-1], d.shape[-1] // block)
        cache.insert(buf, cfg, make_params(
            num_bits=bits, quantized_dimension=d.ndim - 1, scale=scale.reshape(sshape),
            zero_point=np.zeros(sshape, dtype=uqt.numpy_dtype_for(bits)), symmetric=True,
            quantized_data=q.reshape(d.shape), block_size=block))
    stats["quantized"] += len(members)
    stats["batched_calls"] += 1
  return stats


def _gran(cfg):
  """Our QuantGranularity for a config of either package (enum values are identical)."""
  return qtyping.QuantGranularity(_name(cfg.granularity))
