"""`HADAMARD_GPTQ`: block-diagonal Hadamard rotation composed with GPTQ (BASELINE.json configs[4]).

The reference ships the two halves as separate algorithms — `hadamard_rotation`
(ai_edge_quantizer/algorithms/uniform_quantize/hadamard_rotation.py:93-203: rotate the weight's
last axis with R = diag(H_n / sqrt(n), ...), OCTAV on the result, the runtime rotates the
activation with the same R) and `gptq` (gptq.py:55-300: Hessian of the op's input, damped inverse,
OBS column loop).  Composing them is arithmetic only: the rotated op sees the activation x R, so
its Hessian is (2 / n_s) (X R)^T (X R) = R^T H R, and GPTQ runs on the pair (W R, R^T H R) with the
min/max scales of the ROTATED weight (gptq.py:257-271).  On the device:

  W R        one pass of the Hadamard tile kernel            (aeqb_hadamard_rows_f32)
  R^T H R    H is symmetric: rotate its rows, transpose, rotate the rows again — two more passes of
             the same kernel and one aeqb_swap_axes            (fp32, like the weight)
  GPTQ       aeqb_hessian_inverse_f64 + aeqb_gptq_quantize_f32 on the rotated pair

Results carry `HadamardRotationParams`, so the reference's materialisers insert the activation-side
rotation exactly as they do for HADAMARD_ROTATION.
"""
from __future__ import annotations

from typing import Any, Mapping, Optional

import numpy as np

from ... import hostio
from ... import qtyping
from . import gptq
from . import hadamard_rotation
from . import uniform_quantize_tensor as uqt

ALGORITHM_KEY = "HADAMARD_GPTQ"


def rotate_hessian_device(hessian, hadamard_size: int):
  """float64 [K, K] device tensor R^T H R for the block-diagonal R of size `hadamard_size`."""
  import torch
  from ... import device
  h = hessian if isinstance(hessian, torch.Tensor) else hostio.to_device(np.asarray(hessian, np.float64))
  k = h.shape[0]
  if hadamard_size <= 1:
    return h.double().clone()
  h32 = h.float().contiguous()
  a = device.hadamard_rows(h32, hadamard_size)                 # H R
  at = device.swap_axes(a, k, k, 1).reshape(k, k)              # (H R)^T = R^T H   (H symmetric)
  return device.hadamard_rows(at, hadamard_size).double()      # R^T H R


def quantize_device(w_dev, hessian, num_bits: int, symmetric: bool = True,
                    max_hadamard_size: Optional[int] = None, damp: float = 0.01):
  """Device-resident pipeline: (q int8 [R, K], scale [R, 1], zero_point, hadamard_size).  `hessian`
  is the UNROTATED float64 Hessian of the op's input (host array or device tensor); it is not
  modified."""
  from ... import device
  rows, k = w_dev.shape
  rot, n = hadamard_rotation.rotate_with_diagonal_hadamard_device(w_dev, (rows, k), max_hadamard_size)
  h_rot = rotate_hessian_device(hessian, n)
  hinv = device.hessian_inverse(h_rot, damp)
  mn, mx, _ = device.row_stats(rot)
  zp, scale, _ = device.scale_zp_from_minmax(mn, mx, num_bits, symmetric, False)
  q = device.gptq_quantize(rot, hinv, scale.reshape(-1), zp.reshape(-1), 0, num_bits, symmetric)
  return q, scale, zp, n


def quantize_layer_device(weights, feeds, hessians, num_bits: int, symmetric: bool = True,
                          max_hadamard_size: Optional[int] = None, damp: float = 0.01,
                          concurrent: bool = False):
  """Rotated GPTQ of all the FC weights of one layer: [(q, scale, zero_point, hadamard_size)].

  weights:  device float32 [R_i, K_i] matrices (e.g. q, k, v, o, gate, up, down of a decoder layer);
  feeds:    for each weight, the key of the Hessian of ITS input in `hessians`;
  hessians: {key: UNROTATED float64 [K, K] device Hessian} (shared by the weights it feeds).
  The weights of a layer are independent problems given their Hessians (the reference walks them
  one op at a time, params_generator.py:110-183, with no state in between), and both halves of the
  work are chains of short dependent launches — the factorisation's diagonal blocks, the OBS
  loop's 64-column steps — that leave most of the GPU idle.  So every Hessian is rotated, inverted
  and every OBS loop runs on its OWN stream: the inverses overlap each other, an OBS loop starts
  as soon as its inverse is there.  Same kernels, same launches per problem, same results as
  `quantize_device` one weight at a time; concurrent=False (the default) runs exactly that on the
  caller's stream -- each factorisation and OBS loop still overlaps its own lookahead side stream.
  The stream-per-problem arrangement is opt-in: over the end-of-round-2 runs it measured between
  0.7x and 2.3x the one-stream time on the Llama-7B layer (70 / 79 / 88 / 188 ms against 82-99 ms; more
  than eight streams share the device's hardware queues), and one 2-GPU bench run out of three met
  a non-positive pivot in its first concurrent step that neither a single-GPU replay of the same
  Hessians nor two further 2-GPU runs reproduced."""
  import torch
  from ... import device
  main = torch.cuda.current_stream()
  keys = list(dict.fromkeys(feeds))
  rot = []
  for w in weights:
    r, n = hadamard_rotation.rotate_with_diagonal_hadamard_device(w, tuple(w.shape), max_hadamard_size)
    rot.append((r, n))
  hinv, infos, done = {}, [], {}
  for key in keys:
    n = hadamard_rotation.hadamard_size_for(int(hessians[key].shape[0]), max_hadamard_size)
    s = torch.cuda.Stream() if concurrent else main
    s.wait_stream(main)
    with torch.cuda.stream(s):
      h_rot = rotate_hessian_device(hessians[key], n)
      hi, info = device.hessian_inverse(h_rot, damp, check=False)
    hinv[key], done[key] = hi, s
    infos.append(info)
  out = []
  streams = []
  for (r, n), key in zip(rot, feeds):
    s = torch.cuda.Stream() if concurrent else main
    s.wait_stream(main)
    s.wait_stream(done[key])
    with torch.cuda.stream(s):
      mn, mx, _ = device.row_stats(r)
      zp, scale, _ = device.scale_zp_from_minmax(mn, mx, num_bits, symmetric, False)
      q = device.gptq_quantize(r, hinv[key], scale.reshape(-1), zp.reshape(-1), 0, num_bits, symmetric)
    for t in (q, scale, zp):
      t.record_stream(main)
    hinv[key].record_stream(s)
    out.append((q, scale, zp, n))
    streams.append(s)
  for s in streams + list(done.values()):
    main.wait_stream(s)
  for info in infos:  # one read-back each, after everything is queued
    device.check_hessian_info(info)
  return out


def get_tensor_quant_params(
    op_info: qtyping.OpInfo,
    tensor_quant_config: qtyping.TensorQuantizationConfig,
    tensor_content: Optional[np.ndarray] = None,
    tensor_qsv: Optional[Mapping[str, Any]] = None,
) -> qtyping.UniformQuantParams:
  """Rotated GPTQ for a constant FULLY_CONNECTED weight; runtime tensors and weights without a
  Hessian fall back to what `gptq.get_tensor_quant_params` does for them."""
  cfg = tensor_quant_config
  act_qsv = tensor_qsv.get("activation_tensor_qsv") if tensor_qsv else None
  if tensor_content is None or act_qsv is None or "hessian" not in act_qsv or (tensor_qsv and "min" in tensor_qsv):
    return gptq.get_tensor_quant_params(op_info, cfg, tensor_content, tensor_qsv)
  if tensor_content.ndim != 2:
    raise ValueError("Hadamard rotation + GPTQ expects a 2-D FULLY_CONNECTED weight")
  if tensor_content.dtype != np.float32:
    raise ValueError(f"only float32 weights are quantised, got {tensor_content.dtype}")
  if cfg.granularity != qtyping.QuantGranularity.CHANNELWISE:
    raise ValueError("Hadamard rotation + GPTQ supports CHANNELWISE granularity")
  if cfg.num_bits > 8:
    raise ValueError(f"device GPTQ supports num_bits <= 8, got {cfg.num_bits}")
  rows, k = tensor_content.shape
  hessian = act_qsv["hessian"]
  if hessian.shape[0] != k:
    raise ValueError(f"Hessian order {hessian.shape[0]} does not match the weight's {k} input features")
  q, scale, zp, n = quantize_device(hostio.to_device(tensor_content, np.float32), hessian, cfg.num_bits,
                                    bool(cfg.symmetric), cfg.algorithm_params.get("max_hadamard_size"))
  return qtyping.UniformQuantParams(
      num_bits=cfg.num_bits, quantized_dimension=0, scale=hostio.to_host(scale).reshape(rows, 1),
      zero_point=hostio.to_host(zp).reshape(rows, 1).astype(uqt.numpy_dtype_for(cfg.num_bits)),
      symmetric=cfg.symmetric, quantized_data=hostio.to_host(q), block_size=0,
      hadamard=qtyping.UniformQuantParams.HadamardRotationParams(
          random_binary_vector=np.ones(n, dtype=np.int8), hadamard_size=n))
