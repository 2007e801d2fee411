"""`HADAMARD_ROTATION`: block-diagonal Hadamard rotation of the last axis, then OCTAV.

Mirror of ai_edge_quantizer/algorithms/uniform_quantize/hadamard_rotation.py
(`_make_hadamard_matrix` :48-90, `_rotate_with_diagonal_hadamard` :93-134,
`get_tensor_quant_params` :137-203).  The rotation is a shared-memory fast
Walsh-Hadamard transform on the device (`aeqb_hadamard_rows_f32`); the rotated
weight stays in HBM and feeds the OCTAV + fused requantisation kernels directly.
"""
from __future__ import annotations

import math
from typing import Any, Optional

import numpy as np

from ... import hostio
from ... import qtyping
from . import octav

ALGORITHM_KEY = "HADAMARD_ROTATION"


def hadamard_size_for(last_dim: int, max_size: Optional[int] = None) -> int:
  """Largest power of two dividing `last_dim`, capped to a power of two <= max_size (:121-123)."""
  n = math.gcd(int(last_dim), 2**30)
  if max_size:
    n = min(n, 1 << (int(max_size).bit_length() - 1))
  return n


def rotate_with_diagonal_hadamard_device(x_dev, shape, max_size: Optional[int] = None):
  """(rotated device tensor [prod(shape[:-1]), shape[-1]], hadamard_size)."""
  from ... import device
  n = hadamard_size_for(shape[-1], max_size)
  x2 = x_dev.reshape(-1, shape[-1])
  if n == 1:  # odd last dimension: H_1 / sqrt(1) = [[1]], the rotation is the identity
    return x2, n
  return device.hadamard_rows(x2, n), n


def _rotate_with_diagonal_hadamard(tensor_content: np.ndarray, axis: int,
                                   max_size: Optional[int] = None):
  """NumPy-facing mirror: (rotated array, hadamard_size, random_vector)."""
  if axis != tensor_content.ndim - 1:
    raise ValueError(
        "Hadamard rotation is only supported for tensors with quantized"
        " dimension 0 (rotate last dimension).")
  rot, n = rotate_with_diagonal_hadamard_device(
      hostio.to_device(tensor_content, np.float32), tensor_content.shape, max_size)
  return (hostio.to_host(rot).reshape(tensor_content.shape), n, np.ones(n, dtype=np.int8))


def get_tensor_quant_params(
    op_info: qtyping.OpInfo,
    tensor_quant_config: qtyping.TensorQuantizationConfig,
    tensor_content: Optional[np.ndarray] = None,
    tensor_qsv: Optional[dict[str, Any]] = None,
) -> qtyping.UniformQuantParams:
  if tensor_content is None:
    raise ValueError("Hadamard rotation is only supported for weight tensors.")
  if tensor_qsv is not None:
    raise ValueError("Hadamard rotation is not supported for static quantization.")
  if tensor_content.ndim < 2:
    raise ValueError("Hadamard rotation is only supported for tensors with rank >= 2.")
  if tensor_content.dtype != np.float32:
    raise ValueError(f"only float32 weights are quantised, got {tensor_content.dtype}")
  rot, n = rotate_with_diagonal_hadamard_device(
      hostio.to_device(tensor_content, np.float32), tensor_content.shape,
      tensor_quant_config.algorithm_params.get("max_hadamard_size"))
  q = octav.get_tensor_quant_params(op_info, tensor_quant_config, tensor_content, None,
                                    x_dev=rot.reshape(tensor_content.shape))
  return qtyping.UniformQuantParams(
      quantized_dimension=q.quantized_dimension, num_bits=q.num_bits, scale=q.scale,
      zero_point=q.zero_point, symmetric=q.symmetric, quantized_data=q.quantized_data,
      block_size=q.block_size,
      hadamard=qtyping.UniformQuantParams.HadamardRotationParams(
          random_binary_vector=np.ones(n, dtype=np.int8), hadamard_size=n))
