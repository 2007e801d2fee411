"""Core tensor math with the reference's NumPy-facing signatures, computed on the GPU.

Mirror of ai_edge_quantizer/algorithms/uniform_quantize/uniform_quantize_tensor.py
(`uqt`): `tensor_zp_scale_from_min_max` :492-586, `uniform_quantize` :273-362,
`uniform_dequantize` :365-409, `fix_quantization_params_rank` :112-161,
`_is_valid_quantization_params` :589-638, range / dtype helpers :26-109.
Arrays arrive and leave as NumPy (the plug-in contract); every number is
produced by libaeqb200.so.  Shape / rank validation and error texts are the
reference's and stay on the host.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np

from ... import hostio
from ... import qtyping
from ...utils import tfl_flatbuffer_utils


@dataclasses.dataclass(frozen=True)
class IntType:
  num_bits: int
  signed: bool


def is_blockwise(granularity) -> bool:
  return "BLOCKWISE" in str(granularity)


def get_quantized_range(qtype: IntType) -> tuple[float, float]:
  if qtype.signed:
    half = 2 ** (qtype.num_bits - 1)
    return float(-half), float(half - 1)
  return 0.0, float(2**qtype.num_bits - 1)


_BLOCK_OF = {
    qtyping.QuantGranularity.BLOCKWISE_32: 32,
    qtyping.QuantGranularity.BLOCKWISE_64: 64,
    qtyping.QuantGranularity.BLOCKWISE_128: 128,
    qtyping.QuantGranularity.BLOCKWISE_256: 256,
}


def extract_block_size_from_granularity(granularity) -> int:
  return _BLOCK_OF.get(granularity, 0)


def numpy_dtype_for(num_bits: int, signed: bool = True):
  for width, s, u in ((8, np.int8, np.uint8), (16, np.int16, np.uint16),
                      (32, np.int32, np.uint32)):
    if num_bits <= width:
      return s if signed else u
  return np.int64 if signed else np.uint64


def assign_quantized_type(tensor: np.ndarray, qtype: IntType) -> np.ndarray:
  return tensor.astype(numpy_dtype_for(qtype.num_bits, qtype.signed), copy=False)


def _blockwise_shape(shape, quantized_dim: int, block_size: int) -> list[int]:
  out = []
  for axis, extent in enumerate(shape):
    if axis != quantized_dim:
      out.append(extent)
      continue
    if extent % block_size:
      raise ValueError(
          f"Quantized dimension {extent} in tensor shape {shape} is not"
          f" divisible by block size {block_size}.")
    out += [extent // block_size, block_size]
  return out


def reshape_data_for_blockwise(tensor_data: np.ndarray, op_name, granularity):
  """[.., C, ..] -> [.., C/B, B, ..] view and the new reduce axis (uqt:197-219)."""
  qdim = tfl_flatbuffer_utils.TFL_OP_TO_BLOCKWISE_WEIGHT_QUANTIZED_DIM[op_name]
  block = extract_block_size_from_granularity(granularity)
  return tensor_data.reshape(_blockwise_shape(tensor_data.shape, qdim, block)), qdim + 1


def fix_quantization_params_rank(tensor_data: np.ndarray,
                                 quantization_params: qtyping.UniformQuantParams):
  """Gives scale / zero_point the tensor's rank (1s everywhere but the quantised axis)."""
  scales, zps = quantization_params.scale, quantization_params.zero_point
  if tensor_data.ndim == scales.ndim:
    return quantization_params
  if tensor_data.ndim == 0:
    if scales.size != 1 or zps.size != 1:
      raise ValueError(
          "Scale and zero_point must contain single element for scalar tensor."
          f" Got scale: {scales}, zero_point: {zps}")
    scales, zps = np.array(scales.item()), np.array(zps.item())
  else:
    dims = [d for d in range(tensor_data.ndim)
            if d != quantization_params.quantized_dimension]
    scales, zps = np.expand_dims(scales, axis=dims), np.expand_dims(zps, axis=dims)
  return qtyping.UniformQuantParams(
      scale=scales, zero_point=zps, num_bits=quantization_params.num_bits,
      symmetric=quantization_params.symmetric,
      quantized_dimension=quantization_params.quantized_dimension,
      quantized_data=quantization_params.quantized_data,
      block_size=quantization_params.block_size)


def _is_valid_quantization_params(tensor_data, quantization_params) -> None:
  scale, zp = quantization_params.scale, quantization_params.zero_point
  if scale.shape != zp.shape and zp.size != 1:
    raise ValueError(
        "scale and zero_point must have the same shape or zero_point must have"
        f" only one element. Got {scale.shape} and {zp.shape}")
  if tensor_data.ndim != scale.ndim or tensor_data.ndim != zp.ndim:
    raise ValueError(
        f"Ranks of scales ({scale.ndim}) and zps"
        f" ({zp.ndim}) must be the same as the tensor rank"
        f" ({tensor_data.ndim}).")
  block = quantization_params.block_size
  if block != 0 and tensor_data.shape[quantization_params.quantized_dimension] % block != 0:
    raise ValueError(
        "Tensor dimension must be divisible by block size. Got dimension:"
        f" {tensor_data.shape[quantization_params.quantized_dimension]} and"
        f" block size: {block}")


def _param_layout(shape, scale_shape) -> tuple[int, int]:
  """(channels, inner) such that element i uses parameter (i // inner) % channels."""
  if int(np.prod(scale_shape)) == 1:
    return 1, 1
  axes = [a for a, (s, t) in enumerate(zip(scale_shape, shape)) if s != 1]
  if len(axes) != 1 or scale_shape[axes[0]] != shape[axes[0]]:
    raise ValueError(
        f"scale of shape {tuple(scale_shape)} does not select a single axis of a"
        f" tensor of shape {tuple(shape)}")
  axis = axes[0]
  return int(shape[axis]), int(np.prod(shape[axis + 1:], dtype=np.int64))


def uniform_quantize(tensor_data: np.ndarray,
                     quantization_params: qtyping.UniformQuantParams,
                     is_blockwise_quant: bool = False) -> np.ndarray:
  """clip(rint(x / scale + zp)) cast to the narrowest signed int (uqt:273-362)."""
  tensor_data = np.asarray(tensor_data)
  scale = np.asarray(quantization_params.scale)
  zp = quantization_params.zero_point
  bits = quantization_params.num_bits
  if is_blockwise_quant:
    if quantization_params.quantized_dimension is None:
      raise ValueError("Quantized dimension must be specified.")
    block = quantization_params.block_size
    if block is None or block <= 0:
      raise ValueError("Block size must be specified and positive.")
    qdim = quantization_params.quantized_dimension
    _blockwise_shape(tensor_data.shape, qdim, block)  # divisibility check + message
    if qdim != tensor_data.ndim - 1:
      raise ValueError("blockwise quantisation is supported along the last axis only")
    channels, inner = tensor_data.size // block, block
    if zp is None or np.size(zp) == 0:
      zp = None
  else:
    quantization_params = fix_quantization_params_rank(tensor_data, quantization_params)
    _is_valid_quantization_params(tensor_data, quantization_params)
    scale, zp = quantization_params.scale, quantization_params.zero_point
    channels, inner = _param_layout(tensor_data.shape, scale.shape)
  if zp is not None and not np.issubdtype(np.asarray(zp).dtype, np.signedinteger):
    raise ValueError(
        f"zero_points need to be {np.signedinteger}. But the actual type is"
        f" {np.asarray(zp).dtype}.")
  if bits > 16:
    raise ValueError(f"device quantisation supports num_bits <= 16, got {bits}")
  from ... import device
  x = hostio.to_device(tensor_data, np.float32)
  q = device.quantize(
      x, hostio.to_device(scale.reshape(-1), np.float32),
      None if zp is None else hostio.to_device(np.asarray(zp).reshape(-1), np.int32),
      bits, bool(quantization_params.symmetric), channels, inner)
  return hostio.to_host(q).reshape(tensor_data.shape)


def uniform_dequantize(tensor_data: np.ndarray,
                       quantization_params: qtyping.UniformQuantParams) -> np.ndarray:
  """(q - zp) * scale in fp32 (uqt:365-409), blockwise scales re-broadcast."""
  tensor_data = np.asarray(tensor_data)
  qp = quantization_params
  if qp.block_size != 0:
    qdim = qp.quantized_dimension
    if qdim == 0:  # XNNPack-style dimension, uqt:383-387
      qdim = 1
    if qdim != tensor_data.ndim - 1:
      raise ValueError("blockwise dequantisation is supported along the last axis only")
    _blockwise_shape(tensor_data.shape, qdim, qp.block_size)
    channels, inner = tensor_data.size // qp.block_size, qp.block_size
    scale, zp = np.asarray(qp.scale), qp.zero_point
  else:
    qp = fix_quantization_params_rank(tensor_data, qp)
    _is_valid_quantization_params(tensor_data, qp)
    scale, zp = qp.scale, qp.zero_point
    channels, inner = _param_layout(tensor_data.shape, scale.shape)
  if tensor_data.dtype.itemsize > 4 or not np.issubdtype(tensor_data.dtype, np.signedinteger):
    tensor_data = tensor_data.astype(np.int32)
  zp_arr = None if zp is None or np.size(zp) == 0 else np.asarray(zp)
  wrap8 = (tensor_data.dtype == np.int8 and zp_arr is not None and zp_arr.dtype == np.int8)
  from ... import device
  out = device.dequantize(
      hostio.to_device(tensor_data),
      hostio.to_device(np.asarray(scale).reshape(-1), np.float32),
      None if zp_arr is None else hostio.to_device(zp_arr.reshape(-1), np.int32),
      channels, inner, wrap8=wrap8)
  return hostio.to_host(out).reshape(tensor_data.shape)


def tensor_zp_scale_from_min_max(min_value, max_value, num_bits: int, symmetric: bool,
                                 granularity, clipping_values: Optional[np.ndarray] = None):
  """(zero_point, scale) from min / max, any shape (uqt:492-586)."""
  from ... import device
  mn = np.asarray(min_value, dtype=np.float32)
  mx = np.asarray(max_value, dtype=np.float32)
  clip = None if clipping_values is None else hostio.to_device(
      np.broadcast_to(np.asarray(clipping_values, dtype=np.float32), mn.shape))
  zp, scale, _ = device.scale_zp_from_minmax(
      hostio.to_device(mn), hostio.to_device(mx), num_bits, bool(symmetric),
      is_blockwise(granularity), clip)
  zp_np = hostio.to_host(zp).reshape(mn.shape)
  return (assign_quantized_type(zp_np, IntType(num_bits, True)),
          hostio.to_host(scale).reshape(mn.shape))
