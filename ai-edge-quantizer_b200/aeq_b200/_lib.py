"""ctypes binding of libaeqb200.so — the only way arithmetic is reached.

There is deliberately no fallback: if the shared library is missing or a
symbol cannot be resolved this module raises, and every product entry point
above it fails with it.
"""
from __future__ import annotations

import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AEQB200_LIB") or os.path.join(_HERE, "libaeqb200.so")  # env: A/B builds
HEADER_PATH = os.path.normpath(
    os.path.join(_HERE, "..", "..", "include", "aeqb200.h"))

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_L = _c.c_int64
_F = _c.c_float
_D = _c.c_double

# name -> (restype, argtypes); must list every function include/aeqb200.h declares
# (tests/test_cabi_symbols.py parses the header and checks both directions).
SIGNATURES = {
    "aeqb_version": (_I, []),
    "aeqb_last_error": (_c.c_char_p, []),
    "aeqb_launch_count": (_L, []),
    "aeqb_host_requant_rows_batch_f32": (_I, [_P, _L, _I, _I]),
    "aeqb_host_requant_blocks_batch_f32": (_I, [_P, _L, _I, _I]),
    "aeqb_host_requant_mse_rows_batch_f32": (_I, [_P, _L, _I, _F]),
    "aeqb_host_set_devices": (_I, [_P, _I]),
    "aeqb_host_worker_threads": (_I, []),
    "aeqb_host_copy_in": (_I, [_P, _P, _c.c_size_t, _P]),
    "aeqb_host_copy_out": (_I, [_P, _P, _c.c_size_t, _P]),
    "aeqb_host_alloc": (_P, [_c.c_size_t]),
    "aeqb_host_free": (None, [_P]),
    "aeqb_host_release": (None, []),
    "aeqb_requant_rows_f32": (_I, [_P, _L, _L, _I, _I, _P, _P, _P, _P, _P, _P]),
    "aeqb_requant_given_minmax_f32":
        (_I, [_P, _L, _L, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "aeqb_requant_blocks_f32": (_I, [_P, _L, _L, _I, _I, _P, _P, _P, _P, _P, _P]),
    "aeqb_requant_rows_batch_f32": (_I, [_P, _L, _I, _I, _P]),
    "aeqb_requant_blocks_batch_f32": (_I, [_P, _L, _I, _I, _P]),
    "aeqb_requant_rows_batch_mirror_f32": (_I, [_P, _L, _I, _I, _P, _I, _P]),
    "aeqb_requant_blocks_batch_mirror_f32": (_I, [_P, _L, _I, _I, _P, _I, _P]),
    "aeqb_ema_sequence_f32": (_I, [_P, _L, _D, _P, _P]),
    "aeqb_peer_alloc": (_I, [_c.c_size_t, _P, _P]),
    "aeqb_peer_open": (_I, [_P, _P]),
    "aeqb_peer_close": (_I, [_P]),
    "aeqb_peer_free": (_I, [_P]),
    "aeqb_minmax_workspace_bytes": (_c.c_size_t, []),
    "aeqb_minmax_tensors_f32": (_I, [_P, _L, _F, _F, _I, _I, _P, _P]),
    "aeqb_minmax_tensor_f32": (_I, [_P, _L, _F, _F, _I, _I, _P, _P, _P]),
    "aeqb_hist_accumulate_f32": (_I, [_P, _L, _F, _F, _I, _I, _P, _P]),
    "aeqb_row_stats_f32": (_I, [_P, _L, _L, _P, _P, _P, _P]),
    "aeqb_minmax_blocks_f32": (_I, [_P, _L, _L, _I, _P, _P, _P]),
    "aeqb_octav_workspace_bytes": (_c.c_size_t, [_L, _I]),
    "aeqb_octav_clip_rows_f32": (_I, [_P, _L, _L, _I, _I, _F, _I, _P, _P, _P]),
    "aeqb_octav_clip_blocks_f32": (_I, [_P, _L, _L, _I, _I, _I, _F, _I, _P, _P, _P]),
    "aeqb_mse_workspace_bytes": (_c.c_size_t, []),
    "aeqb_mse_scale_rows_f32": (_I, [_P, _L, _L, _F, _P, _P, _P]),
    "aeqb_hadamard_rows_f32": (_I, [_P, _L, _L, _L, _P, _P]),
    "aeqb_xtx_workspace_bytes": (_c.c_size_t, [_L, _L]),
    "aeqb_xtx_f32": (_I, [_P, _L, _L, _D, _P, _P, _P]),
    "aeqb_hessian_inverse_workspace_bytes": (_c.c_size_t, [_L]),
    "aeqb_hessian_inverse_f64": (_I, [_P, _L, _D, _I, _P, _P, _P, _P]),
    "aeqb_gptq_workspace_bytes": (_c.c_size_t, [_L, _L]),
    "aeqb_gptq_quantize_f32": (_I, [_P, _L, _L, _P, _P, _P, _L, _I, _I, _I, _I, _P, _P, _P]),
    "aeqb_hessian_merge_f64": (_I, [_P, _D, _P, _D, _P, _L, _P]),
    "aeqb_scale_zp_from_minmax": (_I, [_P, _P, _P, _L, _I, _I, _I, _P, _P, _P, _P]),
    "aeqb_quantize_f32": (_I, [_P, _L, _L, _L, _P, _P, _I, _I, _I, _P, _P]),
    "aeqb_dequantize_f32": (_I, [_P, _I, _L, _L, _L, _P, _P, _I, _I, _P, _P]),
    "aeqb_pack_bits": (_I, [_P, _L, _I, _P, _P]),
    "aeqb_swap_axes": (_I, [_P, _L, _L, _L, _I, _P, _P]),
    "aeqb_requant_mse_rows_f32": (_I, [_P, _L, _L, _I, _F, _P, _P, _P, _P, _P]),
    "aeqb_dwr_workspace_bytes": (_c.c_size_t, [_L, _L]),
    "aeqb_dwr_scales_f32": (_I, [_P, _L, _L, _P, _P, _P]),
    "aeqb_max_abs_diff_f32": (_I, [_P, _P, _L, _P, _P, _P]),
    "aeqb_cast_f32_f16": (_I, [_P, _L, _P, _P]),
    "aeqb_colsq_workspace_bytes": (_c.c_size_t, [_L, _L]),
    "aeqb_colsq_f64": (_I, [_P, _L, _L, _D, _P, _P, _P]),
    "aeqb_oscar_pass_workspace_bytes": (_c.c_size_t, [_L, _L, _L]),
    "aeqb_oscar_pass_f32": (_I, [_P, _L, _L, _L, _P, _P, _P, _P, _P]),
    "aeqb_oscar_clip_workspace_bytes": (_c.c_size_t, [_L, _L, _L]),
    "aeqb_oscar_clip_f32": (_I, [_P, _L, _L, _L, _P, _P, _P, _D, _I, _P, _P, _P]),
    "aeqb_oscar_scale_f64": (_I, [_P, _L, _I, _I, _P, _P]),
    "aeqb_oscar_quantize_f32": (_I, [_P, _L, _L, _L, _P, _P, _I, _P, _P]),
}


class RowsJob(_c.Structure):
  """aeqb_rows_job (include/aeqb200.h)."""
  _fields_ = [("x", _P), ("rows", _L), ("cols", _L), ("clip", _P), ("q", _P),
              ("packed", _P), ("scale", _P), ("zp", _P)]


class BlocksJob(_c.Structure):
  """aeqb_blocks_job (include/aeqb200.h)."""
  _fields_ = [("x", _P), ("rows", _L), ("cols", _L), ("clip", _P), ("q", _P),
              ("packed", _P), ("scale", _P), ("scale_f16", _P)]


class MinmaxJob(_c.Structure):
  """aeqb_minmax_job (include/aeqb200.h)."""
  _fields_ = [("x", _P), ("n", _L), ("out2", _P)]


class AeqbError(RuntimeError):
  """A C-ABI call returned non-zero."""


_lib = None


def header_functions(path: str = HEADER_PATH) -> list[str]:
  """Names of all functions declared in the public header."""
  text = open(path).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(aeqb_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
  """Loads the library once and types every entry point."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python"
        " ai-edge-quantizer_b200/build.py` (nvcc, sm_100a). aeq_b200 has no"
        " CPU fallback.")
  lib = ctypes.CDLL(LIB_PATH)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)  # AttributeError if the symbol is not exported
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib


def call(name: str, *args) -> None:
  """Calls an int-returning entry point and raises AeqbError on failure."""
  lib = load()
  rc = getattr(lib, name)(*args)
  if rc != 0:
    msg = lib.aeqb_last_error().decode("utf-8", "replace")
    raise AeqbError(f"{name} failed: {msg}")
