"""CPU oracle for the AI Edge Quantizer numeric hot path (NumPy restatement).

TEST INFRASTRUCTURE — NOT PRODUCT CODE. Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this module, and only as the checker / the timed CPU baseline.
The product path (`ai-edge-quantizer_b200/`) never imports it and has no CPU
fallback.

Every function restates, array-in / array-out, what one piece of the
reference's NumPy path computes, and cites the reference file:line it follows
(paths relative to /root/reference/ai_edge_quantizer/; `uqt` =
algorithms/uniform_quantize/uniform_quantize_tensor.py). FC/EMBEDDING weight
layout is `[rows=out_features, cols=in_features]`: per-channel quantises dim 0,
blockwise splits dim 1 into blocks (utils/tfl_flatbuffer_utils.py:95-106).

Parity pin: `tests/test_oracle_vs_reference.py` runs this module against the
unmodified reference (imported through `oracle/refshim`) on seeded inputs when
/root/reference is mounted, and `tests/test_oracle_golden.py` checks it against
the reference's own literal test vectors and against
`tests/golden/*.npz`, which were produced from the reference by
`tests/golden/make_golden.py`.
"""
from __future__ import annotations

import math

import ml_dtypes
import numpy as np

F32 = np.float32
VALID_LO, VALID_HI = -3e38, 3e38  # naive_min_max_quantize.py:187, gptq.py:61
MSE_MULTIPLIER = {8: 0.05408, 4: 0.37755}  # mse.py:30-33


# --------------------------------------------------------------------------
# integer ranges / dtypes
# --------------------------------------------------------------------------
def qrange(bits: int) -> tuple[float, float]:
  """Signed two's-complement range, as floats (uqt:37-45)."""
  half = 2 ** (bits - 1)
  return float(-half), float(half - 1)


def qdtype(bits: int):
  """Narrowest signed NumPy int that holds `bits` (uqt:88-109)."""
  for width, dt in ((8, np.int8), (16, np.int16), (32, np.int32)):
    if bits <= width:
      return dt
  return np.int64


# --------------------------------------------------------------------------
# a1: weight min/max per granularity  (common_quantize.py:1311-1359)
# --------------------------------------------------------------------------
def weight_minmax(w: np.ndarray, block: int = 0, per_channel: bool = True):
  """min & max of a 2-D weight.

  block > 0  -> shape [R, C/block] (no keepdims, common_quantize.py:1345-1353)
  per_channel -> shape [R, 1]      (keepdims,    common_quantize.py:1337-1344)
  else        -> shape [1, 1]      (tensorwise,  common_quantize.py:1334-1336)
  """
  if block:
    r, c = w.shape
    if c % block:
      raise ValueError(
          f"Quantized dimension {c} in tensor shape {w.shape} is not"
          f" divisible by block size {block}."
      )  # uqt:185-189
    v = w.reshape(r, c // block, block)
    return v.min(axis=2), v.max(axis=2)
  if per_channel:
    axes = tuple(range(1, w.ndim))
    return w.min(axis=axes, keepdims=True), w.max(axis=axes, keepdims=True)
  return w.min(keepdims=True), w.max(keepdims=True)


# --------------------------------------------------------------------------
# a2: scale / zero-point from min/max  (uqt:492-586)
# --------------------------------------------------------------------------
def round_scale_bf16_fp16(scale: np.ndarray) -> np.ndarray:
  """fp32 -> bf16 -> fp16 -> fp32 (uqt:577-581, quantize_tensor.py:129-133)."""
  return scale.astype(ml_dtypes.bfloat16).astype(np.float16).astype(F32)


def scale_zp(mn, mx, bits: int, symmetric: bool, blockwise: bool, clip=None):
  """Returns (zero_point, scale) exactly as uqt:492-586 does."""
  qmin, qmax = qrange(bits)
  floor = 1e-9  # uqt:525
  pos, neg = (None, None) if clip is None else (clip, -clip)
  if blockwise:  # fp16-representable range, uqt:529-550
    hi = np.broadcast_to(np.array(65280) * (2**bits - 1), mx.shape)
    lo = np.broadcast_to(np.array(-65280) * (2**bits), mn.shape)
    pos = hi if pos is None else np.minimum(pos, hi)
    neg = lo if neg is None else np.maximum(neg, lo)
  if symmetric:  # uqt:552-563
    bound = np.maximum(np.maximum(np.abs(mn), np.abs(mx)), floor)
    if clip is not None:
      bound = np.clip(bound, neg, pos)
    scale = bound / qmax
    zp = np.zeros_like(scale, dtype=np.int32)
  else:  # uqt:564-575
    top = np.maximum(mx, np.zeros_like(mx))
    bot = np.minimum(mn, np.zeros_like(mn))
    bound = np.maximum(top - bot, floor)
    if clip is not None:
      bound = np.clip(bound, -clip, clip)
    scale = bound / (qmax - qmin)
    zp = np.rint(qmin - bot / scale)
  if blockwise:
    scale = round_scale_bf16_fp16(scale)
  return zp.astype(qdtype(bits), copy=False), scale  # uqt:585


# --------------------------------------------------------------------------
# a3 / a4: quantise, dequantise  (uqt:273-362, 365-409)
# --------------------------------------------------------------------------
def _spread_blocks(p: np.ndarray, shape, block: int) -> np.ndarray:
  """[R, C/B] -> [R, C] by repeating each entry B times (uqt:246-262)."""
  r, c = shape
  return np.broadcast_to(p[:, :, None], (r, c // block, block)).reshape(r, c)


def quantize(x, scale, zp, bits: int, symmetric: bool, block: int = 0):
  """clip(rint(x / scale + zp)) cast to int (uqt:273-362).

  `scale`/`zp` must already have x's rank (or [R, C/B] when block > 0).
  Narrow range (qmin + 1) only when symmetric and bits >= 8 (uqt:313-315).
  Rows are processed in <= 32 MiB slabs like uqt:323-354 so the CPU-baseline
  timing has the reference's memory behaviour (without its empty iterations).
  """
  if block:
    if x.shape[1] % block:
      raise ValueError(
          "Tensor dimension must be divisible by block size. Got dimension:"
          f" {x.shape[1]} and block size: {block}"
      )  # uqt:633-638
    scale = _spread_blocks(scale, x.shape, block)
    zp = _spread_blocks(zp, x.shape, block)
  if not np.issubdtype(zp.dtype, np.signedinteger):
    raise ValueError(
        f"zero_points need to be {np.signedinteger}. But the actual type is"
        f" {zp.dtype}."
    )  # uqt:316-321
  qmin, qmax = qrange(bits)
  lo = qmin + 1 if (symmetric and bits >= 8) else qmin
  out_dt = qdtype(bits)

  def slab(xs, ss, zs):
    q = np.divide(xs, ss)
    q = np.add(q, zs, out=q)
    np.rint(q, out=q)
    np.clip(q, lo, qmax, out=q)
    with np.errstate(invalid="ignore"):
      return q.astype(out_dt)

  big = x.ndim > 1 and x.nbytes > 32 * 2**20
  if not big:
    with np.errstate(divide="ignore", invalid="ignore"):
      return slab(x, scale, zp)
  flat = x.reshape(-1, x.shape[-1])
  sb = np.broadcast_to(scale, x.shape).reshape(flat.shape)
  zb = np.broadcast_to(zp, x.shape).reshape(flat.shape)
  out = np.empty(flat.shape, dtype=out_dt)
  step = max(1, (32 * 2**20) // (flat.shape[1] * flat.dtype.itemsize))
  with np.errstate(divide="ignore", invalid="ignore"):
    for k in range(0, flat.shape[0], step):
      out[k:k + step] = slab(flat[k:k + step], sb[k:k + step], zb[k:k + step])
  return out.reshape(x.shape)


def dequantize(q, scale, zp, block: int = 0):
  """(q - zp) * scale with blockwise re-broadcast (uqt:365-409)."""
  if block:
    scale = _spread_blocks(scale.reshape(q.shape[0], -1), q.shape, block)
    zp = _spread_blocks(zp.reshape(q.shape[0], -1), q.shape, block)
  return np.multiply(q - zp, scale)


# --------------------------------------------------------------------------
# a5: min-max requantisation  (naive_min_max_quantize.py:34-110)
# --------------------------------------------------------------------------
def minmax_requant(w, bits: int, symmetric: bool = True, block: int = 0,
                   per_channel: bool = True):
  """Returns dict(scale, zero_point, q) for a 2-D fp32 weight."""
  mn, mx = weight_minmax(w, block, per_channel)
  zp, scale = scale_zp(mn, mx, bits, symmetric, bool(block))
  return dict(scale=scale, zero_point=zp,
              q=quantize(w, scale, zp, bits, symmetric, block))


# --------------------------------------------------------------------------
# a6: OCTAV clipping search  (octav.py:30-112, 115-227)
# --------------------------------------------------------------------------
def octav_clip(x, bits: int, axis, max_iterations: int = 10,
               divisor: float = 3.0, early_stop: bool = True,
               return_trace: bool = False):
  """Newton iterations of OCTAV eq. (6) (octav.py:62-112).

  c <- sum{|x| : |x| >= c} / ((1 - s) * count{|x| >= c} + s * N),
  s = f32(4^-bits / divisor), start c = 1; global allclose early stop.
  """
  axis = (axis,) if isinstance(axis, int) else tuple(axis)
  # np.prod -> np.int64 scalar, so `s * n` below is float64 (NEP 50) and the
  # add is evaluated in float64 then stored to fp32 (octav.py:58, 107).
  n = np.prod([x.shape[a] for a in axis])
  s = np.asarray(4.0 ** (-bits) / divisor, dtype=F32)
  c = np.ones([1 if i in axis else d for i, d in enumerate(x.shape)], F32)
  trace = []
  for _ in range(max_iterations):
    prev = c
    hi = x >= prev
    cnt = np.count_nonzero(hi, axis=axis, keepdims=True).astype(F32)
    num = np.sum(x, axis=axis, where=hi, keepdims=True, dtype=F32)
    lo = x <= -prev
    cnt = np.add(cnt, np.count_nonzero(lo, axis=axis, keepdims=True), out=cnt)
    num = np.subtract(num, np.sum(x, axis=axis, where=lo, keepdims=True), out=num)
    den = np.multiply(cnt, 1.0 - s, out=cnt)
    den = np.add(den, s * n, out=den)
    c = np.divide(num, den, out=num)
    trace.append(c)
    if early_stop and np.allclose(prev, c):
      break
  return (c, trace) if return_trace else c


def octav_requant(w, bits: int, block: int = 0, per_channel: bool = True):
  """OCTAV-clipped symmetric requantisation (octav.py:115-227)."""
  mn, mx = weight_minmax(w, block, per_channel)
  if block:
    r, c = w.shape
    clip = octav_clip(w.reshape(r, c // block, block), bits, 2).reshape(mn.shape)
  elif per_channel:
    clip = octav_clip(w, bits, tuple(range(1, w.ndim)))
  else:
    clip = octav_clip(w, bits, tuple(range(w.ndim))).reshape(mn.shape)
  zp, scale = scale_zp(mn, mx, bits, True, bool(block), clip)
  return dict(scale=scale, zero_point=zp, clip=clip,
              q=quantize(w, scale, zp, bits, True, block))


# --------------------------------------------------------------------------
# a7: MSE closed form  (mse.py:36-128)
# --------------------------------------------------------------------------
def mse_requant(w, bits: int, per_channel: bool = True):
  """scale = k * sqrt(mean(w^2)); int32 zero zp (mse.py:105-109)."""
  axes = tuple(range(1, w.ndim)) if per_channel else None
  scale = MSE_MULTIPLIER[bits] * np.sqrt(np.mean(w**2, axis=axes, keepdims=True))
  zp = np.zeros_like(scale, dtype=np.int32)
  return dict(scale=scale, zero_point=zp,
              q=quantize(w, scale, zp, bits, True))


# --------------------------------------------------------------------------
# a8: block-diagonal Hadamard rotation  (hadamard_rotation.py:48-203)
# --------------------------------------------------------------------------
def hadamard_size(cols: int, max_size: int | None = None) -> int:
  """Largest power of two dividing cols, capped (hadamard_rotation.py:121-123)."""
  n = math.gcd(cols, 2**30)
  if max_size:
    n = min(n, 1 << (max_size.bit_length() - 1))
  return n


def hadamard_matrix(n: int) -> np.ndarray:
  """Sylvester H_n / sqrt(n) as fp32 (hadamard_rotation.py:79-89)."""
  if n < 2 or n & (n - 1):
    raise ValueError("Hadamard matrix size must be a power of 2. ")
  h2 = np.array([[1, 1], [1, -1]], dtype=np.int8)
  h = h2
  while h.shape[0] < n:
    h = np.kron(h, h2)
  return h / np.sqrt(n, dtype=F32)  # int8 / f32 scalar -> fp32 entries +-fl(1/sqrt n)


def hadamard_rotate(w, max_size: int | None = None):
  """W.reshape(-1, n) @ H_n/sqrt(n), fp32 (hadamard_rotation.py:128-129)."""
  n = hadamard_size(w.shape[-1], max_size)
  rot = np.matmul(w.reshape(-1, n), hadamard_matrix(n)).reshape(w.shape)
  return rot, n


def hadamard_requant(w, bits: int, max_size: int | None = None,
                     per_channel: bool = True):
  """Rotate, then OCTAV (hadamard_rotation.py:137-203)."""
  rot, n = hadamard_rotate(w, max_size)
  out = octav_requant(rot, bits, 0, per_channel)
  out.update(hadamard_size=n, random_binary_vector=np.ones(n, np.int8), rotated=rot)
  return out


# --------------------------------------------------------------------------
# a12: activation min/max with the open-interval filter
#      (common_quantize.py:1362-1413, 1416-1456)
# --------------------------------------------------------------------------
def activation_minmax(x, lo: float | None = VALID_LO, hi: float | None = VALID_HI):
  """Scalar min over x>lo / max over x<hi, raw fallback if all were dropped."""
  shape = (1,) * x.ndim
  if np.issubdtype(x.dtype, np.integer):
    return np.reshape(x.min(), shape), np.reshape(x.max(), shape)
  if lo is None:
    t_min = x.min()
  else:
    t_min = np.min(x, where=x > lo, initial=np.inf)
    if t_min == np.inf:
      t_min = x.min()
  if hi is None:
    t_max = x.max()
  else:
    t_max = np.max(x, where=x < hi, initial=-np.inf)
    if t_max == -np.inf:
      t_max = x.max()
  return np.reshape(t_min, shape), np.reshape(t_max, shape)


def activation_qsv(x, lo=VALID_LO, hi=VALID_HI):
  """min / max / num_samples QSV (common_quantize.py:1448-1456)."""
  mn, mx = activation_minmax(x, lo, hi)
  return {"min": mn, "max": mx,
          "num_samples": np.array(x.shape[0] if x.ndim > 0 else 1)}


# --------------------------------------------------------------------------
# a11: QSV merge rules  (utils/qsv_utils.py:25-122)
# --------------------------------------------------------------------------
def ema_update(qsv, new, smoothing: float = 0.95):
  """0.95*old + 0.05*new on min/max; first observation kept (qsv_utils.py:43-68)."""
  if not qsv:
    return new
  return {k: smoothing * qsv[k] + (1.0 - smoothing) * new[k] for k in ("min", "max")}


def minmax_union(qsv, new):
  """Elementwise min of mins, max of maxes (qsv_utils.py:105-122)."""
  if not qsv:
    return new
  return {"min": np.minimum(qsv["min"], new["min"]),
          "max": np.maximum(qsv["max"], new["max"])}


def gptq_update(qsv, new):
  """EMA on min/max + sample-weighted Hessian mean (qsv_utils.py:71-102)."""
  if not qsv:
    return new
  out = ema_update(qsv, new)
  a, b = qsv["num_samples"], new["num_samples"]
  tot = a + b
  if tot == 0:
    out["hessian"], out["num_samples"] = new["hessian"], 0
  else:
    out["hessian"] = (qsv["hessian"] * a + new["hessian"] * b) / tot
    out["num_samples"] = tot
  return out


def ema_sequence(mins, maxs, smoothing: float = 0.95):
  """Folds per-batch (min, max) in batch order; returns the final pair."""
  q = {}
  for mn, mx in zip(mins, maxs):
    q = ema_update(q, {"min": mn, "max": mx}, smoothing)
  return q["min"], q["max"]


# --------------------------------------------------------------------------
# a9 / a10: GPTQ  (gptq.py:55-108, 111-128, 131-216, 219-300)
# --------------------------------------------------------------------------
def gptq_hessian(x):
  """(2 / num_samples) * X^T X with X = x.reshape(-1, K) (gptq.py:100-106).

  num_samples is the *batch* dim (common_quantize.py:1453) as a 0-d int array,
  so the scalar is float64 and the product promotes the fp32 GEMM to float64.
  """
  n = np.array(x.shape[0] if x.ndim > 0 else 1)
  x2 = x.reshape(-1, x.shape[-1])
  return (2.0 / n) * x2.T.dot(x2)


def gptq_damped_diagonal(h, damp: float = 0.01):
  """diag with zeros -> 1, plus damp * mean (gptq.py:115-117)."""
  d = np.diag(h)
  d = np.where(d, d, 1.0)
  return d + damp * np.mean(d)


def gptq_hessian_inverse(h, damp: float = 0.01, mutate: bool = False):
  """Damped inverse through Cholesky + triangular inverse (gptq.py:111-128).

  Reference quirk: `orig_diag = np.diag(hessian)` is a view, so the "restore"
  at gptq.py:123 writes the damped diagonal back onto itself and the CALLER's
  Hessian keeps the damping (it accumulates when one activation feeds several
  FC ops).  `mutate=True` reproduces that side effect on `h`.
  """
  import scipy.linalg
  d = gptq_damped_diagonal(h, damp)
  if mutate:
    np.fill_diagonal(h, d)
  h = np.array(h, copy=True)
  np.fill_diagonal(h, d)
  low = np.linalg.cholesky(h)
  low_inv, info = scipy.linalg.lapack.strtri(low, lower=True)
  assert info == 0
  return np.einsum("ji,jk->ik", low_inv, low_inv)


def gptq_quantize(w, scale, zp, hinv, bits: int, symmetric: bool = True,
                  block: int = 0, blocksize: int = 64):
  """Lazy-block OBS column loop (gptq.py:131-216). Returns int array [R, K]."""
  fp = w.copy()
  r, k = fp.shape
  q_all = np.zeros((r, k), dtype=qdtype(bits))
  for b0 in range(0, k, blocksize):
    b1 = min(b0 + blocksize, k)
    wb = fp[:, b0:b1]
    err = np.zeros_like(wb)
    for i in range(b1 - b0):
      col = b0 + i
      if block:  # per-column scale pick, gptq.py:177-189
        sc, z = scale[:, col // block], zp[:, col // block]
      else:
        sc, z = scale.reshape(-1), zp.reshape(-1)
      if sc.size == 1:
        sc, z = np.broadcast_to(sc, (r,)), np.broadcast_to(z, (r,))
      wc = wb[:, i]
      qc = quantize(wc[:, None], sc[:, None], z[:, None], bits, symmetric)
      dq = np.multiply(qc - z[:, None], sc[:, None]).reshape(-1)
      q_all[:, col] = qc.reshape(-1)
      np.subtract(wc, dq, out=err[:, i])
      err[:, i] /= hinv[col, col]
      if i < b1 - b0 - 1:
        wb[:, i + 1:] -= np.outer(err[:, i], hinv[col, col + 1:b1])
    fp[:, b1:] -= np.matmul(err, hinv[b0:b1, b1:])
  return q_all


def gptq_requant(w, hessian, bits: int, symmetric: bool = True, block: int = 0,
                 per_channel: bool = True):
  """min/max scales on the original W, then the OBS loop (gptq.py:219-300)."""
  mn, mx = weight_minmax(w, block, per_channel)
  zp, scale = scale_zp(mn, mx, bits, symmetric, bool(block))
  hinv = gptq_hessian_inverse(hessian)
  return dict(scale=scale, zero_point=zp, hinv=hinv,
              q=gptq_quantize(w, scale, zp, hinv, bits, symmetric, block))


def hadamard_rotate_hessian(hessian, n: int):
  """R^T H R in float64 for the block-diagonal R = diag(H_n / sqrt(n), ...) that
  `hadamard_rotate` applies to the weight's last axis.  No reference function computes this: it
  follows from the two it composes — the runtime feeds the rotated op x R (the activation-side
  INSERT_HADAMARD_ROTATION, hadamard_rotation.py:206-283), and gptq.calibrate's Hessian of x R is
  (2 / n_s) (X R)^T (X R) = R^T H R (gptq.py:100-106)."""
  h = np.asarray(hessian, dtype=np.float64)
  k = h.shape[0]
  if n <= 1:
    return h.copy()
  r = hadamard_matrix(n).astype(np.float64)
  out = h.reshape(k // n, n, k // n, n)
  out = np.einsum("ab,ibjc,cd->iajd", r.T, out, r, optimize=True)
  return out.reshape(k, k)


def hadamard_gptq_requant(w, hessian, bits: int, max_size: int | None = None, symmetric: bool = True):
  """BASELINE.json configs[4] as one algorithm: rotate W (hadamard_rotation.py:93-134), rotate the
  Hessian with the same R, then GPTQ on the pair (gptq.py:219-300: min/max scales of the ROTATED
  weight, damped inverse, OBS loop).  Each stage is the pinned oracle of its reference function."""
  rot, n = hadamard_rotate(w, max_size)
  h_rot = hadamard_rotate_hessian(hessian, n)
  out = gptq_requant(rot, h_rot, bits, symmetric)
  out.update(hadamard_size=n, random_binary_vector=np.ones(n, np.int8), rotated=rot, hessian_rotated=h_rot)
  return out


# --------------------------------------------------------------------------
# §8(f) row 3: dequantized_weight_recovery / float_casting
# --------------------------------------------------------------------------
def dwr_group_scales(groups: np.ndarray, min_scale: float = 1e-9) -> np.ndarray:
  """One recovered scale per row of `groups` [n_groups, group_len]
  (dequantized_weight_recovery.py:181-209): sort |x| with a 0 appended, smallest
  adjacent difference > 1e-9, floored at min_scale; min_scale when all equal."""
  a = np.abs(groups)
  a = np.hstack([a, np.zeros((a.shape[0], 1), dtype=a.dtype)])
  d = np.diff(np.sort(a, axis=1), axis=1)
  m = np.min(np.where(d > 1e-9, d, np.inf), axis=1)
  s = np.maximum(m, min_scale)
  s[s == np.inf] = min_scale
  return s


def dwr_requant(w: np.ndarray, bits: int, block: int = 0, per_channel: bool = True):
  """scale / zero_point / q of dequantized_weight_recovery.get_tensor_quant_params
  (:220-262) for a 2-D FC weight (quantised dim 0, blocks along dim 1)."""
  if block:
    scale = dwr_group_scales(w.reshape(-1, block)).reshape(w.shape[0], w.shape[1] // block)
  elif per_channel:
    scale = dwr_group_scales(w).reshape(w.shape[0], 1)
  else:
    u = np.unique(np.append(np.abs(np.ravel(w)), 0))  # _get_scale (:66-77)
    v = float(np.maximum(np.min(np.diff(u)), 1e-9)) if u.size > 1 else 1e-9
    scale = np.array([[v]])
  zp = np.zeros_like(scale, dtype=np.int32)
  q = quantize(w, scale, zp, bits, True, block)
  return {"scale": scale, "zero_point": zp, "q": q}


def float_cast(w: np.ndarray) -> np.ndarray:
  """float_casting.materialize_fc_conv's weight (float_casting.py:160-162)."""
  return w.astype(np.float16)


def fake_quantized_weight(rows: int, cols: int, bits: int, block: int = 0, index: int = 0,
                          per_channel: bool = True) -> np.ndarray:
  """A QAT-style weight: integers in the signed range times a per-group fp32 scale."""
  rng = np.random.default_rng(3000 + index)
  lo, hi = int(qrange(bits)[0]) + 1, int(qrange(bits)[1])
  q = rng.integers(lo, hi + 1, size=(rows, cols)).astype(F32)
  if block:
    s = rng.uniform(0.001, 0.05, size=(rows, cols // block)).astype(F32)
    return (q.reshape(rows, cols // block, block) * s[:, :, None]).reshape(rows, cols).astype(F32)
  s = rng.uniform(0.001, 0.05, size=(rows, 1) if per_channel else (1, 1)).astype(F32)
  return (q * s).astype(F32)


# --------------------------------------------------------------------------
# §8(f) row 3: OSCAR (float64 throughout, FULLY_CONNECTED weights [out, in])
# --------------------------------------------------------------------------
OSCAR_EPS = 1e-12  # oscar.py:52


def oscar_floor(mu2):
  """Dead-channel guard (oscar.py:56-59)."""
  mu2 = np.asarray(mu2, np.float64)
  return np.maximum(mu2, float(np.max(mu2)) * 1e-8 + OSCAR_EPS)


def oscar_mu2(x):
  """calibrate's per-channel second moment (oscar.py:271-277)."""
  x2 = np.asarray(x, np.float64).reshape(-1, x.shape[-1])
  return np.mean(x2 * x2, axis=0)


def oscar_group_clip(a, m, qmax: int):
  """Per-row exact minimiser of c^2 M / (12 q^2) + sum_j max(a_j - c, 0)^2 m_j (oscar.py:62-104).

  a: [n, g] magnitudes, m: [g] masses.  Breakpoint scan: with the k+1 largest magnitudes
  clipped the objective is a parabola in c, minimised in closed form and clamped to its
  segment; candidate 0 is "no clipping"."""
  n, g = a.shape
  order = np.argsort(-a, axis=1)
  hi = np.take_along_axis(a, order, 1)
  w = m[order]
  mass = float(m.sum()) + OSCAR_EPS
  sm, sam, sa2m = np.cumsum(w, 1), np.cumsum(hi * w, 1), np.cumsum(hi * hi * w, 1)
  c = 2.0 * sam / (mass / (6.0 * qmax * qmax) + 2.0 * sm)
  lo = np.concatenate([hi[:, 1:], np.zeros((n, 1))], 1)
  c = np.clip(c, lo, hi)
  e = (c ** 2) * (mass / (12.0 * qmax * qmax)) + sa2m - 2.0 * c * sam + (c ** 2) * sm
  c0 = hi[:, :1]
  e0 = (c0 ** 2) * (mass / (12.0 * qmax * qmax))
  cc, ee = np.concatenate([c0, c], 1), np.concatenate([e0, e], 1)
  return cc[np.arange(n), np.argmin(ee, 1)]


def oscar_objective(w, s, mu2, block: int) -> float:
  """sum over column groups of (sum_i max_j (|w_ij| s_j)^2) * sum_j mu2_j / s_j^2 (oscar.py:172-189)."""
  m, a = mu2 / (s * s), np.abs(w) * s
  d = w.shape[1]
  g = block if (block and d % block == 0) else d
  total = 0.0
  for b in range(d // g):
    mx = a[:, b * g:(b + 1) * g].max(1)
    total += float((mx * mx).sum()) * float(m[b * g:(b + 1) * g].sum())
  return total


def oscar_channel_scales(w, mu2, block: int = 0, iters: int = 3):
  """Alternating fixed point for the per-input-channel scales (oscar.py:192-251).
  Returns None when no candidate beats s = 1."""
  w = np.asarray(w, np.float64)
  mu2 = oscar_floor(mu2)
  mu = np.sqrt(mu2)
  n, d = w.shape
  g = block if (block and d % block == 0) else d

  def norm(v):
    return np.clip(v / np.exp(np.mean(np.log(v))), 1e-4, 1e4)

  a_base = (w * w).sum(0) + OSCAR_EPS
  identity = oscar_objective(w, np.ones(d), mu2, block)
  s = norm(np.sqrt(mu / np.sqrt(a_base)))
  best = (oscar_objective(w, s, mu2, block), s)
  rows = np.arange(n)
  for _ in range(iters):
    a_eff = np.zeros(d)
    mag = np.abs(w) * s
    for b in range(d // g):
      j = b * g + np.argmax(mag[:, b * g:(b + 1) * g], 1)
      np.add.at(a_eff, j, w[rows, j] ** 2)
    a_eff = np.maximum(a_eff, 0.25 * a_base)
    s = norm(np.sqrt(s * norm(np.sqrt(mu / np.sqrt(a_eff)))))
    loss = oscar_objective(w, s, mu2, block)
    if loss < best[0]:
      best = (loss, s)
  return None if best[0] >= identity else best[1]


def oscar_requant(w, mu2, bits: int, block: int = 0, per_channel: bool = True):
  """oscar.get_tensor_quant_params for a 2-D FC weight (oscar.py:373-458)."""
  w64 = np.asarray(w, np.float64)
  n, d = w64.shape
  s = np.ones(d)
  if mu2 is not None:
    got = oscar_channel_scales(w64, np.asarray(mu2, np.float64).ravel(), block)
    s = got if got is not None else s
  ws = w64 * s
  m = oscar_floor(np.ones(d) if mu2 is None else np.asarray(mu2, np.float64).ravel() / (s * s))
  qmax = 2 ** (bits - 1) - 1
  a = np.abs(ws)
  if block:
    bound = np.stack([oscar_group_clip(a[:, b * block:(b + 1) * block], m[b * block:(b + 1) * block], qmax)
                      for b in range(d // block)], axis=1)
  elif per_channel:
    bound = oscar_group_clip(a, m, qmax).reshape(n, 1)
  else:
    bound = oscar_group_clip(a.reshape(1, n * d), np.tile(m, n), qmax).reshape(1, 1)
  scale = np.maximum(bound, 1e-9) / qmax  # uqt:553-563 on float64 bounds
  if block:
    scale = round_scale_bf16_fp16(scale)
  zp = np.zeros(scale.shape, dtype=qdtype(bits))
  q = quantize(ws, scale, zp, bits, True, block)
  return {"scale": scale, "zero_point": zp, "q": q, "bound": bound, "channel_scale": s,
          "multiplier": (1.0 / s).astype(F32)}


# --------------------------------------------------------------------------
# a13 / a14: bit packing and the serialised blockwise scale
#            (transformations/transformation_utils.py:293-353,
#             transformations/quantize_tensor.py:107-147)
# --------------------------------------------------------------------------
def pack_bits(bits: int, data: np.ndarray) -> np.ndarray:
  """INT4: two nibbles per byte, even index low; INT2: four crumbs, index 0 lowest.

  Odd tails are zero-padded; any other width is returned flattened, untouched.
  """
  flat = data.reshape(-1)
  if bits not in (2, 4):
    return flat
  per = 8 // bits
  mask = (1 << bits) - 1
  n_out = -(-flat.size // per)
  padded = np.zeros(n_out * per, dtype=np.uint8)
  padded[:flat.size] = flat.astype(np.uint8) & mask
  lanes = padded.reshape(n_out, per)
  out = np.zeros(n_out, dtype=np.uint8)
  for j in range(per):
    out |= lanes[:, j] << (bits * j)
  return out


def blockwise_scale_fp16(scale: np.ndarray) -> np.ndarray:
  """The fp16 scale tensor the flatbuffer stores (quantize_tensor.py:129-133)."""
  return scale.astype(ml_dtypes.bfloat16).astype(np.float16)


# --------------------------------------------------------------------------
# synthetic inputs shared by tests and bench
# --------------------------------------------------------------------------
def synthetic_weight(rows: int, cols: int, index: int = 0) -> np.ndarray:
  """N(0,1)*0.02 fp32 with one x20 outlier per 1024 elements; seed 1000+index."""
  rng = np.random.default_rng(1000 + index)
  w = rng.standard_normal((rows, cols), dtype=F32) * F32(0.02)
  flat = w.reshape(-1)
  flat[::1024] *= F32(20.0)
  return w


def synthetic_activation(shape, index: int = 0) -> np.ndarray:
  """N(0,1) fp32 activations; seed 2000+index."""
  return np.random.default_rng(2000 + index).standard_normal(shape, dtype=F32)
