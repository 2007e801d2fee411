"""Generates tests/golden/*.npz from the UNMODIFIED reference (imported via oracle/refshim).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Every fixture stores the seeded inputs next to the reference's outputs so the
tests need neither the reference nor the RNG.  Shapes are small (seconds on CPU).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import aeq_oracle as O  # noqa: E402
from oracle import refshim  # noqa: E402

Q = refshim.ref("qtyping")
NMM = refshim.ref("algorithms.uniform_quantize.naive_min_max_quantize")
OCT = refshim.ref("algorithms.uniform_quantize.octav")
MSE = refshim.ref("algorithms.uniform_quantize.mse")
HAD = refshim.ref("algorithms.uniform_quantize.hadamard_rotation")
GPTQ = refshim.ref("algorithms.uniform_quantize.gptq")
UQT = refshim.ref("algorithms.uniform_quantize.uniform_quantize_tensor")
CQ = refshim.ref("algorithms.uniform_quantize.common_quantize")
QSV = refshim.ref("utils.qsv_utils")
TU = refshim.ref("transformations.transformation_utils")

G = Q.QuantGranularity
GRAN = {0: G.CHANNELWISE, -1: G.TENSORWISE, 32: G.BLOCKWISE_32, 64: G.BLOCKWISE_64,
        128: G.BLOCKWISE_128, 256: G.BLOCKWISE_256}


def weight(rows, cols, index, special=True):
  w = O.synthetic_weight(rows, cols, index)
  if special and rows >= 4:
    w[1, :] = 0.0
    w[2, : cols // 2] = 1e-8
    w[3, 0] = 3.0e38
  return w


def cfg(bits, sym, gran_key, **params):
  return Q.TensorQuantizationConfig(num_bits=bits, symmetric=sym, granularity=GRAN[gran_key],
                                    algorithm_params=params)


def run(mod, w, c, qsv=None, op="FULLY_CONNECTED"):
  with np.errstate(all="ignore"):
    return mod.get_tensor_quant_params(refshim.fc_op_info(c, op), c, w, qsv)


def save(name, **arrays):
  np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
  print("wrote", name, len(arrays), "arrays")


def main():
  only = set(sys.argv[1:])
  for name, fn in (("minmax", gen_minmax), ("octav", gen_octav), ("mse", gen_mse),
                   ("hadamard", gen_hadamard), ("gptq", gen_gptq),
                   ("calibration", gen_calibration), ("pack", gen_pack),
                   ("histogram", gen_histogram), ("recovery", gen_recovery), ("oscar", gen_oscar)):
    if not only or name in only:
      fn()


def gen_minmax():
  # ---- min-max: every granularity / bit width / symmetry the policy allows
  out = {}
  cases = []
  for i, (rows, cols) in enumerate([(16, 8), (12, 256), (8, 4096), (5, 11008)]):
    w = weight(rows, cols, 10 + i)
    out[f"w{i}"] = w
    for bits in (2, 4, 8):
      for sym in (True, False):
        for gk in (0, -1, 32, 64, 128, 256):
          if gk > 0 and (not sym or cols % gk):
            continue
          r = run(NMM, w, cfg(bits, sym, gk))
          key = f"w{i}_b{bits}_s{int(sym)}_g{gk}"
          out[key + "_q"], out[key + "_scale"], out[key + "_zp"] = r.quantized_data, r.scale, r.zero_point
          cases.append(key)
  out["cases"] = np.array(cases)
  save("minmax", **out)



def gen_octav():
  # ---- OCTAV (no literal goldens exist in the reference tests)
  out, cases = {}, []
  for i, (rows, cols) in enumerate([(16, 256), (8, 4096), (6, 1024)]):
    w = weight(rows, cols, 20 + i, special=False)
    out[f"w{i}"] = w
    for bits in (4, 8):
      for gk in (0, -1, 32, 128):
        if gk > 0 and cols % gk:
          continue
        r = run(OCT, w, cfg(bits, True, gk))
        key = f"w{i}_b{bits}_g{gk}"
        out[key + "_q"], out[key + "_scale"] = r.quantized_data, r.scale
        if gk > 0:
          x = w.reshape(rows, cols // gk, gk)
          clip = OCT._guess_clipping_with_octav(x, bits, 2, 10, 3.0, True)
        elif gk == 0:
          clip = OCT._guess_clipping_with_octav(w, bits, (1,), 10, 3.0, True)
        else:
          clip = OCT._guess_clipping_with_octav(w, bits, (0, 1), 10, 3.0, True)
        out[key + "_clip"] = np.asarray(clip)
        cases.append(key)
  out["cases"] = np.array(cases)
  save("octav", **out)



def gen_mse():
  # ---- MSE
  out, cases = {}, []
  for i, (rows, cols) in enumerate([(16, 256), (8, 4096)]):
    w = weight(rows, cols, 30 + i, special=False)
    out[f"w{i}"] = w
    for bits in (4, 8):
      for gk in (0, -1):
        r = run(MSE, w, cfg(bits, True, gk))
        key = f"w{i}_b{bits}_g{gk}"
        out[key + "_q"], out[key + "_scale"] = r.quantized_data, r.scale
        cases.append(key)
  out["cases"] = np.array(cases)
  save("mse", **out)



def gen_hadamard():
  # ---- Hadamard rotation (+ OCTAV)
  out, cases = {}, []
  for i, (rows, cols, cap) in enumerate([(8, 256, None), (6, 1024, 64), (4, 4096, None), (4, 2752, None)]):
    w = weight(rows, cols, 40 + i, special=False)
    out[f"w{i}"] = w
    params = {} if cap is None else {"max_hadamard_size": cap}
    for bits in (4, 8):
      c = cfg(bits, True, 0, **params)
      r = run(HAD, w, c)
      key = f"w{i}_b{bits}"
      out[key + "_q"], out[key + "_scale"] = r.quantized_data, r.scale
      out[key + "_hsize"] = np.array(r.hadamard.hadamard_size)
      cases.append(key)
    out[f"w{i}_rot"] = HAD._rotate_with_diagonal_hadamard(w, 1, cap)[0]
    out[f"w{i}_cap"] = np.array(-1 if cap is None else cap)
  out["cases"] = np.array(cases)
  save("hadamard", **out)



def gen_gptq():
  # ---- GPTQ (Hessian from seeded activations; small K so the CPU loop is quick)
  out, cases = {}, []
  for i, (rows, k, tokens) in enumerate([(16, 128, 256), (8, 256, 512), (12, 192, 300)]):
    w = weight(rows, k, 50 + i, special=False)
    x = O.synthetic_activation((4, tokens // 4, k), 50 + i)
    hess = (2.0 / np.array(x.shape[0])) * x.reshape(-1, k).T.dot(x.reshape(-1, k))
    # _prepare_hessian_inverse leaves its damped diagonal in the caller's array
    # (np.diag returns a view, so the "restore" at gptq.py:123 is a no-op): hand
    # every call its own copy and keep the pristine Hessian in the fixture.
    out[f"w{i}"], out[f"x{i}"], out[f"h{i}"] = w, x, hess.copy()
    out[f"hinv{i}"] = GPTQ._prepare_hessian_inverse(hess.copy())
    mutated = hess.copy()
    GPTQ._prepare_hessian_inverse(mutated)
    out[f"hdiag_after{i}"] = np.diag(mutated).copy()
    for bits in (4, 8):
      for gk in (0, 32):
        if gk and k % gk:
          continue
        qsv = {"activation_tensor_qsv": {"hessian": hess.copy(), "num_samples": x.shape[0]}}
        r = run(GPTQ, w, cfg(bits, True, gk), qsv)
        key = f"w{i}_b{bits}_g{gk}"
        out[key + "_q"], out[key + "_scale"] = r.quantized_data, r.scale
        cases.append(key)
  out["cases"] = np.array(cases)
  save("gptq", **out)



def gen_calibration():
  # ---- activation calibration filter + EMA over a batch sequence
  out = {}
  acts = [O.synthetic_activation((4, 33, 65), j) for j in range(6)]
  acts[1][0, 0, 0] = np.inf
  acts[2][1, 2, 3] = -np.inf
  acts[3][0, 1, 1] = 3.39e38
  acts[4][:] = 3.2e38  # everything filtered -> raw fallback
  q = {}
  for j, a in enumerate(acts):
    mm = CQ.get_activation_min_max(a, -3e38, 3e38)
    out[f"a{j}"], out[f"a{j}_min"], out[f"a{j}_max"] = a, mm["min"], mm["max"]
    q = QSV.moving_average_update(q, mm)
  out["ema_min"], out["ema_max"] = q["min"], q["max"]
  save("calibration", **out)



def gen_recovery():
  """dequantized_weight_recovery on QAT-style weights (every granularity) + the fp16 cast."""
  DWR = refshim.ref("algorithms.uniform_quantize.dequantized_weight_recovery")
  out = {}
  cases = []
  for i, (bits, gk, shape) in enumerate([(4, 0, (16, 512)), (8, 0, (8, 11008)), (4, 32, (16, 512)),
                                         (8, 256, (8, 1024)), (4, -1, (16, 512)), (8, -1, (32, 1024))]):
    w = O.fake_quantized_weight(shape[0], shape[1], bits, block=max(gk, 0), index=i, per_channel=(gk == 0))
    w[1, :] = 0.0
    w[2, :] = w[2, 0]
    r = run(DWR, w, cfg(bits, True, gk))
    out[f"w{i}"], out[f"scale{i}"], out[f"q{i}"] = w, r.scale, r.quantized_data
    cases.append((bits, gk))
  out["cases"] = np.array(cases)
  w = weight(32, 300, 77)
  w[5, 0], w[5, 1], w[5, 2] = 70000.0, -65520.0, 6e-8  # overflow to inf, the rounding boundary, a subnormal
  out["cast_in"], out["cast_out"] = w, w.astype(np.float16)
  save("recovery", **out)


def oscar_mu2(d, seed):
  rng = np.random.default_rng(seed)
  mu2 = rng.standard_normal(d) ** 2 * 0.5 + 0.01
  mu2[::17] *= 40.0
  mu2[3] = 0.0
  return mu2


def gen_oscar():
  """OSCAR end to end (channel scales, clip bounds, scale, integers) + calibrate's mu2."""
  OSC = refshim.ref("algorithms.uniform_quantize.oscar")
  out, cases = {}, []
  specs = [(4, 0, (16, 512), True), (8, 0, (16, 512), True), (4, 32, (16, 512), True),
           (4, 128, (16, 512), False), (4, -1, (64, 512), True), (8, 0, (6, 11008), False),
           (4, 0, (5, 100), True)]
  for i, (bits, gk, shape, with_mu2) in enumerate(specs):
    w = weight(shape[0], shape[1], 40 + i, special=False)
    w[1, :] = 0.0
    mu2 = oscar_mu2(shape[1], i) if with_mu2 else None
    r = run(OSC, w, cfg(bits, True, gk), qsv=({"mu2": mu2} if with_mu2 else None))
    out[f"w{i}"], out[f"scale{i}"], out[f"q{i}"] = w, r.scale, r.quantized_data
    out[f"mult{i}"] = r.custom_algorithm_param["multiplier"]
    if with_mu2:
      out[f"mu2_{i}"] = mu2
    cases.append((bits, gk, int(with_mu2)))
  out["cases"] = np.array(cases)
  x = O.synthetic_activation((3, 70, 96), 8)
  out["act"], out["act_mu2"] = x, np.mean(np.asarray(x, np.float64).reshape(-1, 96) ** 2, axis=0)
  save("oscar", **out)


def gen_pack():
  # ---- pack_data
  out = {}
  rng = np.random.default_rng(7)
  for bits, n in ((4, 15), (4, 4096), (2, 13), (2, 1024)):
    half = 2 ** (bits - 1)
    v = rng.integers(-half, half, size=n, dtype=np.int8)
    out[f"b{bits}_n{n}_in"] = v
    out[f"b{bits}_n{n}_out"] = TU.pack_data(bits, v.view(np.uint8).copy())
  save("pack", **out)


def gen_histogram():
  # ---- DynamicHistogram (utils/histogram_utils.py): per-tensor growth / compaction, per-channel,
  # initial_bin_width, non-finite filtering and a resampled merge
  HU = refshim.ref("utils.histogram_utils")
  out = {}
  rng = np.random.default_rng(77)
  batches = [rng.standard_normal((4, 33, 65)).astype(np.float32) * s for s in (1.0, 0.5, 3.0, 40.0, 2.0)]
  batches[1][0, 0, 0] = np.inf
  batches[2][1, 2, 3] = np.nan
  batches[2][1, 2, 4] = -np.inf
  for j, b in enumerate(batches):
    out[f"b{j}"] = b

  def dump(prefix, h):
    for c, impl in enumerate(h._impls):
      out[f"{prefix}_c{c}_counts"] = np.array(impl.counts, copy=True)  # add() updates in place
      out[f"{prefix}_c{c}_lb"] = np.array(impl.lower_bound)
      out[f"{prefix}_c{c}_bw"] = np.array(impl.bin_width)
      out[f"{prefix}_c{c}_min"] = np.array(impl.global_min)
      out[f"{prefix}_c{c}_max"] = np.array(impl.global_max)

  h = HU.DynamicHistogram(max_tensor_bins=2048)
  for j, b in enumerate(batches):
    h.add(b)
    dump(f"t_after{j}", h)
  h2 = HU.DynamicHistogram(max_tensor_bins=256, initial_bin_width=0.05)
  for b in batches[:3]:
    h2.add(b)
  dump("w", h2)
  hc = HU.DynamicHistogram(max_tensor_bins=2048, axis=0)
  for b in batches[:4]:
    hc.add(b)
  dump("ch", hc)
  a, bb = HU.DynamicHistogram(max_tensor_bins=512), HU.DynamicHistogram(max_tensor_bins=512)
  a.add(batches[0]); a.add(batches[2])
  bb.add(batches[3])
  a.merge(bb)
  dump("merged", a)
  save("histogram", **out)


if __name__ == "__main__":
  main()
