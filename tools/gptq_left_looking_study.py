"""Feasibility study for DESIGN.md §8 item 1 (CPU, NumPy; test infrastructure, not product).

The reference's OBS loop is RIGHT-looking (gptq.py:131-216): after each 64-column block it
subtracts `err_b @ hinv[b, end:]` from every later column, in fp32, block after block.  A
LEFT-looking schedule applies all pending updates to block b just before it is quantised:
    W[:, b] -= sum_{p < b} err_p @ hinv[p, b]
which is one long-contraction product per block (what a tensor-core kernel wants) and writes every
column once.  The two are the same in exact arithmetic; in fp32 the summation order differs, and a
flipped rounding decision propagates through its row.  This script measures how far the integers
and the proxy loss tr(E H E^T) move, for the accumulation orders a kernel could use:
  seq   : subtract the 64-wide partial products one after the other in fp32 (the reference's order
          per column, only the loop nest is interchanged: bit-identical by construction)
  gemm  : one fp32 GEMM over the whole contraction, then one subtraction
  gemm64: the same with a float64 accumulator (an upper bound on what 3xTF32 + segmented
          fp32 drains can deliver)
Usage: python tools/gptq_left_looking_study.py [rows cols tokens]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import aeq_oracle as O  # noqa: E402


def left_looking(w, scale, zp, hinv, bits, mode, blocksize=64):
  fp = w.copy()
  r, k = fp.shape
  q_all = np.zeros((r, k), dtype=O.qdtype(bits))
  err_all = np.zeros((r, k), dtype=np.float32)
  sc, z = scale.reshape(-1), zp.reshape(-1)
  for b0 in range(0, k, blocksize):
    b1 = min(b0 + blocksize, k)
    if b0:
      if mode == "seq":
        for p0 in range(0, b0, blocksize):
          fp[:, b0:b1] -= np.matmul(err_all[:, p0:p0 + blocksize], hinv[p0:p0 + blocksize, b0:b1])
      elif mode == "gemm":
        fp[:, b0:b1] -= np.matmul(err_all[:, :b0], hinv[:b0, b0:b1])
      else:
        fp[:, b0:b1] -= np.matmul(err_all[:, :b0].astype(np.float64),
                                  hinv[:b0, b0:b1].astype(np.float64)).astype(np.float32)
    wb = fp[:, b0:b1]
    for i in range(b1 - b0):
      col = b0 + i
      wc = wb[:, i]
      qc = O.quantize(wc[:, None], sc[:, None], z[:, None], bits, True)
      dq = np.multiply(qc - z[:, None], sc[:, None]).reshape(-1)
      q_all[:, col] = qc.reshape(-1)
      e = (wc - dq) / hinv[col, col]
      err_all[:, col] = e
      if i < b1 - b0 - 1:
        wb[:, i + 1:] -= np.outer(e, hinv[col, col + 1:b1])
  return q_all


def proxy_loss(w, q, scale, h):
  e = (w - q.astype(np.float32) * scale).astype(np.float64)
  return float(np.einsum("ij,jk,ik->", e, h.astype(np.float64), e))


def main():
  rows, cols, tokens = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 1024, 4096)
  w = O.synthetic_weight(rows, cols, 5)
  x = O.synthetic_activation((4, tokens // 4, cols), 5)
  if os.environ.get("CORRELATED", "1") != "0":
    # i.i.d. activations give a near-diagonal Hessian (GPTQ ~ round to nearest, nothing to
    # propagate); real layers see strongly correlated features: mix through a random low-rank map
    rng = np.random.default_rng(11)
    mix = (rng.standard_normal((cols, cols // 8), dtype=np.float32) @
           rng.standard_normal((cols // 8, cols), dtype=np.float32)) / np.float32(cols // 8) ** 0.5
    x = (x.reshape(-1, cols) @ (0.3 * np.eye(cols, dtype=np.float32) + mix)).reshape(4, -1, cols)
  h = O.gptq_hessian(x)
  ref = O.gptq_requant(w, h, 4)
  hinv = ref["hinv"].astype(np.float32)
  scale, zp = ref["scale"], ref["zero_point"]
  base = O.gptq_quantize(w, scale, zp, hinv, 4)
  rtn = O.quantize(w, scale, zp, 4, True)
  l0 = proxy_loss(w, base, scale, h)
  print(f"[{rows},{cols}] INT4 per-channel, {tokens} tokens; round-to-nearest proxy loss / GPTQ = "
        f"{proxy_loss(w, rtn, scale, h) / l0:.3f}")
  for mode in ("seq", "gemm", "gemm64"):
    q = left_looking(w, scale, zp, hinv, 4, mode)
    mism = float(np.mean(q != base))
    print(f"  left-looking {mode:6s}: integers differing {mism:.4%}, max |dq| {int(np.abs(q.astype(int) - base.astype(int)).max())}, "
          f"proxy loss ratio {proxy_loss(w, q, scale, h) / l0:.6f}")


if __name__ == "__main__":
  main()
