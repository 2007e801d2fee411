"""Debug/timing tool for the Hessian contraction (aeqb_xtx_f32): error against an fp64 torch
product (checker only) and time per call.  AEQB_XTX_SIMT=1 selects the SIMT kernel.

  python tools/xtx_check.py [--time]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")]
import torch  # noqa: E402

from aeq_b200 import device  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--time", action="store_true")
  ap.add_argument("--shapes", default="4096x512,4100x1028,2048x256,20000x2048,16384x4096")
  a = ap.parse_args()
  dev = torch.device("cuda:0")
  mode = "simt" if os.environ.get("AEQB_XTX_SIMT") == "1" else "tcgen05"
  for s in a.shapes.split(","):
    T, K = (int(v) for v in s.split("x"))
    g = torch.Generator(device=dev).manual_seed(T + K)
    x = torch.randn(T, K, device=dev, generator=g) * 0.5 + 0.1
    h = device.xtx(x, 2.0 / 7)
    torch.cuda.synchronize()
    ref = (2.0 / 7) * (x.double().T @ x.double())
    err = (h - ref).abs().max().item()
    scale = ref.diagonal().abs().max().item()
    sym = (h - h.T).abs().max().item()
    line = f"{mode} T={T} K={K}: max|err|/max diag = {err / scale:.3e}  asym={sym:.1e}"
    if a.time:
      for _ in range(2):
        device.xtx(x, 1.0)
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      n = 5
      for _ in range(n):
        device.xtx(x, 1.0)
      e1.record()
      torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / n
      line += f"  {ms:.3f} ms  {2.0 * T * K * K / ms / 1e9:.1f} TFLOP/s (full square)"
    print(line, flush=True)


if __name__ == "__main__":
  main()
