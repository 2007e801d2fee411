"""Per-shape timing of the batched fused requantisation (device-resident), to find the shapes
that fall off the roofline.  python tools/shape_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")]
import torch  # noqa: E402

from aeq_b200 import device  # noqa: E402


def main():
  peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
  dev = torch.device("cuda:0")
  shapes = [(4096, 4096), (11008, 4096), (4096, 11008), (2048, 2048), (256, 2048), (16384, 2048),
            (2048, 16384), (4096, 8192), (4096, 14336), (4096, 5120), (4096, 6144), (4096, 3584),
            (4096, 2560), (8192, 1536), (4096, 13824)]
  for r, c in shapes:
    n = max(2, int(2e9 // (r * c * 4)))  # ~2 GB per stack: far beyond L2
    ws = [torch.randn(r, c, device=dev) * 0.02 for _ in range(n)]
    for name, fn, bpw in (
        ("rows int8", lambda st: device.requant_rows_batch(ws, 8, True, outs=st.get("o")), 5.0),
        ("rows int4 packed", lambda st: device.requant_rows_batch(ws, 4, True, want_q=False, want_packed=True,
                                                                  outs=st.get("o")), 4.5),
        ("blocks32 int4 packed", lambda st: device.requant_blocks_batch(ws, 32, 4, outs=st.get("o")), 4.5625)):
      st = {}
      for _ in range(3):
        st["o"] = fn(st)
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      reps = 10
      for _ in range(reps):
        st["o"] = fn(st)
      e1.record()
      torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / reps
      gbs = n * r * c * bpw / ms / 1e6
      print(f"[{r:6d},{c:6d}] x{n:3d} {name:22s} {ms:7.3f} ms  {gbs:7.1f} GB/s  frac {gbs / peak:.3f}", flush=True)
    del ws


if __name__ == "__main__":
  main()
