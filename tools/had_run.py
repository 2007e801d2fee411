"""Profiling driver: a few Hadamard rotations of one [4096, 4096] tensor.  python tools/had_run.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")]
import torch  # noqa: E402

from aeq_b200 import device  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
ws = [torch.randn(4096, 4096, device=dev) for _ in range(4)]
out = torch.empty_like(ws[0])
for _ in range(3):
  for w in ws:
    device.hadamard_rows(w, n, out=out)
torch.cuda.synchronize()
