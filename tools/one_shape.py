"""Runs the batched per-channel requantisation on one shape a few times (profiling driver).
python tools/one_shape.py ROWS COLS [BITS] [PACKED]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")]
import torch  # noqa: E402

from aeq_b200 import device  # noqa: E402

r, c = int(sys.argv[1]), int(sys.argv[2])
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 8
packed = len(sys.argv) > 4 and sys.argv[4] == "1"
dev = torch.device("cuda:0")
ws = [torch.randn(r, c, device=dev) * 0.02 for _ in range(8)]
outs = None
for _ in range(3):
  outs = device.requant_rows_batch(ws, bits, True, want_q=not packed, want_packed=packed, outs=outs)
torch.cuda.synchronize()
