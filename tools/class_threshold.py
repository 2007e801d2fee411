"""Where does the 96 KiB-tile class start to pay?  Times n x [4096,4096] INT8 per-channel batches under the
class rule of this process's environment (run once with AEQB_ROWS_SMALL_TILES=1, once with AEQB_ROWS_MIN_CLASS=4)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")):
  sys.path.insert(0, p)
import torch  # noqa: E402
from aeq_b200 import device  # noqa: E402

dev = torch.device("cuda:0")
ws = [torch.randn(4096, 4096, device=dev) * 0.02 for _ in range(16)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for n in (1, 2, 3, 4, 6, 8, 12, 16):
  w = ws[:n]
  outs = device.requant_rows_batch(w, 8, True)
  for _ in range(3):
    device.requant_rows_batch(w, 8, True, outs=outs)
  torch.cuda.synchronize()
  ts = []
  for _ in range(12):
    flush.zero_()  # L2 flush between launches (inputs of <= 126 MB would otherwise be re-read from L2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    device.requant_rows_batch(w, 8, True, outs=outs)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
  ts.sort()
  out.append(f"{n * 64} MiB: {ts[len(ts) // 2]:.1f} us")
print({k: v for k, v in os.environ.items() if k.startswith("AEQB_ROWS")}, " | ".join(out))
