"""Sustained series of the headline launch (64- and 477-tensor sets): algorithmic HBM GB/s per window of launches,
first and last four windows.  Run once per variant (environment switches of the rows kernel) to A/B them:
  for v in "" AEQB_ROWS_NO_FOLD=1; do env $v python tools/sustain_ab.py; done
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")):
  sys.path.insert(0, p)
import torch  # noqa: E402
from aeq_b200 import device  # noqa: E402

dev = torch.device("cuda:0")
n_big = int(os.environ.get("SUSTAIN_TENSORS", "477"))
flat = torch.empty(n_big * 4096 * 4096, device=dev)
flat.normal_(0.0, 0.02, generator=torch.Generator(device=dev).manual_seed(1))
ws = list(flat.view(n_big, 4096, 4096).unbind(0))
print({k: v for k, v in os.environ.items() if k.startswith("AEQB_")})
kind = os.environ.get("SUSTAIN_KIND", "rows")  # rows: INT8 per-channel; blocks: INT4 block-32 packed + fp16 scales
for name, n, per_win, wins in ((f"{kind}64", 64, 20, 40), (f"{kind}{n_big}", n_big, 3, 40)):
  w = ws[:n]
  if kind == "rows":
    outs = device.requant_rows_batch(w, 8, True)
    fn = lambda: device.requant_rows_batch(w, 8, True, outs=outs)
    bpw = 5.0 + 8.0 / 4096
  else:
    outs = device.requant_blocks_batch(w, 32, 4)
    fn = lambda: device.requant_blocks_batch(w, 32, 4, outs=outs)
    bpw = 4.5625
  fn()
  torch.cuda.synchronize()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(wins + 1)]
  ev[0].record()
  for k in range(wins):
    for _ in range(per_win):
      fn()
    ev[k + 1].record()
  torch.cuda.synchronize()
  nbytes = n * 4096 * 4096 * bpw
  gbs = [round(nbytes * per_win / ev[k].elapsed_time(ev[k + 1]) / 1e6) for k in range(wins)]
  print(f"{name:10s} alg GB/s first {gbs[:4]} ... last {gbs[-4:]}  mean {sum(gbs) / len(gbs):.0f}")
  del outs
  torch.cuda.synchronize()
