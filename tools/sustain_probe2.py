"""Sustained behaviour of the headline kernel per stream class, with 20 ms clock / power samples."""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")):
  sys.path.insert(0, p)
import torch  # noqa: E402
from aeq_b200 import device  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
ws = [torch.randn(4096, 4096, device=dev, generator=g) * 0.02 for _ in range(64)]
outs = device.requant_rows_batch(ws, 8, True)
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader,nounits",
                         "-lms", "20", "-i", "0"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.perf_counter(), l.strip())) for l in proc.stdout], daemon=True).start()
time.sleep(0.3)
mode = sys.argv[1] if len(sys.argv) > 1 else "rows"
src = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
dst = torch.empty_like(src)
fn = (lambda: device.requant_rows_batch(ws, 8, True, outs=outs)) if mode == "rows" else (lambda: dst.copy_(src))
nbytes = 64 * 4096 * 4096 * 5 if mode == "rows" else 4 * (1 << 30)
fn()
torch.cuda.synchronize()
n_win = 40
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_win + 1)]
t_start = time.perf_counter()
ev[0].record()
for k in range(n_win):
  for _ in range(20):
    fn()
  ev[k + 1].record()
torch.cuda.synchronize()
t_end = time.perf_counter()
gbs = [round(nbytes * 20 / ev[k].elapsed_time(ev[k + 1]) / 1e6) for k in range(n_win)]
time.sleep(0.1)
proc.terminate()
samples = [l for t, l in rows if t_start - 0.05 <= t <= t_end + 0.05]
print(json.dumps({"mode": mode, "class_env": os.environ.get("AEQB_ROWS_MIN_CLASS"), "gbs": gbs, "smi_sm_mem_power": samples}))
