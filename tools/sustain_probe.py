"""Is the 477-tensor step slower per byte than the 64-tensor one because of its footprint or
because it runs long enough for the power cap to bite?  Times (a) the 64-tensor launch repeated
for ~0.5 s, in windows, (b) a plain device copy over the same windows, (c) the 477-tensor launch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")):
  sys.path.insert(0, p)
import torch  # noqa: E402
from aeq_b200 import device  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
T = int(os.environ.get("T", "477"))
ws = [torch.randn(4096, 4096, device=dev, generator=g) * 0.02 for _ in range(T)]
outs = device.requant_rows_batch(ws, 8, True)
o64 = outs[:64]
w64 = ws[:64]


def windows(fn, reps_per_window, n_windows, bytes_per_rep):
  fn()
  torch.cuda.synchronize()
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_windows + 1)]
  ev[0].record()
  for k in range(n_windows):
    for _ in range(reps_per_window):
      fn()
    ev[k + 1].record()
  torch.cuda.synchronize()
  return [round(bytes_per_rep * reps_per_window / ev[k].elapsed_time(ev[k + 1]) / 1e6, 1) for k in range(n_windows)]


res = {}
b64 = 64 * 4096 * 4096 * 5
res["rows64_alg_gbs_windows_of_20_launches"] = windows(lambda: device.requant_rows_batch(w64, 8, True, outs=o64), 20, 30, b64)
src = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
dst = torch.empty_like(src)
res["torch_copy_gbs_windows_of_20_copies"] = windows(lambda: dst.copy_(src), 20, 30, 2 * src.numel() * 2)
res["rows477_alg_gbs_per_step"] = windows(lambda: device.requant_rows_batch(ws, 8, True, outs=outs), 1, 30, T * 4096 * 4096 * 5)
# same 64 tensors, but spread over the whole 40 GB footprint (every 7th tensor)
wsp = ws[::7][:64]
osp = outs[::7][:64]
res["rows64_spread_alg_gbs_windows_of_20_launches"] = windows(lambda: device.requant_rows_batch(wsp, 8, True, outs=osp), 20, 10, len(wsp) * 4096 * 4096 * 5)
print(json.dumps(res))
