import json, os, sys
ROOT = os.getcwd()
sys.path[:0] = [ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")]
import torch
from aeq_b200 import device
peak = 6549.1
dev = torch.device("cuda:0")
for r, c in [(2048, 2048), (256, 2048), (4096, 4096)]:
    n = max(2, int(2e9 // (r * c * 4)))
    ws = [torch.randn(r, c, device=dev) * 0.02 for _ in range(n)]
    for name, fn, bpw in (
        ("rows int8", lambda st: device.requant_rows_batch(ws, 8, True, outs=st.get("o")), 5.0),
        ("rows int4 packed", lambda st: device.requant_rows_batch(ws, 4, True, want_q=False, want_packed=True, outs=st.get("o")), 4.5),
        ("blocks32 int4 packed", lambda st: device.requant_blocks_batch(ws, 32, 4, outs=st.get("o")), 4.5625)):
      for trial in range(3):
        st = {}
        for _ in range(3):
            st["o"] = fn(st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            st["o"] = fn(st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"[{r},{c}] x{n} {name:22s} trial {trial} {ms:7.3f} ms frac {n*r*c*bpw/ms/1e6/peak:.3f}", flush=True)
    del ws
