"""Kernel micro-benchmark used while tuning (not the contract bench; see bench.py).

python tools/kbench.py [--rows 4096 --cols 4096 --tensors 64 --iters 20]
Prints one line per kernel variant: ms per pass over the stack, fp32-weight GB/s,
algorithmic HBM GB/s and fraction of MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")]
import torch  # noqa: E402

from aeq_b200 import _lib  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--rows", type=int, default=4096)
  ap.add_argument("--cols", type=int, default=4096)
  ap.add_argument("--tensors", type=int, default=64)
  ap.add_argument("--iters", type=int, default=20)
  ap.add_argument("--only", default="")
  a = ap.parse_args()
  peak = 6558.1
  try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
  except Exception:
    pass
  dev = torch.device("cuda:0")
  R, C, T = a.rows, a.cols, a.tensors
  ws = [torch.randn(R, C, device=dev) * 0.02 for _ in range(T)]
  q = [torch.empty(R, C, dtype=torch.int8, device=dev) for _ in range(T)]
  pk = [torch.empty(R * C // 2, dtype=torch.uint8, device=dev) for _ in range(T)]
  sc = [torch.empty(R, dtype=torch.float32, device=dev) for _ in range(T)]
  zp = [torch.empty(R, dtype=torch.int32, device=dev) for _ in range(T)]
  bs = [torch.empty(R * C // 32, dtype=torch.float32, device=dev) for _ in range(T)]
  bh = [torch.empty(R * C // 32, dtype=torch.float16, device=dev) for _ in range(T)]
  st = torch.cuda.current_stream().cuda_stream
  lib = _lib.load()

  def rows8():
    for i in range(T):
      lib.aeqb_requant_rows_f32(ws[i].data_ptr(), R, C, 8, 1, None, q[i].data_ptr(), None,
                                sc[i].data_ptr(), zp[i].data_ptr(), st)

  def rows4p():
    for i in range(T):
      lib.aeqb_requant_rows_f32(ws[i].data_ptr(), R, C, 4, 1, None, None, pk[i].data_ptr(),
                                sc[i].data_ptr(), zp[i].data_ptr(), st)

  def blk4p():
    for i in range(T):
      lib.aeqb_requant_blocks_f32(ws[i].data_ptr(), R, C, 32, 4, None, None, pk[i].data_ptr(),
                                  None, bh[i].data_ptr(), st)

  def blk4q():
    for i in range(T):
      lib.aeqb_requant_blocks_f32(ws[i].data_ptr(), R, C, 32, 4, None, q[i].data_ptr(), None,
                                  bs[i].data_ptr(), None, st)

  from aeq_b200 import device
  state = {}

  def rows8_batch():
    state["r8"] = device.requant_rows_batch(ws, 8, True, outs=state.get("r8"))

  def blk4p_batch():
    state["b4"] = device.requant_blocks_batch(ws, 32, 4, outs=state.get("b4"))

  def copy():
    for i in range(T):
      q[i].view(torch.float32).copy_(ws[i].view(-1)[: R * C // 4].view(R, C // 4))

  n = R * C * T
  cases = [("rows_int8", rows8, 5.0), ("rows_int4_packed", rows4p, 4.5),
           ("blocks32_int4_packed", blk4p, 4.5625), ("blocks32_int4_unpacked", blk4q, 5.125),
           ("rows_int8_batch", rows8_batch, 5.0), ("blocks32_int4_packed_batch", blk4p_batch, 4.5625)]
  for name, fn, bpw in cases:
    if a.only and name not in a.only.split(","):
      continue
    for _ in range(3):
      fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
      fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    wgb = n * 4 / ms / 1e6
    hbm = n * bpw / ms / 1e6
    print(f"{name:26s} {ms:8.3f} ms/pass  weight {wgb:8.1f} GB/s  hbm {hbm:8.1f} GB/s  frac {hbm/peak:.3f}",
          flush=True)


if __name__ == "__main__":
  main()
