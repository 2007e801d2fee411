"""Per-kernel timings through the raw C ABI with preallocated buffers (CUDA events around
back-to-back calls, so the GPU — not Python — is the bottleneck for anything above ~10 us).

  python tools/ktime.py [--only substr] [--iters 20]
Prints: name, us per call, algorithmic GB/s, fraction of MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")]
import torch  # noqa: E402

from aeq_b200 import _lib  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--only", default="")
  ap.add_argument("--iters", type=int, default=20)
  ap.add_argument("--rows", type=int, default=4096)
  ap.add_argument("--cols", type=int, default=4096)
  a = ap.parse_args()
  peak = 6558.1
  try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
  except Exception:
    pass
  lib = _lib.load()
  dev = torch.device("cuda:0")
  R, C = a.rows, a.cols
  n = R * C
  NT = 8  # distinct input tensors, cycled so that L2 (126 MB) does not serve the reads
  ws_in = []
  for i in range(NT):
    w = torch.randn(R, C, device=dev) * 0.02
    w.view(-1)[::1024] *= 20.0
    ws_in.append(w)
  st = torch.cuda.current_stream().cuda_stream
  f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
  q = torch.empty(n, dtype=torch.int8, device=dev)
  pk = torch.empty(n // 2, dtype=torch.uint8, device=dev)
  # valid parameters (not torch.empty garbage: degenerate scales take the IEEE-divide path)
  scale, zp = torch.full((R,), 1e-3, device=dev), torch.zeros(R, dtype=torch.int32, device=dev)
  clip_r, clip_b = torch.full((R,), 0.05, device=dev), torch.full((n // 32,), 0.05, device=dev)
  bscale = torch.empty(n // 32, dtype=torch.float16, device=dev)
  out2 = f32(NT, 2)
  rot = f32(R, C)
  mm_ws = torch.zeros(lib.aeqb_minmax_workspace_bytes(), dtype=torch.uint8, device=dev)
  oct_ws_r = torch.empty(lib.aeqb_octav_workspace_bytes(R, 10), dtype=torch.uint8, device=dev)
  oct_ws_b = torch.empty(lib.aeqb_octav_workspace_bytes(n // 32, 10), dtype=torch.uint8, device=dev)
  P = lambda t: t.data_ptr()

  jobs = (_lib.MinmaxJob * NT)()
  for i, w in enumerate(ws_in):
    jobs[i] = _lib.MinmaxJob(P(w), n, out2[i].data_ptr())
  jobs_p = ctypes.cast(jobs, ctypes.c_void_p)

  cases = []  # (name, fn(i), bytes per call)
  cases.append(("minmax_tensor 64MiB", lambda i: lib.aeqb_minmax_tensor_f32(
      P(ws_in[i % NT]), n, -3e38, 3e38, 1, 1, P(out2), P(mm_ws), st), n * 4))
  cases.append((f"minmax_tensors batch of {NT}", lambda i: lib.aeqb_minmax_tensors_f32(
      jobs_p, NT, -3e38, 3e38, 1, 1, P(mm_ws), st), NT * n * 4))
  for bits in (4, 8):
    cases.append((f"octav_clip_rows b{bits}", lambda i, b=bits: lib.aeqb_octav_clip_rows_f32(
        P(ws_in[i % NT]), R, C, b, 10, 3.0, 1, P(clip_r), P(oct_ws_r), st), n * 4))
  cases.append(("octav_clip_blocks32 b4", lambda i: lib.aeqb_octav_clip_blocks_f32(
      P(ws_in[i % NT]), R, C, 32, 4, 10, 3.0, 1, P(clip_b), P(oct_ws_b), st), n * 4 + 11 * n // 8))
  cases.append(("requant_rows int8 +clip", lambda i: lib.aeqb_requant_rows_f32(
      P(ws_in[i % NT]), R, C, 8, 1, P(clip_r), P(q), None, P(scale), P(zp), st), n * 5))
  cases.append(("requant_rows int8", lambda i: lib.aeqb_requant_rows_f32(
      P(ws_in[i % NT]), R, C, 8, 1, None, P(q), None, P(scale), P(zp), st), n * 5))
  cases.append(("requant_blocks32 int4 packed +clip", lambda i: lib.aeqb_requant_blocks_f32(
      P(ws_in[i % NT]), R, C, 32, 4, P(clip_b), None, P(pk), None, P(bscale), st), n * 4.5625))
  cases.append(("mse_scale_rows", lambda i: lib.aeqb_mse_scale_rows_f32(
      P(ws_in[i % NT]), R, C, 0.05408, P(scale), None, st), n * 4))
  cases.append(("quantize (given scale) int8", lambda i: lib.aeqb_quantize_f32(
      P(ws_in[i % NT]), n, R, C, P(scale), None, 1, 8, 1, P(q), st), n * 5))
  cases.append((f"hadamard_rows n={C}", lambda i: lib.aeqb_hadamard_rows_f32(
      P(ws_in[i % NT]), R, C, C, P(rot), st), n * 8))
  cases.append(("hadamard_rows n=128", lambda i: lib.aeqb_hadamard_rows_f32(
      P(ws_in[i % NT]), R, C, 128, P(rot), st), n * 8))
  cases.append(("row_stats minmax", lambda i: lib.aeqb_row_stats_f32(
      P(ws_in[i % NT]), R, C, P(scale), P(clip_r), None, st), n * 4))

  # GPTQ pieces on one [R, C] layer
  if not a.only or "gptq" in a.only or "xtx" in a.only or "hess" in a.only:
    x = torch.randn(8192, C, device=dev)
    h = torch.empty(C, C, dtype=torch.float64, device=dev)
    xws_n = lib.aeqb_xtx_workspace_bytes(8192, C)
    xws = torch.empty(max(xws_n, 1), dtype=torch.uint8, device=dev)
    lib.aeqb_xtx_f32(P(x), 8192, C, 0.25, P(h), P(xws) if xws_n else None, st)
    hinv = f32(C, C)
    hws = torch.empty(lib.aeqb_hessian_inverse_workspace_bytes(C), dtype=torch.uint8, device=dev)
    info = torch.zeros(1, dtype=torch.int32, device=dev)
    lib.aeqb_hessian_inverse_f64(P(h), C, 0.01, 0, P(hinv), P(hws), P(info), st)
    work = f32(R, C)
    gws = torch.empty(lib.aeqb_gptq_workspace_bytes(R, C), dtype=torch.uint8, device=dev)
    sc = (ws_in[0].abs().amax(dim=1) / 7.0).contiguous()
    cases.append(("xtx 8192 tokens (gptq hessian)", lambda i: lib.aeqb_xtx_f32(
        P(x), 8192, C, 0.25, P(h), P(xws) if xws_n else None, st), 0))
    cases.append(("hessian_inverse (gptq)", lambda i: lib.aeqb_hessian_inverse_f64(
        P(h), C, 0.01, 0, P(hinv), P(hws), P(info), st), 0))

    def gptq_call(i):
      work.copy_(ws_in[0])
      return lib.aeqb_gptq_quantize_f32(P(work), R, C, P(hinv), P(sc), None, 1, 0, 4, 1, 64, P(q), P(gws), st)
    cases.append(("gptq_quantize OBS loop (+copy)", gptq_call, 0))

  for name, fn, nbytes in cases:
    if a.only and a.only not in name:
      continue
    iters = a.iters if nbytes else 3
    for i in range(3):
      rc = fn(i)
      if rc:
        print(name, "FAILED:", lib.aeqb_last_error().decode())
        break
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
      fn(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    gbs = nbytes / us / 1e3 if nbytes else 0.0
    print(f"{name:40s} {us:10.1f} us  {gbs:8.1f} GB/s  frac {gbs / peak:5.3f}")


if __name__ == "__main__":
  main()
