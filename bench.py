"""Contract benchmark: fp32 weight GB/s requantised (INT8 per-channel, INT4 block-32).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): per-channel symmetric INT8 weight-only over a
synthetic stack of T [4096, 4096] fp32 FC weights per GPU (default T = 64 = 4 GiB,
far larger than the 126 MB L2, so no flush is needed between steps).  One step =
one pass of the hot path over the whole stack.  The same stack is also pushed
through the INT4 block-32 packed path and reported under "modes".

  value      device-resident: inputs already in HBM, one batched C-ABI call per step
  e2e        the same stack through the host-buffer C-ABI call (page-locked host
             arrays in and out; H2D + kernel + D2H inside the timed region)
  roofline   algorithmic bytes of the dominant kernel / its CUDA-event duration
             vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline   oracle (NumPy port of the reference path) on a bounded sample

--impl reference times the reference's CPU path (the oracle port; the Python
reference itself cannot travel to the GPU box) on the host cores.
Weights are sharded by tensor across ranks (weak scaling: T per GPU); the only
collective is one NCCL all-gather of the per-channel scales per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")):
  if _p not in sys.path:
    sys.path.insert(0, _p)

import numpy as np  # noqa: E402

ROWS, COLS = 4096, 4096
METRIC = "weight GB/s requantized (INT8-perch, INT4-blk32)"
UNIT = "GB/s"


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--tensors", type=int, default=64, help="[4096,4096] tensors per GPU")
  ap.add_argument("--e2e-tensors", type=int, default=16, help="stack size of the host-buffer arm")
  ap.add_argument("--cpu-sample", type=int, default=12, help="tensors timed by the CPU baseline")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  return ap.parse_args()


def workload_name(t):
  return (f"configs[1]: per-channel symmetric INT8 weight-only over synthetic {ROWS}x{COLS} FC stack,"
          f" {t} tensors = {t * ROWS * COLS * 4 / 2**30:.0f} GiB fp32 per GPU")


# ------------------------------------------------------------------ CPU arms
def cpu_pass(weights, mode):
  from oracle import aeq_oracle as O
  for w in weights:
    if mode == "int8":
      O.minmax_requant(w, 8, True)
    else:
      r = O.minmax_requant(w, 4, True, block=32)
      O.pack_bits(4, r["q"])
      O.blockwise_scale_fp16(r["scale"])


def cpu_weights(n):
  from oracle import aeq_oracle as O
  return [O.synthetic_weight(ROWS, COLS, i) for i in range(n)]


def cpu_threads():
  try:
    from threadpoolctl import threadpool_info
    n = [p.get("num_threads", 1) for p in threadpool_info()]
    return max(n) if n else 1
  except Exception:
    return 1


def cpu_baseline(n_sample):
  """Oracle port of the reference path, bounded sample, single NumPy thread (as shipped)."""
  ws = cpu_weights(n_sample)
  out = {}
  for mode in ("int8", "int4"):
    cpu_pass(ws[:1], mode)  # warm-up
    t0 = time.perf_counter()
    cpu_pass(ws, mode)
    dt = time.perf_counter() - t0
    out[mode] = n_sample * ROWS * COLS * 4 / dt / 1e9
  return out


def run_reference(a):
  """--impl reference: the reference's CPU implementation of the path (oracle port)."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  n = max(1, min(a.cpu_sample, 4))
  ws = cpu_weights(n)
  for _ in range(max(1, min(a.warmup, 1))):
    cpu_pass(ws[:1], "int8")
  steps = max(1, min(a.steps, 5))
  t0 = time.perf_counter()
  for _ in range(steps):
    cpu_pass(ws, "int8")
  dt = (time.perf_counter() - t0) / steps
  v = n * ROWS * COLS * 4 / dt / 1e9
  t1 = time.perf_counter()
  cpu_pass(ws, "int4")
  v4 = n * ROWS * COLS * 4 / (time.perf_counter() - t1) / 1e9
  sample = f"{n} of the workload's [{ROWS},{COLS}] tensors per step, {steps} steps"
  print(json.dumps({
      "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
      "steps": steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": {"workload": workload_name(a.tensors), "sample": sample},
      "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                       "blas_threads_available": cpu_threads(), "host_cores": os.cpu_count()},
      "modes": {"int8_perchannel": {"value": v}, "int4_block32_packed": {"value": v4}},
      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
      "note": "reference arm = NumPy port of the reference CPU path (oracle/aeq_oracle.py, pinned"
              " bit-exact to the reference); the reference is single-threaded NumPy on this path",
  }), flush=True)


# ------------------------------------------------------------------ clocks
class ClockSampler:
  Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
       "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
       "clocks_event_reasons.sw_power_cap")

  def __init__(self, index):
    self.index = index
    self.rows = []
    self.proc = None

  def start(self):
    try:
      self.proc = subprocess.Popen(
          ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
           "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      threading.Thread(target=self._read, daemon=True).start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(",")])

  def stop(self):
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    time.sleep(0.15)
    self.proc.terminate()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in self.rows:
      try:
        sm.append(float(r[1]))
        mx.append(float(r[2]))
        for name, v in zip(names, r[4:8]):
          if v.lower().startswith("active"):
            reasons.add(name)
      except Exception:
        continue
    return {"sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
            "samples": len(sm)}


# ------------------------------------------------------------------ GPU arm
def main():
  a = parse()
  if a.impl == "reference":
    run_reference(a)
    return

  import torch
  import torch.distributed as dist
  from aeq_b200 import _lib, device, host

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a CUDA device: aeq_b200 has no CPU fallback")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  lib = _lib.load()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  T = a.tensors
  gen = torch.Generator(device=dev).manual_seed(1000 + rank)
  ws = []
  for _ in range(T):  # N(0, 0.02) with one x20 outlier per 1024 elements
    w = torch.randn(ROWS, COLS, device=dev, generator=gen) * 0.02
    w.view(-1)[::1024] *= 20.0
    ws.append(w)
  n_bytes = T * ROWS * COLS * 4

  state = {}
  gathered = (torch.empty(world * T * ROWS, dtype=torch.float32, device=dev) if world > 1 else None)

  def step_int8():
    state["r8"] = device.requant_rows_batch(ws, 8, True, outs=state.get("r8"))
    if world > 1:  # the path's one collective: all ranks learn every per-channel scale
      flat = torch.cat([o.scale.view(-1) for o in state["r8"]])
      dist.all_gather_into_tensor(gathered, flat)

  def step_int4():
    state["b4"] = device.requant_blocks_batch(ws, 32, 4, outs=state.get("b4"))

  def timed(fn, steps, warmup):
    for _ in range(warmup):
      fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.aeqb_launch_count()
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = lib.aeqb_launch_count() - l0
    if world > 1:
      t = torch.tensor([ms], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = float(t.item())
    return ms, launches

  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  ms8, launches8 = timed(step_int8, a.steps, max(a.warmup, 3))
  ms4, launches4 = timed(step_int4, a.steps, max(a.warmup, 3))
  clocks = sampler.stop() if rank == 0 else None
  value8 = world * n_bytes / ms8 / 1e6
  value4 = world * n_bytes / ms4 / 1e6

  # ---- roofline of the dominant kernel (one persistent launch per step at T <= 64)
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
  except Exception:
    pass
  peak = float(peaks.get("hbm_gbs", 6650.0))
  peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  l0 = lib.aeqb_launch_count()
  e0.record()
  for _ in range(a.steps):
    state["r8"] = device.requant_rows_batch(ws, 8, True, outs=state["r8"])
  e1.record()
  torch.cuda.synchronize()
  k_launches = lib.aeqb_launch_count() - l0
  k_ms = e0.elapsed_time(e1) / max(k_launches, 1)
  alg_bytes = (n_bytes / 4) * 5 + T * ROWS * 8  # 4 B read + 1 B written per weight, 8 B/row scale+zp
  alg_per_launch = alg_bytes * a.steps / max(k_launches, 1)
  achieved = alg_per_launch / k_ms / 1e6
  e0.record()
  for _ in range(a.steps):
    state["b4"] = device.requant_blocks_batch(ws, 32, 4, outs=state["b4"])
  e1.record()
  torch.cuda.synchronize()
  k4_ms = e0.elapsed_time(e1) / a.steps
  achieved4 = (n_bytes / 4) * 4.5625 / k4_ms / 1e6

  # ---- e2e: host-buffer C-ABI call, pinned host arrays in and out
  Te = max(1, min(a.e2e_tensors, T))
  h_in = [host.pinned_empty((ROWS, COLS), np.float32) for _ in range(Te)]
  for i, h in enumerate(h_in):
    torch.from_numpy(h).copy_(ws[i])
  torch.cuda.synchronize()
  e_outs = host.requant_rows(h_in, 8, True, alloc=host.pinned_empty)

  def e2e_step():
    host.requant_rows(h_in, 8, True, outs=e_outs)

  for _ in range(2):
    e2e_step()
  barrier()
  e2e_steps = max(3, min(a.steps, 10))
  t0 = time.perf_counter()
  for _ in range(e2e_steps):
    e2e_step()  # returns when every output is in host memory
  e2e_s = (time.perf_counter() - t0) / e2e_steps
  if world > 1:
    t = torch.tensor([e2e_s], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
  e2e_val = world * Te * ROWS * COLS * 4 / e2e_s / 1e9
  h2d = Te * ROWS * COLS * 4
  d2h = Te * (ROWS * COLS + ROWS * 8)
  # spot-check the e2e result against the device-resident path (same arithmetic)
  ok = bool((torch.from_numpy(e_outs[0][0]).to(dev) == state["r8"][0].q).all())

  cpu = None
  if rank == 0 and not a.no_cpu_baseline:
    c = cpu_baseline(a.cpu_sample)
    cpu = {"value": c["int8"], "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"{a.cpu_sample} of the workload's [{ROWS},{COLS}] tensors, one pass, oracle/aeq_oracle.py"
                     " (NumPy restatement pinned bit-exact to the reference; single-threaded like the reference)",
           "int4_block32_packed": c["int4"], "host_cores": os.cpu_count(),
           "blas_threads_available": cpu_threads()}

  if rank == 0:
    print(json.dumps({
        "metric": METRIC, "value": value8, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms8, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(T), "tensors_per_gpu": T, "shape": [ROWS, COLS],
                   "l2": "inputs are 4 GiB per GPU per step, >> 126 MB L2: no flush needed",
                   "collective": "one NCCL all-gather of per-channel scales per step" if world > 1 else "none (N=1)",
                   "sharding": "tensors partitioned across ranks, no data-path collective"},
        "modes": {
            "int8_perchannel": {"value": value8, "ms_per_step": ms8, "launches_per_step": launches8 / a.steps},
            "int4_block32_packed": {"value": value4, "ms_per_step": ms4, "launches_per_step": launches4 / a.steps,
                                    "roofline_frac": achieved4 / peak, "achieved_hbm_gbs": achieved4,
                                    "bytes_per_weight": 4.5625}},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None,
                     "kernel": "requant_rows_stream<32768,8>", "bytes_per_weight": 5.0,
                     "algorithmic_bytes_per_launch": alg_per_launch, "launch_ms": k_ms,
                     "peak_source": peak_src},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "tensors": Te, "ms_per_step": e2e_s * 1e3, "matches_device_path": ok,
                "api": "aeq_b200.host.requant_rows -> aeqb_host_requant_rows_batch_f32 (pinned host buffers)"},
        "cpu_baseline": cpu,
        "gpu_launches": int(launches8),
        "clocks": clocks,
    }), flush=True)
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
