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
  ap.add_argument("--cpu-sample", type=int, default=32, help="tensors timed by the CPU baseline")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--modes", default="headline", choices=["headline", "all"],
                  help="all: also time OCTAV, MSE, Hadamard, calibration min/max and GPTQ (extra keys under modes)")
  return ap.parse_args()


def workload_name(t):
  return (f"configs[1]: per-channel symmetric INT8 weight-only over synthetic {ROWS}x{COLS} FC stack,"
          f" {t} tensors = {t * ROWS * COLS * 4 / 2**30:.0f} GiB fp32 per GPU")


# ------------------------------------------------------------------ CPU arms
def cpu_one(w, mode):
  from oracle import aeq_oracle as O
  if mode == "int8":
    O.minmax_requant(w, 8, True)
  else:
    r = O.minmax_requant(w, 4, True, block=32)
    O.pack_bits(4, r["q"])
    O.blockwise_scale_fp16(r["scale"])


def cpu_pass(weights, mode, threads=1):
  """One pass of the reference's CPU path over `weights`.  The reference itself is a
  single-threaded NumPy loop over tensors (params_generator.py:110-183); threads > 1 gives it
  every host core by running independent tensors concurrently (NumPy releases the GIL)."""
  if threads <= 1:
    for w in weights:
      cpu_one(w, mode)
    return
  from concurrent.futures import ThreadPoolExecutor
  with ThreadPoolExecutor(max_workers=threads) as ex:
    list(ex.map(lambda w: cpu_one(w, mode), weights))


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


def host_threads():
  try:
    return max(1, len(os.sched_getaffinity(0)))
  except Exception:
    return max(1, os.cpu_count() or 1)


def cpu_baseline(n_sample):
  """Oracle port of the reference path on a bounded sample: all host cores (tensors in
  parallel) and, for the record, one core (how the reference ships)."""
  threads = host_threads()
  ws = cpu_weights(n_sample)
  out = {"threads": threads}
  for mode in ("int8", "int4"):
    cpu_pass(ws[:threads], mode, threads)  # warm-up
    t0 = time.perf_counter()
    cpu_pass(ws, mode, threads)
    dt = time.perf_counter() - t0
    out[mode] = n_sample * ROWS * COLS * 4 / dt / 1e9
  t0 = time.perf_counter()
  cpu_pass(ws[:2], "int8", 1)
  out["int8_single_thread"] = 2 * ROWS * COLS * 4 / (time.perf_counter() - t0) / 1e9
  return out


def run_reference(a):
  """--impl reference: the reference's CPU implementation of the path (oracle port)."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  threads = host_threads()
  n = max(threads, min(a.cpu_sample, 32))
  ws = cpu_weights(n)
  warm = max(1, min(a.warmup, 2))
  for _ in range(warm):
    cpu_pass(ws, "int8", threads)
  steps = max(1, min(a.steps, 5))
  t0 = time.perf_counter()
  for _ in range(steps):
    cpu_pass(ws, "int8", threads)
  dt = (time.perf_counter() - t0) / steps
  v = n * ROWS * COLS * 4 / dt / 1e9
  t1 = time.perf_counter()
  cpu_pass(ws, "int4", threads)
  v4 = n * ROWS * COLS * 4 / (time.perf_counter() - t1) / 1e9
  t1 = time.perf_counter()
  cpu_pass(ws[:2], "int8", 1)
  v1 = 2 * ROWS * COLS * 4 / (time.perf_counter() - t1) / 1e9
  sample = (f"{n} of the workload's [{ROWS},{COLS}] tensors per step, {steps} steps,"
            f" {threads} threads over independent tensors")
  emit(({
      "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
      "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": {"workload": workload_name(a.tensors), "sample": sample},
      "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                       "single_thread_value": v1, "host_cores": os.cpu_count()},
      "modes": {"int8_perchannel": {"value": v}, "int4_block32_packed": {"value": v4},
                "int8_perchannel_single_thread": {"value": v1}},
      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
      "note": "reference arm = NumPy port of the reference CPU path (oracle/aeq_oracle.py, pinned"
              " bit-exact to the reference; the Python reference cannot travel to the GPU box)."
              " The reference is single-threaded NumPy on this path; this arm additionally runs"
              " independent tensors on every host core",
  }))


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


# ------------------------------------------------------------------ the other rows of the path
def extra_modes(dev, ws, peak, steps):
  """Device-resident timings of the remaining hot-path rows on the same stack."""
  import torch
  from aeq_b200 import device
  out = {}
  T = min(len(ws), 16)
  n_bytes = T * ROWS * COLS * 4

  def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

  reps = max(3, min(steps, 10))

  def octav(bits, block):
    def run():
      for w in ws[:T]:
        if block:
          c = device.octav_clip_blocks(w, block, bits)
          device.requant_blocks(w, block, bits, clip=c, want_q=False, want_packed=True)
        else:
          c = device.octav_clip_rows(w, bits)
          device.requant_rows(w, bits, True, clip=c)
    return run

  for name, fn, bpw in (("octav_int4_perchannel", octav(4, 0), 9.0),
                        ("octav_int4_block32_packed", octav(4, 32), 8.5625 + 10 * 4 / 32)):
    ms = timeit(fn, reps)
    out[name] = {"value": n_bytes / ms / 1e6, "ms_per_step": ms, "tensors": T, "bytes_per_weight": bpw,
                 "roofline_frac": (n_bytes / 4) * bpw / ms / 1e6 / peak}

  def mse():
    for w in ws[:T]:
      device.requant_mse_rows(w, 8, 0.05408)
  ms = timeit(mse, reps)
  out["mse_int8_perchannel"] = {"value": n_bytes / ms / 1e6, "ms_per_step": ms, "tensors": T,
                                "bytes_per_weight": 5.0, "roofline_frac": (n_bytes / 4) * 5.0 / ms / 1e6 / peak,
                                "note": "fused: sum of squares, scale and integers in one pass (one launch per tensor)"}

  rot = torch.empty_like(ws[0])

  def hadamard():
    for w in ws[:T]:
      device.hadamard_rows(w, COLS, out=rot)
      c = device.octav_clip_rows(rot, 4)
      device.requant_rows(rot, 4, True, clip=c)
  ms = timeit(hadamard, reps)
  out["hadamard4096_octav_int4"] = {"value": n_bytes / ms / 1e6, "ms_per_step": ms, "tensors": T,
                                    "bytes_per_weight": 17.0,
                                    "roofline_frac": (n_bytes / 4) * 17.0 / ms / 1e6 / peak}

  def had_only():
    for w in ws[:T]:
      device.hadamard_rows(w, COLS, out=rot)
  ms = timeit(had_only, reps)
  out["hadamard4096_rotate_only"] = {"value": n_bytes / ms / 1e6, "ms_per_step": ms, "tensors": T,
                                     "bytes_per_weight": 8.0,
                                     "roofline_frac": (n_bytes / 4) * 8.0 / ms / 1e6 / peak}

  # calibration: per-batch min/max with the (-3e38, 3e38) filter over [8, 512, 4096] activations
  acts = [w.view(8, 512, 4096) for w in ws[:T]]

  def calib():
    for a in acts:
      device.minmax_tensor(a, -3e38, 3e38)
  ms = timeit(calib, reps)
  out["calibration_minmax"] = {"value": n_bytes / ms / 1e6, "unit": "activation GB/s", "ms_per_step": ms,
                               "batches": T, "bytes_per_element": 4.0,
                               "roofline_frac": n_bytes / ms / 1e6 / peak}

  # the same batches, eight per launch (aeqb_minmax_tensors_f32): min / max of a batch do not depend
  # on other batches, only the EMA that consumes them is sequential (and O(1) per batch on the host)
  def calib8():
    for i in range(0, len(acts), 8):
      device.minmax_tensors(acts[i:i + 8], -3e38, 3e38)
  ms = timeit(calib8, reps)
  out["calibration_minmax_8_per_launch"] = {
      "value": n_bytes / ms / 1e6, "unit": "activation GB/s", "ms_per_step": ms, "batches": T,
      "bytes_per_element": 4.0, "roofline_frac": n_bytes / ms / 1e6 / peak}

  # GPTQ on one [4096, 4096] layer: Hessian from 8192 tokens, damped inverse, OBS loop
  x = torch.randn(8, 1024, COLS, device=dev)
  ms_h = timeit(lambda: device.xtx(x, 2.0 / 8), 3)
  h = device.xtx(x, 2.0 / 8)
  ms_inv = timeit(lambda: device.hessian_inverse(h, 0.01), 2)
  hinv = device.hessian_inverse(h, 0.01)
  w = ws[0]
  sc = w.abs().amax(dim=1) / 7.0
  ms_q = timeit(lambda: device.gptq_quantize(w, hinv, sc, None, 0, 4, True), 2)
  out["gptq_int4_4096x4096"] = {
      "value": ROWS * COLS * 4 / (ms_inv + ms_q) / 1e6, "ms_hessian_8192_tokens": ms_h,
      "hessian_tflops": 2.0 * 8192 * COLS * COLS / ms_h / 1e9,
      "ms_hessian_inverse": ms_inv, "ms_obs_loop": ms_q,
      "obs_loop_tflops": 1.0 * ROWS * COLS * COLS / ms_q / 1e9,
      "note": "value = fp32 weight bytes / (inverse + OBS loop) for one layer, Hessian given"}
  del x, h, hinv
  out.update(config_sets(dev, peak, reps))
  return out


def headline_8b(dev, peak):
  """BASELINE.json's target set at full size on ONE GPU: 477 x [4096, 4096] fp32 = 8.00 B parameters =
  32.0 GB resident in HBM, requantised per channel to INT8 and in blocks of 32 to
  packed INT4; eight persistent launches of <= 64 tensors per pass."""
  import torch
  from aeq_b200 import device
  n = 477
  free, _ = torch.cuda.mem_get_info(dev)
  if free < 60e9:
    return {"headline_8b_params": {"skipped": f"only {free / 1e9:.0f} GB of HBM free"}}
  g = torch.Generator(device=dev).manual_seed(8)
  ws = []
  for _ in range(n):
    w = torch.randn(ROWS, COLS, device=dev, generator=g) * 0.02
    w.view(-1)[::1024] *= 20.0
    ws.append(w)
  n_bytes = n * ROWS * COLS * 4
  out = {}
  for name, bpw, fn in (
      ("int8_perchannel", 5.0, lambda st: device.requant_rows_batch(ws, 8, True, outs=st.get("o"))),
      ("int4_block32_packed", 4.5625, lambda st: device.requant_blocks_batch(ws, 32, 4, outs=st.get("o")))):
    st = {}
    for _ in range(2):
      st["o"] = fn(st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
      st["o"] = fn(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[name] = {"value": n_bytes / ms / 1e6, "ms_per_pass": ms,
                 "roofline_frac": (n_bytes / 4) * bpw / ms / 1e6 / peak}
    del st
  return {"headline_8b_params": {"tensors": n, "fp32_bytes": n_bytes, "unit": "GB/s of fp32 weight bytes", **out}}


def config_sets(dev, peak, reps):
  """BASELINE.json configs[2] and the weight side of configs[4] at their own tensor shapes
  (SURVEY.md §8d): the Gemma-2B FC set through INT4 block-32, the Llama-7B FC set through INT8 /
  INT4 per-channel (11008-wide rows take the rows kernel's 64 KiB stage class)."""
  import torch
  from aeq_b200 import device
  out = {}

  def make(shapes, layers):
    g = torch.Generator(device=dev).manual_seed(4242)
    ws = []
    for _ in range(layers):
      for r, c in shapes:
        w = torch.randn(r, c, device=dev, generator=g) * 0.02
        w.view(-1)[::1024] *= 20.0
        ws.append(w)
    return ws

  def timeit(fn):
    for _ in range(2):
      fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

  gemma = [(2048, 2048), (2048, 2048), (256, 2048), (256, 2048), (16384, 2048), (16384, 2048), (2048, 16384)]
  ws = make(gemma, 18)
  n_bytes = sum(w.numel() for w in ws) * 4
  st = {}

  def g4():
    st["o"] = device.requant_blocks_batch(ws, 32, 4, outs=st.get("o"))
  ms = timeit(g4)
  out["cfg3_gemma2b_int4_block32_packed"] = {
      "value": n_bytes / ms / 1e6, "ms_per_step": ms, "tensors": len(ws), "fp32_bytes": n_bytes,
      "bytes_per_weight": 4.5625, "roofline_frac": (n_bytes / 4) * 4.5625 / ms / 1e6 / peak}
  del ws, st

  llama = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
  ws = make(llama, 4)
  n_bytes = sum(w.numel() for w in ws) * 4
  for bits, bpw in ((8, 5.0), (4, 4.5)):
    st = {}

    def rows():
      st["o"] = device.requant_rows_batch(ws, bits, True, want_q=(bits == 8), want_packed=(bits == 4),
                                           outs=st.get("o"))
    ms = timeit(rows)
    out[f"cfg5_llama7b_int{bits}_perchannel"] = {
        "value": n_bytes / ms / 1e6, "ms_per_step": ms, "tensors": len(ws), "fp32_bytes": n_bytes,
        "bytes_per_weight": bpw, "roofline_frac": (n_bytes / 4) * bpw / ms / 1e6 / peak}
    del st
  del ws

  # ---- configs[4], one Llama-7B decoder layer: Hadamard-rotated INT4 and GPTQ INT4 over its seven
  # FC weights.  GPTQ needs four Hessians (q/k/v share their input, gate/up too): X^T X over
  # `tok` tokens each, the damped inverse, then the OBS loop per weight.
  def gcd_pow2(n):
    return n & -n

  layer = make(llama, 1)
  rot = [torch.empty_like(w) for w in layer]

  def hadamard_layer():
    for w, r in zip(layer, rot):
      device.hadamard_rows(w, min(gcd_pow2(w.shape[1]), 4096), out=r)
      c = device.octav_clip_rows(r, 4)
      device.requant_rows(r, 4, True, clip=c)
  ms = timeit(hadamard_layer)
  lbytes = sum(w.numel() for w in layer) * 4
  out["cfg5_llama7b_layer_hadamard_octav_int4"] = {
      "value": lbytes / ms / 1e6, "ms_per_layer": ms, "tensors": len(layer), "fp32_bytes": lbytes,
      "bytes_per_weight": 17.0, "roofline_frac": (lbytes / 4) * 17.0 / ms / 1e6 / peak}
  del rot

  tok = 16384
  t0 = torch.cuda.Event(enable_timing=True)
  t1 = torch.cuda.Event(enable_timing=True)
  parts = {"hessian": 0.0, "inverse": 0.0, "obs_loop": 0.0}
  g = torch.Generator(device=dev).manual_seed(99)

  def timed(fn):
    t0.record()
    r = fn()
    t1.record()
    torch.cuda.synchronize()
    return r, t0.elapsed_time(t1)

  for rep in range(2):  # second pass is the measurement
    for key in parts:
      parts[key] = 0.0
    hinv = {}
    for name, k in (("attn_in", 4096), ("attn_out", 4096), ("mlp_in", 4096), ("mlp_mid", 11008)):
      x = torch.randn(tok, k, device=dev, generator=g)
      h, ms_h = timed(lambda: device.xtx(x, 2.0 / 8))
      del x
      hi, ms_i = timed(lambda: device.hessian_inverse(h, 0.01))
      del h
      hinv[name] = hi
      parts["hessian"] += ms_h
      parts["inverse"] += ms_i
    feeds = ["attn_in"] * 3 + ["attn_out"] + ["mlp_in"] * 2 + ["mlp_mid"]
    for w, feed in zip(layer, feeds):
      sc = w.abs().amax(dim=1) / 7.0
      _, ms_q = timed(lambda: device.gptq_quantize(w, hinv[feed], sc, None, 0, 4, True))
      parts["obs_loop"] += ms_q
    del hinv
  total = sum(parts.values())
  del layer
  out.update(headline_8b(dev, peak))
  out["cfg5_llama7b_layer_gptq_int4"] = {
      "value": lbytes / total / 1e6, "ms_per_layer": total, "ms_hessian_4x": parts["hessian"],
      "ms_inverse_4x": parts["inverse"], "ms_obs_loop_7x": parts["obs_loop"], "tokens_per_hessian": tok,
      "note": "value = fp32 weight bytes of the layer / (4 Hessians over 16384 tokens + 4 inverses + 7 OBS loops);"
              " configs[4] uses 262144 tokens per Hessian: scale ms_hessian_4x by 16"}
  return out


# ------------------------------------------------------------------ GPU arm
_REAL_STDOUT = None


def quiet_stdout():
  """stdout must carry exactly ONE JSON line.  Libraries print there too (NCCL writes
  "NCCL version ..." with a bare printf when NCCL_DEBUG=VERSION), so file descriptor 1 is pointed
  at stderr for the whole run and the result is written to a duplicate of the original stdout."""
  global _REAL_STDOUT
  sys.stdout.flush()
  _REAL_STDOUT = os.dup(1)
  os.dup2(2, 1)


def emit(obj) -> None:
  line = (json.dumps(obj) + "\n").encode()
  if _REAL_STDOUT is None:
    sys.stdout.write(line.decode())
    sys.stdout.flush()
  else:
    os.write(_REAL_STDOUT, line)


def main():
  a = parse()
  quiet_stdout()
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
    # stdout carries exactly one JSON line: NCCL's own log lines ("NCCL version ...", INFO output
    # when the caller asks for it) go to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
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
  # Per-channel scales of this rank's tensors live in ONE flat buffer (each tensor's scale output
  # is a view into it), so the path's single exchange needs no packing step.  At N > 1 that
  # buffer is this rank's row of a [world, T * ROWS] buffer every rank maps (aeq_b200/peer.py):
  # the requantisation kernel stores each row's scale into every peer's copy from its own
  # epilogue, so the all-gather of scales costs no launch.  If peer mapping is not available on
  # the box (or AEQB_BENCH_NCCL_GATHER is set) the exchange is one NCCL all-gather per step.
  mirror, mirror_note = None, None
  if world > 1 and not os.environ.get("AEQB_BENCH_NCCL_GATHER"):
    try:
      from aeq_b200 import peer
      mirror = peer.PeerScales(T * ROWS, dev)
    except Exception as e:  # pylint: disable=broad-except
      mirror, mirror_note = None, f"{type(e).__name__}: {e}"[:200]
  flat_scales = mirror.local if mirror is not None else torch.empty(T * ROWS, dtype=torch.float32, device=dev)
  gathered = (torch.empty(world * T * ROWS, dtype=torch.float32, device=dev) if world > 1 else None)
  state["r8"] = [device.Requantized(
      torch.empty((ROWS, COLS), dtype=torch.int8, device=dev), None,
      flat_scales[i * ROWS:(i + 1) * ROWS].view(ROWS, 1),
      torch.empty((ROWS, 1), dtype=torch.int32, device=dev)) for i in range(T)]

  def step_int8():
    if mirror is not None:  # scales reach every rank's gathered buffer from inside the kernel
      device.requant_rows_batch(ws, 8, True, outs=state["r8"], mirror=mirror)
      return
    device.requant_rows_batch(ws, 8, True, outs=state["r8"])
    if world > 1:  # the path's one collective: all ranks learn every per-channel scale
      dist.all_gather_into_tensor(gathered, flat_scales)

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
  exchange_ok = None
  if mirror is not None:  # the in-kernel exchange against NCCL's all-gather of the same scales
    mirror.sync()
    dist.all_gather_into_tensor(gathered, flat_scales)
    okt = torch.tensor([int(torch.equal(gathered.view(world, -1), mirror.gathered))], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    exchange_ok = bool(okt.item())
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
    device.requant_rows_batch(ws, 8, True, outs=state["r8"])
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

  # dram bytes per launch from the committed `ncu --set full` capture of this kernel, scaled by
  # tensor count when the capture used a different stack size (profiles/traffic.json)
  traffic = None
  try:
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["requant_rows_stream"]
    traffic = tj["dram_bytes_per_launch"] * (alg_per_launch / tj["algorithmic_bytes_per_launch"])
  except Exception:
    pass

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
  # what bounds the e2e arm: this box's host->device copy rate from pinned memory, measured with
  # a device->host copy of a quarter of the bytes running beside it (the arm's own 4 : 1 mix)
  pin_in = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
  pin_out = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
  d_in, d_out = torch.empty_like(pin_in, device=dev), torch.empty(64 << 20, dtype=torch.uint8, device=dev)
  s2 = torch.cuda.Stream()
  best = 0.0
  for _ in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_in.copy_(pin_in, non_blocking=True)
    with torch.cuda.stream(s2):
      pin_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    best = max(best, pin_in.numel() / (time.perf_counter() - t0) / 1e9)
  pcie_h2d = best
  del pin_in, pin_out, d_in, d_out

  extra = {}
  if a.modes == "all" and rank == 0 and world == 1:
    extra = extra_modes(dev, ws, peak, a.steps)

  cpu = None
  if rank == 0 and not a.no_cpu_baseline:
    c = cpu_baseline(a.cpu_sample)
    cpu = {"value": c["int8"], "unit": UNIT, "cores": c["threads"], "kind": "port",
           "sample": f"{a.cpu_sample} of the workload's [{ROWS},{COLS}] tensors, one pass, oracle/aeq_oracle.py"
                     f" (NumPy restatement pinned bit-exact to the reference), {c['threads']} threads over"
                     " independent tensors",
           "int4_block32_packed": c["int4"], "single_thread_value": c["int8_single_thread"],
           "host_cores": os.cpu_count()}

  if rank == 0:
    emit(({
        "metric": METRIC, "value": value8, "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms8, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(T), "tensors_per_gpu": T, "shape": [ROWS, COLS],
                   "l2": "inputs are 4 GiB per GPU per step, >> 126 MB L2: no flush needed",
                   "collective": ("none (N=1)" if world == 1 else
                                  "per-channel scales stored into every peer's gathered buffer by the requantisation kernel itself (NVLink peer memory, aeqb_requant_rows_batch_mirror_f32); no collective launch" if mirror is not None else
                                  "one NCCL all-gather of per-channel scales per step"),
                   "scale_exchange_matches_nccl_all_gather": exchange_ok,
                   "peer_mapping_error": mirror_note,
                   "sharding": "tensors partitioned across ranks, no data-path collective"},
        "modes": {
            "int8_perchannel": {"value": value8, "ms_per_step": ms8, "launches_per_step": launches8 / a.steps},
            "int4_block32_packed": {"value": value4, "ms_per_step": ms4, "launches_per_step": launches4 / a.steps,
                                    "roofline_frac": achieved4 / peak, "achieved_hbm_gbs": achieved4,
                                    "bytes_per_weight": 4.5625}, **extra},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "kernel": "requant_rows_stream<16384,4,8,3,false>", "bytes_per_weight": 5.0,
                     "algorithmic_bytes_per_launch": alg_per_launch, "launch_ms": k_ms,
                     "peak_source": peak_src},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "tensors": Te, "ms_per_step": e2e_s * 1e3, "matches_device_path": ok,
                "bound": "pcie", "h2d_gbs_measured": pcie_h2d,
                "frac_of_h2d": (e2e_val / world) / pcie_h2d if pcie_h2d > 0 else None,
                "api": "aeq_b200.host.requant_rows -> aeqb_host_requant_rows_batch_f32 (pinned host buffers)"},
        "cpu_baseline": cpu,
        "gpu_launches": int(launches8),
        "clocks": clocks,
    }))
  if mirror is not None:
    mirror.close()
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
