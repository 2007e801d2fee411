"""Contract benchmark: fp32 weight GB/s requantised (INT8 per-channel, INT4 block-32).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): per-channel symmetric INT8 weight-only over a
synthetic stack of [4096, 4096] fp32 FC weights, T = 477 tensors per GPU = 8.00 B parameters =
32 GB (the set BASELINE's metric is quoted on; far larger than the 126 MB L2, so no flush between
steps).  One step = one pass of the hot path over the whole stack.  Weak scaling: T per GPU.

  value          device-resident: inputs already in HBM, batched C-ABI calls, CUDA events
  e2e            the same arithmetic through the host-buffer C-ABI call (page-locked host arrays in
                 and out; H2D + kernel + D2H inside the timed region); `e2e.pageable` = the boundary
                 the reference presents (read-only np.memmap views in, fresh NumPy arrays out);
                 `e2e.quantizer` = Quantizer(model.tflite).quantize() on a 4 GiB synthetic FC model
  roofline       algorithmic bytes of the dominant kernel / its CUDA-event duration vs
                 MEASURED_PEAKS.json hbm_gbs
  cpu_baseline   the reference's own CPU path on a bounded sample (oracle/_ref = the unmodified
                 reference sources copied by oracle/make_ref.py; the NumPy port if they are absent)
  modes          the other BASELINE.json configs at the same N, each with its exchange:
                 gemma2b_int4b32 (fp16 block scales stored into every peer from the kernel),
                 calib512 (all-gather of per-batch (min, max) + replicated EMA),
                 llama7b_gptq (Hadamard-rotated INT4 GPTQ, one decoder layer per rank, partial X^T X
                 reduced to the layer's owner)

--impl reference times the reference's CPU implementation of the path on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
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
GEMMA_LAYER = [(2048, 2048), (2048, 2048), (256, 2048), (256, 2048), (16384, 2048), (16384, 2048), (2048, 16384)]
LLAMA_LAYER = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
LLAMA_FEEDS = ["attn_in"] * 3 + ["attn_out"] + ["mlp_in"] * 2 + ["mlp_mid"]
LLAMA_INPUTS = [("attn_in", 4096), ("attn_out", 4096), ("mlp_in", 4096), ("mlp_mid", 11008)]
WORKLOADS = ("fc4096_int8", "gemma2b_int4b32", "calib512", "llama7b_gptq")


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=25)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default="fc4096_int8", choices=WORKLOADS,
                  help="which BASELINE.json config is the line's `value`; the others go under `modes`")
  ap.add_argument("--tensors", type=int, default=477, help="[4096,4096] tensors per GPU (477 = 8.00 B params)")
  ap.add_argument("--e2e-tensors", type=int, default=16, help="stack size of the host-buffer arms")
  ap.add_argument("--cpu-sample", type=int, default=32, help="tensors timed by the CPU baseline")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--gptq-tokens", type=int, default=262144,
                  help="tokens per Hessian (configs[4]: 128 sequences x 2048 tokens)")
  ap.add_argument("--modes", default="configs", choices=["headline", "configs", "all"],
                  help="headline: configs[1] only; configs (default): + configs[2..4] with their exchanges;"
                       " all: + OCTAV / MSE / Hadamard / single-kernel rows of the path")
  return ap.parse_args()


def workload_name(t):
  return (f"configs[1]: per-channel symmetric INT8 weight-only over synthetic {ROWS}x{COLS} FC stack,"
          f" {t} tensors = {t * ROWS * COLS / 1e9:.2f} B params = {t * ROWS * COLS * 4 / 1e9:.1f} GB fp32 per GPU")


def config_of(a):
  """Identical in both arms (the driver compares them)."""
  return {"workload": workload_name(a.tensors), "tensors_per_gpu": a.tensors, "shape": [ROWS, COLS],
          "l2": "inputs are 32 GB per GPU per step, >> 126 MB L2: no flush needed",
          "sharding": "tensors partitioned across ranks (weak scaling: fixed tensors per GPU), no data-path"
                      " collective; per-channel scales exchanged through NVLink peer stores"}


# ------------------------------------------------------------------ CPU arms
def _reference_modules():
  """The UNMODIFIED reference (oracle/_ref or /root/reference) or None."""
  try:
    from oracle import refshim
    if not refshim.available():
      return None
    import types
    ns = types.SimpleNamespace(kind=refshim.kind())
    ns.q = refshim.ref("qtyping")
    ns.nmm = refshim.ref("algorithms.uniform_quantize.naive_min_max_quantize")
    ns.tu = refshim.ref("transformations.transformation_utils")
    ns.op_info = refshim.fc_op_info
    return ns
  except Exception:  # pylint: disable=broad-except
    return None


def make_cpu_one():
  """(callable(w, mode), kind): one tensor through the reference's CPU path."""
  ref = _reference_modules()
  if ref is not None:
    G = ref.q.QuantGranularity
    cfg8 = ref.q.TensorQuantizationConfig(num_bits=8, symmetric=True, granularity=G.CHANNELWISE)
    cfg4 = ref.q.TensorQuantizationConfig(num_bits=4, symmetric=True, granularity=G.BLOCKWISE_32)
    info8, info4 = ref.op_info(cfg8), ref.op_info(cfg4)

    def one(w, mode):
      # naive_min_max_quantize.get_tensor_quant_params (:34-110) [+ pack_data (:293-353) and the
      # fp16 scale tensor of quantize_tensor._perform_blockwise_quantization (:129-133)]
      if mode == "int8":
        ref.nmm.get_tensor_quant_params(info8, cfg8, w, None)
      else:
        import ml_dtypes
        p = ref.nmm.get_tensor_quant_params(info4, cfg4, w, None)
        ref.tu.pack_data(4, np.ravel(p.quantized_data).view(np.uint8))
        p.scale.astype(ml_dtypes.bfloat16).astype(np.float16)
    return one, ref.kind

  from oracle import aeq_oracle as O

  def one_port(w, mode):
    if mode == "int8":
      O.minmax_requant(w, 8, True)
    else:
      r = O.minmax_requant(w, 4, True, block=32)
      O.pack_bits(4, r["q"])
      O.blockwise_scale_fp16(r["scale"])
  return one_port, "port"


def cpu_pass(one, weights, mode, threads=1):
  """One pass of the reference's CPU path over `weights`.  The reference itself is a
  single-threaded NumPy loop over tensors (params_generator.py:110-183); threads > 1 gives it
  every host core by running independent tensors concurrently (NumPy releases the GIL)."""
  if threads <= 1:
    for w in weights:
      one(w, mode)
    return
  from concurrent.futures import ThreadPoolExecutor
  with ThreadPoolExecutor(max_workers=threads) as ex:
    list(ex.map(lambda w: one(w, mode), weights))


def cpu_weights(n):
  from oracle import aeq_oracle as O
  return [O.synthetic_weight(ROWS, COLS, i) for i in range(n)]


def host_threads():
  try:
    return max(1, len(os.sched_getaffinity(0)))
  except Exception:  # pylint: disable=broad-except
    return max(1, os.cpu_count() or 1)


def cpu_baseline(n_sample):
  """The reference's CPU path on a bounded sample: all host cores (tensors in parallel) and, for
  the record, one core (how the reference ships)."""
  one, kind = make_cpu_one()
  threads = host_threads()
  ws = cpu_weights(n_sample)
  out = {"threads": threads, "kind": kind}
  for mode in ("int8", "int4"):
    cpu_pass(one, ws[:threads], mode, threads)  # warm-up
    t0 = time.perf_counter()
    cpu_pass(one, ws, mode, threads)
    out[mode] = n_sample * ROWS * COLS * 4 / (time.perf_counter() - t0) / 1e9
  t0 = time.perf_counter()
  cpu_pass(one, ws[:2], "int8", 1)
  out["int8_single_thread"] = 2 * ROWS * COLS * 4 / (time.perf_counter() - t0) / 1e9
  return out


def kind_text(kind):
  return {"_ref": "the UNMODIFIED reference (oracle/_ref: its own sources copied byte for byte by oracle/make_ref.py)",
          "reference": "the UNMODIFIED reference (/root/reference)",
          "port": "oracle/aeq_oracle.py (NumPy restatement pinned bit-exact to the reference; oracle/_ref absent)"}[kind]


def spec_kind(kind):
  """The contract's two values: "reference" = the reference's own code, "port" = the oracle."""
  return "port" if kind == "port" else "reference"


def kind_source(kind):
  return {"_ref": "oracle/_ref (travelling byte-for-byte copy made by oracle/make_ref.py)",
          "reference": "/root/reference", "port": "oracle/aeq_oracle.py"}[kind]


def run_reference(a):
  """--impl reference: the reference's own CPU implementation of the path on the host cores."""
  if int(os.environ.get("RANK", "0")) != 0:
    return
  one, kind = make_cpu_one()
  threads = host_threads()
  n = max(1, min(a.cpu_sample, a.tensors))
  ws = cpu_weights(n)
  for _ in range(max(a.warmup, 1)):
    cpu_pass(one, ws[:max(threads, 1)], "int8", threads)
  t0 = time.perf_counter()
  for _ in range(a.steps):
    cpu_pass(one, ws, "int8", threads)
  dt = (time.perf_counter() - t0) / a.steps
  v = n * ROWS * COLS * 4 / dt / 1e9
  t1 = time.perf_counter()
  cpu_pass(one, ws, "int4", threads)
  v4 = n * ROWS * COLS * 4 / (time.perf_counter() - t1) / 1e9
  t1 = time.perf_counter()
  cpu_pass(one, ws[:2], "int8", 1)
  v1 = 2 * ROWS * COLS * 4 / (time.perf_counter() - t1) / 1e9
  sample = (f"each step = {n} of the workload's {a.tensors} [{ROWS},{COLS}] tensors per GPU, {threads} threads over"
            f" independent tensors; {kind_text(kind)}")
  emit(({
      "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus,
      "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": config_of(a),
      "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": spec_kind(kind), "source": kind_source(kind),
                       "sample": sample,
                       "single_thread_value": v1, "host_cores": os.cpu_count()},
      "modes": {"int8_perchannel": {"value": v}, "int4_block32_packed": {"value": v4},
                "int8_perchannel_single_thread": {"value": v1}},
      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
      "note": "reference arm = naive_min_max_quantize.get_tensor_quant_params of the reference (single-threaded"
              " NumPy as shipped: single_thread_value) run on every host core over independent tensors",
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
    except Exception:  # pylint: disable=broad-except
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
      except Exception:  # pylint: disable=broad-except
        continue
    return {"sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
            "samples": len(sm)}


# ------------------------------------------------------------------ shared GPU helpers
class Ctx:
  """Per-process state of the GPU arm."""

  def __init__(self, a):
    import torch
    import torch.distributed as dist
    self.a = a
    self.torch, self.dist = torch, dist
    self.world = int(os.environ.get("WORLD_SIZE", "1"))
    self.rank = int(os.environ.get("RANK", "0"))
    self.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
      raise SystemExit("bench.py needs a CUDA device: aeq_b200 has no CPU fallback")
    torch.cuda.set_device(self.local)
    self.dev = torch.device("cuda", self.local)
    if self.world > 1:
      os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
      # a rank that waits five minutes in a collective is not going to be joined (a peer left its mode
      # on an error): fail instead of holding the launcher for ever
      import datetime
      dist.init_process_group("nccl", device_id=self.dev, timeout=datetime.timedelta(seconds=300))
    from aeq_b200 import _lib
    self.lib = _lib.load()
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # pylint: disable=broad-except
      pass
    self.peak = float(peaks.get("hbm_gbs", 6650.0))
    self.peak_src = ("MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks
                     else "fallback 6650 GB/s (B200_PROFILING.md)")

  def barrier(self):
    if self.world > 1:
      self.dist.barrier()
    self.torch.cuda.synchronize()

  def weights(self, shapes, seed):
    """N(0, 0.02) with one x20 outlier per 1024 elements, generated on the device."""
    torch = self.torch
    gen = torch.Generator(device=self.dev).manual_seed(seed)
    out = []
    # ONE allocation for the whole set, the tensors are views into it — the layout a model's weights have
    # in its flatbuffer (and one large mapping instead of hundreds of 64 MiB ones);
    # AEQB_BENCH_SEPARATE_ALLOCS=1 allocates every tensor on its own
    flat = None
    if not os.environ.get("AEQB_BENCH_SEPARATE_ALLOCS"):
      flat = torch.empty(sum(r * c for r, c in shapes), dtype=torch.float32, device=self.dev)
    off = 0
    for r, c in shapes:
      if flat is None:
        w = torch.randn(r, c, device=self.dev, generator=gen) * 0.02
      else:
        w = flat[off:off + r * c].view(r, c)
        off += r * c
        w.normal_(0.0, 0.02, generator=gen)
      w.view(-1)[::1024] *= 20.0
      out.append(w)
    return out

  def timed(self, fn, steps, warmup):
    """(max-over-ranks ms per step, launches, per-rank ms list): W untimed steps, then exactly K
    steps bracketed by barrier + synchronise, CUDA events on the launching stream."""
    torch = self.torch
    for _ in range(warmup):
      fn()
    self.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = self.lib.aeqb_launch_count()
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    self.barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = self.lib.aeqb_launch_count() - l0
    per_rank = [ms]
    if self.world > 1:
      t = torch.tensor([ms], device=self.dev)
      every = torch.empty(self.world, device=self.dev)
      self.dist.all_gather_into_tensor(every, t)
      per_rank = [float(x) for x in every.cpu()]
      ms = max(per_rank)
    return ms, launches, per_rank

  def free(self):
    import gc
    gc.collect()
    self.torch.cuda.empty_cache()


def peer_buffer(ctx, slots, dtype):
  """(PeerScales or None, note): NVLink peer-mapped gathered buffer, or why it is unavailable."""
  if ctx.world == 1 or os.environ.get("AEQB_BENCH_NCCL_GATHER"):
    return None, None
  try:
    from aeq_b200 import peer
    return peer.PeerScales(slots, ctx.dev, dtype=dtype), None
  except Exception as e:  # pylint: disable=broad-except
    return None, f"{type(e).__name__}: {e}"[:200]


# ------------------------------------------------------------------ configs[1]: fc4096_int8 (headline)
def run_fc4096(ctx, steps, warmup):
  torch, dist = ctx.torch, ctx.dist
  from aeq_b200 import device
  from oracle import aeq_oracle as O
  a, dev, world = ctx.a, ctx.dev, ctx.world
  T = a.tensors
  ws = ctx.weights([(ROWS, COLS)] * T, 1000 + ctx.rank)
  n_bytes = T * ROWS * COLS * 4
  # Per-channel scales of this rank's tensors live in ONE flat buffer (each tensor's scale output
  # is a view into it).  At N > 1 that buffer is this rank's row of a [world, T * ROWS] buffer every
  # rank maps (aeq_b200/peer.py): the requantisation kernel stores each row's scale into every
  # peer's copy from its own epilogue, so the all-gather of scales costs no launch.
  import torch as _t
  mirror, mirror_note = peer_buffer(ctx, T * ROWS, _t.float32)
  flat = mirror.local[:T * ROWS] if mirror is not None else torch.empty(T * ROWS, dtype=torch.float32, device=dev)
  gathered = torch.empty(world * T * ROWS, dtype=torch.float32, device=dev) if world > 1 else None
  if os.environ.get("AEQB_BENCH_SEPARATE_ALLOCS"):
    q_of = lambda i: torch.empty((ROWS, COLS), dtype=torch.int8, device=dev)
  else:
    q_flat = torch.empty(T * ROWS * COLS, dtype=torch.int8, device=dev)
    q_of = lambda i: q_flat[i * ROWS * COLS:(i + 1) * ROWS * COLS].view(ROWS, COLS)
  r8 = [device.Requantized(q_of(i), None,
                           flat[i * ROWS:(i + 1) * ROWS].view(ROWS, 1),
                           torch.empty((ROWS, 1), dtype=torch.int32, device=dev)) for i in range(T)]

  def step_int8():
    if mirror is not None:
      device.requant_rows_batch(ws, 8, True, outs=r8, mirror=mirror)
      return
    device.requant_rows_batch(ws, 8, True, outs=r8)
    if world > 1:
      dist.all_gather_into_tensor(gathered, flat)

  ms8, launches8, per_rank8 = ctx.timed(step_int8, steps, warmup)
  exchange_ok = None
  if mirror is not None:  # the in-kernel exchange against NCCL's all-gather of the same scales
    mirror.sync()
    dist.all_gather_into_tensor(gathered, flat)
    okt = torch.tensor([int(torch.equal(gathered.view(world, -1), mirror.gathered[:, :T * ROWS]))], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    exchange_ok = bool(okt.item())
  # exchange off: what the step costs without any scale exchange (attribution of the N > 1 loss)
  ms8_noex = None
  if world > 1:
    ms8_noex, _, _ = ctx.timed(lambda: device.requant_rows_batch(ws, 8, True, outs=r8), max(5, steps // 2), 2)
  # parity of the TIMED outputs: sampled tensors against the oracle, bit-exact
  parity = True
  for i in sorted({0, T - 1}):
    ref = O.minmax_requant(ws[i].cpu().numpy(), 8, True)
    parity &= bool(np.array_equal(r8[i].q.cpu().numpy(), ref["q"]))
    parity &= bool(np.array_equal(r8[i].scale.cpu().numpy(), ref["scale"]))

  # ---- roofline of the dominant kernel.  A step IS one launch of it (device job table: all T
  # tensors in one persistent launch; at N > 1 the scale exchange rides in the same kernel), so its
  # average launch duration over the TIMED region is this rank's ms per step.  Only when a step
  # carries something else (NCCL all-gather fallback, inline tables -> several launches) is the
  # kernel timed alone in a second loop.
  alg_bytes = (n_bytes / 4) * 5 + T * ROWS * 8  # 4 B read + 1 B written per weight, 8 B/row scale + zp
  step_is_kernel = (launches8 == steps) and world == 1
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  k_steps = max(3, min(steps, 20))
  l0 = ctx.lib.aeqb_launch_count()
  e0.record()
  for _ in range(k_steps):
    device.requant_rows_batch(ws, 8, True, outs=r8)
  e1.record()
  torch.cuda.synchronize()
  k_launches = ctx.lib.aeqb_launch_count() - l0
  again_ms = e0.elapsed_time(e1) / max(k_launches, 1)
  if step_is_kernel:
    k_ms, launches_per_step = per_rank8[ctx.rank], 1.0
    timed_how = "CUDA events over the timed region (one launch per step, nothing else in the step)"
  elif world > 1 and ms8_noex is not None and k_launches == k_steps:
    k_ms, launches_per_step = ms8_noex, 1.0
    timed_how = ("CUDA events over a timed region of the same launch without the scale exchange (max over ranks);"
                 " the step itself also carries the mirror launch")
  else:
    k_ms, launches_per_step = again_ms, k_launches / k_steps
    timed_how = "CUDA events around back-to-back launches after the timed region (the step also carries an exchange)"
  alg_per_launch = alg_bytes / launches_per_step
  achieved = alg_per_launch / k_ms / 1e6
  traffic, traffic_src = None, None
  try:
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["requant_rows_stream"]
    traffic = tj["dram_bytes_per_launch"] * (alg_per_launch / tj["algorithmic_bytes_per_launch"])
    traffic_src = ("profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the"
                   f" committed `ncu --set full` capture ({tj.get('capture', 'profiles/')}), rescaled by the"
                   " algorithmic bytes of this run's launch; not measured in this run")
  except Exception:  # pylint: disable=broad-except
    pass

  # ---- INT4 block-32 packed over the same stack
  b4 = device.requant_blocks_batch(ws, 32, 4)

  def step_int4():
    device.requant_blocks_batch(ws, 32, 4, outs=b4)

  ms4, launches4, _ = ctx.timed(step_int4, max(3, steps // 2), warmup)
  ref4 = O.minmax_requant(ws[T - 1].cpu().numpy(), 4, True, block=32)
  parity4 = bool(np.array_equal(b4[T - 1].packed.cpu().numpy(), O.pack_bits(4, ref4["q"]))) and bool(
      np.array_equal(b4[T - 1].scale_f16.cpu().numpy(), O.blockwise_scale_fp16(ref4["scale"])))
  del b4

  # ---- the 64-tensor stack of round 1 (one persistent launch per step), for continuity
  ms64, _, _ = ctx.timed(lambda: device.requant_rows_batch(ws[:64], 8, True, outs=r8[:64]), max(5, steps), 2)

  res = {
      "value": world * n_bytes / ms8 / 1e6, "ms_per_step": ms8, "launches": launches8, "per_rank_ms": per_rank8,
      "parity_checked": parity, "ms_per_step_without_exchange": ms8_noex,
      "exchange": ("none (N=1)" if world == 1 else
                   "per-channel scales pushed into every peer's gathered buffer over NVLink peer memory by our own"
                   " mirror launch behind the requantisation kernel (aeqb_requant_rows_batch_mirror_f32: 16-byte"
                   " stores, one packet per four scales); no collective"
                   if mirror is not None else "one NCCL all-gather of per-channel scales per step"),
      "scale_exchange_matches_nccl_all_gather": exchange_ok, "peer_mapping_error": mirror_note,
      "roofline": {"bound": "hbm", "achieved": achieved, "peak": ctx.peak, "unit": "GB/s",
                   "frac": achieved / ctx.peak, "traffic": traffic, "traffic_source": traffic_src,
                   "kernel": "requant_rows_stream<98304,16,12,2,false> (six 4096-float rows per 96 KiB tile, 16 consumer warps, 1 CTA per SM: the class a batch of >= 96 MiB takes)", "bytes_per_weight": 5.0,
                   "algorithmic_bytes_per_launch": alg_per_launch, "launch_ms": k_ms,
                   "launches_per_step": launches_per_step, "peak_source": ctx.peak_src, "timed": timed_how,
                   "relaunched_after_timed_region": {
                       "launch_ms": again_ms, "frac": alg_bytes * k_steps / max(k_launches, 1) / again_ms / 1e6 / ctx.peak,
                       "note": "the same launch repeated after the timed region and the parity downloads; on a"
                               " box that has reached its power cap by then (clocks.reasons: sw_power_cap, SM clock"
                               " falling towards ~1.35 GHz as in MEASURED_PEAKS.json clocks_under_load) this kernel"
                               " follows the SM clock, see DESIGN.md 4.1"}},
      "int4": {"value": world * n_bytes / ms4 / 1e6, "ms_per_step": ms4, "launches_per_step": launches4 / max(3, steps // 2),
               "roofline_frac": (n_bytes / 4) * 4.5625 / ms4 / 1e6 / ctx.peak,
               "achieved_hbm_gbs": (n_bytes / 4) * 4.5625 / ms4 / 1e6, "bytes_per_weight": 4.5625,
               "parity_checked": parity4},
      "stack64": {"value": world * 64 * ROWS * COLS * 4 / ms64 / 1e6, "ms_per_step": ms64, "tensors": 64,
                  "roofline_frac": 64 * ROWS * COLS * 5 / ms64 / 1e6 / ctx.peak,
                  "note": "round 1's default stack: 64 tensors = one persistent launch per step"},
  }
  if mirror is not None:
    mirror.close()
  return res, ws, r8


# ------------------------------------------------------------------ e2e (host-buffer C ABI)
def run_e2e(ctx, ws, r8, steps):
  torch, dist = ctx.torch, ctx.dist
  from aeq_b200 import host
  a, dev, world = ctx.a, ctx.dev, ctx.world
  Te = max(1, min(a.e2e_tensors, len(ws)))
  h_in = [host.pinned_empty((ROWS, COLS), np.float32) for _ in range(Te)]
  for i, h in enumerate(h_in):
    torch.from_numpy(h).copy_(ws[i])
  torch.cuda.synchronize()
  e_outs = host.requant_rows(h_in, 8, True, alloc=host.pinned_empty)
  e2e_steps = max(3, min(steps, 10))

  def wall(fn):
    for _ in range(2):
      fn()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
      fn()  # returns when every output is in host memory
    s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
      t = torch.tensor([s], device=dev)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      s = float(t.item())
    return s

  s_pin = wall(lambda: host.requant_rows(h_in, 8, True, outs=e_outs))
  nb = Te * ROWS * COLS * 4
  ok = bool((torch.from_numpy(e_outs[0][0]).to(dev) == r8[0].q).all())
  out = {"value": world * nb / s_pin / 1e9, "unit": UNIT, "h2d_bytes_per_step": nb,
         "d2h_bytes_per_step": Te * (ROWS * COLS + ROWS * 8), "tensors": Te, "ms_per_step": s_pin * 1e3,
         "matches_device_path": ok, "bound": "pcie",
         "api": "aeq_b200.host.requant_rows -> aeqb_host_requant_rows_batch_f32 (page-locked host buffers in and out)"}
  # ---- the boundary the reference presents: read-only views on an mmap'd file in, fresh NumPy out
  try:
    tmpdir = os.environ.get("AEQB_BENCH_TMP") or tempfile.gettempdir()
    with tempfile.NamedTemporaryFile(suffix=f".rank{ctx.rank}.bin", dir=tmpdir) as f:
      for h in h_in:
        f.write(h.tobytes())
      f.flush()
      mm = np.memmap(f.name, dtype=np.float32, mode="r", shape=(Te, ROWS, COLS))
      views = [mm[i] for i in range(Te)]
      got = host.requant_rows(views, 8, True)
      same = bool(np.array_equal(got[0][0], e_outs[0][0]) and np.array_equal(got[-1][0], e_outs[-1][0]))
      s_mm = wall(lambda: host.requant_rows(views, 8, True))
      out["pageable"] = {
          "value": world * nb / s_mm / 1e9, "ms_per_step": s_mm * 1e3, "matches_pinned_path": same,
          "inputs": "read-only np.memmap views of a file (the reference's tensors are views of the mmap'd"
                    " flatbuffer, tfl_flatbuffer_utils.py:254-263)", "outputs": "fresh NumPy arrays every call",
          "staging_threads": int(ctx.lib.aeqb_host_worker_threads()),
          "frac_of_pinned": s_pin / s_mm}
      del mm, views, got
  except Exception as e:  # pylint: disable=broad-except
    out["pageable"] = {"error": f"{type(e).__name__}: {e}"[:200]}
  # what bounds both: this box's host->device copy rate from pinned memory, measured with a
  # device->host copy of a quarter of the bytes running beside it (the arm's own 4 : 1 mix)
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
  out["h2d_gbs_measured"] = best
  out["frac_of_h2d"] = (out["value"] / world) / best if best > 0 else None
  if "value" in out.get("pageable", {}):
    out["pageable"]["frac_of_h2d"] = (out["pageable"]["value"] / world) / best if best > 0 else None
  return out


def run_e2e_quantizer(ctx, ws):
  """Quantizer(model.tflite, dynamic_wi8_afp32).quantize() on a synthetic 4 GiB FC model: file
  mmap -> recipe -> batched host pipeline -> QUANTIZE_TENSOR -> serialised bytes, wall clock."""
  from aeq_b200 import quantizer, recipe
  from aeq_b200.utils import tfl_model as T
  from tests import tfl_fixtures
  n = 64
  weights = [w.cpu().numpy() for w in ws[:n]]
  tmpdir = os.environ.get("AEQB_BENCH_TMP") or tempfile.gettempdir()
  path = os.path.join(tmpdir, f"aeqb_bench_fc{n}.tflite")
  try:
    with open(path, "wb") as f:
      f.write(T.write_model_to_bytes(tfl_fixtures.fc_stack(weights)))
    fsize = os.path.getsize(path)
    del weights
    best, stats, timings = 1e9, None, None
    for _ in range(2):
      t0 = time.perf_counter()
      qz = quantizer.Quantizer(path, recipe.dynamic_wi8_afp32())
      res = qz.quantize()
      dt = time.perf_counter() - t0
      if dt < best:
        best, stats, timings = dt, qz.prefetch_stats, dict(qz.timings, read_model=qz.read_seconds)
      out_bytes = len(res.quantized_model)
      del res, qz
    nb = n * ROWS * COLS * 4
    return {"value": nb / best / 1e9, "unit": UNIT, "seconds": best, "model_bytes": fsize,
            "quantized_model_bytes": out_bytes, "tensors": n, "prefetch": stats, "seconds_by_stage": timings,
            "api": "aeq_b200.quantizer.Quantizer(path, recipe.dynamic_wi8_afp32()).quantize()"}
  finally:
    try:
      os.unlink(path)
    except OSError:
      pass


# ------------------------------------------------------------------ configs[2]: gemma2b_int4b32
def run_gemma(ctx, steps, warmup):
  """Blockwise INT4 (block 32, packed) over Gemma-2B-shaped FC weights.  The global tensor list is
  world x 18 decoder layers; `sharding.assign_tensors` (LPT by bytes) deals it to the ranks, every
  rank requantises what it owns (weak scaling: one Gemma-2B of weights per GPU) and every block's
  fp16 scale is stored into all peers' gathered buffers from the kernel's epilogue."""
  torch, dist = ctx.torch, ctx.dist
  from aeq_b200 import device, sharding
  from oracle import aeq_oracle as O
  world, dev = ctx.world, ctx.dev
  shapes = GEMMA_LAYER * (18 * world)
  sizes = [r * c * 4 for r, c in shapes]
  owner = sharding.assign_tensors(sizes, world)
  mine = sharding.owned(owner, ctx.rank)
  ws = ctx.weights([shapes[i] for i in mine], 4242 + ctx.rank)
  n_bytes_all = sum(sizes)
  counts = [ (w.numel() // 32 + 7) // 8 * 8 for w in ws]  # 16-byte aligned starts inside the rank's row
  slots = max(sum((shapes[i][0] * shapes[i][1] // 32 + 7) // 8 * 8 for i in sharding.owned(owner, r))
              for r in range(world))
  mirror, note = peer_buffer(ctx, slots, torch.float16)
  flat = mirror.local if mirror is not None else torch.empty(slots, dtype=torch.float16, device=dev)
  outs, off = [], 0
  for w, c in zip(ws, counts):
    outs.append(device.Requantized(None, torch.empty(w.numel() // 2, dtype=torch.uint8, device=dev), None, None,
                                   flat[off:off + w.numel() // 32].view(w.shape[0], -1)))
    off += c
  gathered = torch.empty(world * flat.numel(), dtype=torch.float16, device=dev) if world > 1 else None

  def step():
    if mirror is not None:
      device.requant_blocks_batch(ws, 32, 4, outs=outs, mirror=mirror)
      return
    device.requant_blocks_batch(ws, 32, 4, outs=outs)
    if world > 1:
      dist.all_gather_into_tensor(gathered, flat)

  ms, launches, per_rank = ctx.timed(step, steps, warmup)
  ok, ms_noex, ms_nccl = None, None, None
  if world > 1:
    if mirror is not None:
      mirror.sync()
      dist.all_gather_into_tensor(gathered, flat)
      okt = torch.tensor([int(torch.equal(gathered.view(world, -1).view(torch.int16),
                                          mirror.gathered.view(torch.int16)))], device=dev)
      dist.all_reduce(okt, op=dist.ReduceOp.MIN)
      ok = bool(okt.item())
    ms_noex, _, _ = ctx.timed(lambda: device.requant_blocks_batch(ws, 32, 4, outs=outs), steps, 2)

    def step_nccl():
      device.requant_blocks_batch(ws, 32, 4, outs=outs)
      dist.all_gather_into_tensor(gathered, flat)
    ms_nccl, _, _ = ctx.timed(step_nccl, steps, 2)
  i = len(ws) - 1
  ref = O.minmax_requant(ws[i].cpu().numpy(), 4, True, block=32)
  parity = bool(np.array_equal(outs[i].packed.cpu().numpy(), O.pack_bits(4, ref["q"]))) and bool(
      np.array_equal(outs[i].scale_f16.cpu().numpy(), O.blockwise_scale_fp16(ref["scale"])))
  my_bytes = sum(w.numel() for w in ws) * 4
  res = {"value": n_bytes_all / ms / 1e6, "unit": UNIT, "ms_per_step": ms, "per_rank_ms": per_rank,
         "launches_per_step": launches / steps, "tensors_per_rank": len(ws), "fp32_bytes": n_bytes_all,
         "scaling": "weak (one Gemma-2B FC set, 18 layers x 7 tensors = 7.9 GB fp32, per GPU)",
         "imbalance": sharding.imbalance(sizes, owner, world), "bytes_per_weight": 4.5625,
         "roofline_frac": (my_bytes / 4) * 4.5625 / ms / 1e6 / ctx.peak, "parity_checked": parity,
         "exchange": ("none (N=1)" if world == 1 else
                      "fp16 block scales stored into every peer's gathered buffer from the kernel's epilogue"
                      " (aeqb_requant_blocks_batch_mirror_f32), 64 B per peer and consumer warp, sixteen per 64 KiB tile" if mirror is not None
                      else "one NCCL all-gather of fp16 block scales per step"),
         "exchange_bytes_per_rank_per_step": (world - 1) * (my_bytes // 4 // 32) * 2,
         "scale_exchange_matches_nccl_all_gather": ok, "ms_per_step_without_exchange": ms_noex,
         "ms_per_step_with_nccl_all_gather": ms_nccl, "peer_mapping_error": note}
  if mirror is not None:
    mirror.close()
  return res


# ------------------------------------------------------------------ configs[3]: calib512
def run_calib(ctx, ws, steps, warmup):
  """static_wi8_ai8 calibration: per-batch min / max with the (-3e38, 3e38) filter over 512
  activation batches [8, 512, 4096] (64 MiB each), batches sharded over the ranks, ONE all-gather of
  the 512 (min, max) pairs, then the fp32 EMA 0.95 replicated in batch order on every rank."""
  torch, dist = ctx.torch, ctx.dist
  from aeq_b200 import device
  from oracle import aeq_oracle as O
  world, dev = ctx.world, ctx.dev
  n_batches = 512
  per = n_batches // world
  if per * world != n_batches:
    return {"skipped": f"512 batches do not divide over {world} ranks"}
  # the activation batches are views of the resident fp32 stack ([4096,4096] == [8,512,4096] = 64 MiB)
  acts = [ws[i % len(ws)].view(8, 512, 4096) for i in range(per)]
  local = torch.empty((per, 2), dtype=torch.float32, device=dev)
  pairs = torch.empty((n_batches, 2), dtype=torch.float32, device=dev)
  state = {}

  def step():
    local.copy_(device.minmax_tensors(acts, -3e38, 3e38))  # 64 batches per launch
    if world > 1:
      dist.all_gather_into_tensor(pairs.view(-1), local.view(-1))  # rank r holds batches [r*per, (r+1)*per)
      state["qsv"] = device.ema_sequence(pairs)
    else:
      state["qsv"] = device.ema_sequence(local)

  ms, launches, per_rank = ctx.timed(step, steps, warmup)
  src = (pairs if world > 1 else local).cpu().numpy()
  want = O.ema_sequence([np.full((1, 1, 1), v, np.float32) for v in src[:, 0]],
                        [np.full((1, 1, 1), v, np.float32) for v in src[:, 1]])
  got = state["qsv"].cpu().numpy()
  parity = bool(got[0] == want[0].item() and got[1] == want[1].item())
  a0 = acts[0].cpu().numpy()
  mn, mx = O.activation_minmax(a0)
  parity &= bool(src[ctx.rank * per if world > 1 else 0, 0] == mn.item() and src[ctx.rank * per if world > 1 else 0, 1] == mx.item())
  nb = n_batches * 8 * 512 * 4096 * 4
  return {"value": nb / ms / 1e6, "unit": "activation GB/s", "ms_per_step": ms, "per_rank_ms": per_rank,
          "batches": n_batches, "batches_per_rank": per, "launches_per_step": launches / steps,
          "scaling": "strong (512 batches in total)", "bytes_per_element": 4.0,
          "roofline_frac": (nb / world) / ms / 1e6 / ctx.peak, "parity_checked": parity,
          "exchange": "none (N=1)" if world == 1 else
                      "one NCCL all-gather of 512 (min, max) fp32 pairs per step, then aeqb_ema_sequence_f32 on every rank"}


# ------------------------------------------------------------------ configs[4]: llama7b_gptq
def run_gptq(ctx, steps, warmup):
  """Hadamard-rotated INT4 + GPTQ on Llama-7B-shaped FC weights: one decoder layer per rank (weak
  scaling over layers).  Calibration tokens are data-parallel, as calibration inference is: every
  rank holds tokens / world of EVERY layer's four FC inputs, computes the partial X^T X (tcgen05
  3xTF32) and the float64 partials are summed on the layer's owner (NCCL reduce).  The owner then
  rotates weights and Hessians, inverts and runs the OBS loops of its seven weights."""
  torch, dist = ctx.torch, ctx.dist
  from aeq_b200 import device
  from aeq_b200.algorithms.uniform_quantize import hadamard_gptq
  world, dev, rank = ctx.world, ctx.dev, ctx.rank
  tokens = ctx.a.gptq_tokens
  t_local = tokens // world
  layer = ctx.weights(LLAMA_LAYER, 777 + rank)
  lbytes = sum(w.numel() for w in layer) * 4
  gen = torch.Generator(device=dev).manual_seed(99 + rank)
  xs = {}
  for l in range(world):  # this rank's token shard of every layer's inputs
    for name, k in LLAMA_INPUTS:
      xs[(l, name)] = torch.randn(t_local, k, device=dev, generator=gen)
  num_samples = 128  # sequences per Hessian (alpha = 2 / num_samples, gptq.py:105)
  parts = {"hessian": 0.0, "exchange": 0.0, "quantize": 0.0}
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
  state = {}
  # One stream for the layer's seven problems (each factorisation and OBS loop still runs its own lookahead
  # side stream) is what is timed: a stream per problem measured no faster (87.7 against 82.3 ms) or much
  # slower (188 against 82 ms) from run to run; AEQB_BENCH_GPTQ_CONCURRENT=1 times that arrangement instead.
  conc = bool(os.environ.get("AEQB_BENCH_GPTQ_CONCURRENT"))

  def step():
    ev[0].record()
    mine = {}
    hs = {}
    for l in range(world):
      for name, k in LLAMA_INPUTS:
        hs[(l, name)] = device.xtx(xs[(l, name)], 2.0 / num_samples)
    ev[1].record()
    if world > 1:
      for l in range(world):
        for name, k in LLAMA_INPUTS:
          dist.reduce(hs[(l, name)], dst=l, op=dist.ReduceOp.SUM)
    for name, k in LLAMA_INPUTS:
      mine[name] = hs[(rank, name)]
    del hs
    ev[2].record()
    # rotate W and H, invert, OBS loops of the layer's seven problems.  A failure on one rank must not take
    # that rank out of the step sequence (the others would wait in the next reduce for ever): it is kept and
    # raised after the timed region, where every rank leaves together.
    try:
      res = hadamard_gptq.quantize_layer_device(layer, LLAMA_FEEDS, mine, 4, True, 4096, 0.01, concurrent=conc)
      state["res"] = res
    except Exception as e:  # pylint: disable=broad-except
      state.setdefault("errors", []).append(f"{type(e).__name__}: {e}")
      state["failed_h"] = mine
    ev[3].record()
    state["h"] = mine

  ms, launches, per_rank = ctx.timed(step, steps, warmup)
  torch.cuda.synchronize()
  if state.get("errors"):
    # what failed, and whether the same Hessians go through one stream at a time (a timing fault) or not
    diag = []
    for name, h in state["failed_h"].items():
      d = torch.diagonal(h)
      diag.append(f"{name}: finite={bool(torch.isfinite(h).all())} diag=[{float(d.min()):.4g},{float(d.max()):.4g}]")
    try:
      hadamard_gptq.quantize_layer_device(layer, LLAMA_FEEDS, state["failed_h"], 4, True, 4096, 0.01, concurrent=False)
      torch.cuda.synchronize()
      retry = "the same Hessians pass on one stream afterwards"
    except Exception as e:  # pylint: disable=broad-except
      retry = f"the same Hessians fail again on one stream: {e}"
    raise RuntimeError(f"rank {rank}: {len(state['errors'])} of {steps + warmup} steps failed ({state['errors'][0]});"
                       f" {'; '.join(diag)}; {retry}")
  for key, i in (("hessian", 0), ("exchange", 1), ("quantize", 2)):
    parts[key] = ev[i].elapsed_time(ev[i + 1])
  # the other arrangement of the same seven problems (every inverse and OBS loop on its own stream, or one
  # after the other on one stream), for the record
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ms_other, other_err = None, None
  try:
    e0.record()
    hadamard_gptq.quantize_layer_device(layer, LLAMA_FEEDS, state["h"], 4, True, 4096, 0.01, concurrent=not conc)
    e1.record()
    torch.cuda.synchronize()
    ms_other = e0.elapsed_time(e1)
  except Exception as e:  # pylint: disable=broad-except
    other_err = f"{type(e).__name__}: {e}"
  ms_serial = parts["quantize"] if not conc else ms_other
  ms_conc = parts["quantize"] if conc else ms_other
  # sanity of the timed outputs: proxy loss of the o-projection against plain rounding of the same
  # rotated weight (GPTQ must win), and H_rot @ Hinv = I
  q, scale, _, n_o = state["res"][3]
  r = device.hadamard_rows(layer[3], n_o)
  h_rot = hadamard_gptq.rotate_hessian_device(state["h"]["attn_out"], n_o)
  d = torch.diagonal(h_rot)
  d = torch.where(d == 0, torch.ones_like(d), d)
  hd = h_rot.clone()
  hd.diagonal().copy_(d + 0.01 * d.mean())  # the damped rotated Hessian (gptq.py:111-118)
  rtn = torch.clamp(torch.round(r / scale), -8, 7)

  def loss(qq):
    e = r.double() - qq.double() * scale.double()
    return float(((e @ hd) * e).sum())
  l_gptq, l_rtn = loss(q), loss(rtn)
  return {"value": world * lbytes / ms / 1e6, "unit": UNIT, "ms_per_step": ms, "per_rank_ms": per_rank,
          "launches_per_step": launches / steps, "layers": world, "tokens_per_hessian": tokens,
          "tokens_per_rank_per_hessian": t_local, "fp32_bytes": world * lbytes,
          "scaling": "weak (one Llama-7B decoder layer = 7 FC weights per GPU)",
          "ms_hessians": parts["hessian"], "ms_exchange": parts["exchange"],
          "ms_rotate_inverse_obs": parts["quantize"], "ms_rotate_inverse_obs_one_stream": ms_serial,
          "ms_rotate_inverse_obs_stream_per_problem": ms_conc, "stream_per_problem_error": other_err,
          "concurrency": ("timed: 4 rotated-Hessian inverses and 7 OBS loops on their own streams" if conc else
                          "timed: the seven problems one after the other on the caller's stream (lookahead side"
                          " streams inside each factorisation and OBS loop)")
                         + " (hadamard_gptq.quantize_layer_device)",
          "hessian_tflops_fp32_equivalent": 2.0 * t_local * world * sum(k * k for _, k in LLAMA_INPUTS) / parts["hessian"] / 1e9,
          "proxy_loss_vs_round_to_nearest": l_gptq / l_rtn, "parity_checked": bool(l_gptq < l_rtn),
          "exchange": "none (N=1)" if world == 1 else
                      "float64 partial X^T X of every layer summed on the layer's owner (NCCL reduce, 4 per layer:"
                      " 3 x 134 MB + 969 MB)",
          "algorithm": "aeq_b200.algorithms.uniform_quantize.hadamard_gptq (rotate W, rotate H on both sides,"
                       " damped inverse, OBS loop), INT4 per-channel, blocksize 64, damp 0.01"}


# ------------------------------------------------------------------ the other rows of the path
def extra_modes(ctx, ws, steps):
  """Device-resident timings of the remaining hot-path rows on the same stack."""
  torch = ctx.torch
  from aeq_b200 import device
  dev, peak = ctx.dev, ctx.peak
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
                        ("octav_int8_perchannel", octav(8, 0), 9.0),
                        ("octav_int4_block32_packed", octav(4, 32), 8.5625 + 10 * 4 / 32)):
    ms = timeit(fn, reps)
    out[name] = {"value": n_bytes / ms / 1e6, "ms_per_step": ms, "tensors": T, "bytes_per_weight": bpw,
                 "roofline_frac": (n_bytes / 4) * bpw / ms / 1e6 / peak}

  def octav_only():
    for w in ws[:T]:
      device.octav_clip_rows(w, 4)
  ms = timeit(octav_only, reps)
  out["octav_clip_rows_only"] = {"value": n_bytes / ms / 1e6, "us_per_tensor": ms * 1e3 / T, "tensors": T,
                                 "bytes_per_weight": 4.0, "roofline_frac": n_bytes / ms / 1e6 / peak}

  def mse():
    for w in ws[:T]:
      device.requant_mse_rows(w, 8, 0.05408)
  ms = timeit(mse, reps)
  out["mse_int8_perchannel"] = {"value": n_bytes / ms / 1e6, "ms_per_step": ms, "tensors": T,
                                "bytes_per_weight": 5.0, "roofline_frac": (n_bytes / 4) * 5.0 / ms / 1e6 / peak}
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
  acts = [w.view(8, 512, 4096) for w in ws[:T]]

  def calib():
    for x in acts:
      device.minmax_tensor(x, -3e38, 3e38)
  ms = timeit(calib, reps)
  out["calibration_minmax_one_per_launch"] = {"value": n_bytes / ms / 1e6, "unit": "activation GB/s",
                                              "ms_per_step": ms, "batches": T,
                                              "roofline_frac": n_bytes / ms / 1e6 / peak}
  # GPTQ on one [4096, 4096] layer: Hessian from 16384 tokens, damped inverse, OBS loop
  x = torch.randn(16384, COLS, device=dev)
  ms_h = timeit(lambda: device.xtx(x, 2.0 / 8), 3)
  h = device.xtx(x, 2.0 / 8)
  ms_inv = timeit(lambda: device.hessian_inverse(h, 0.01), 2)
  hinv = device.hessian_inverse(h, 0.01)
  w = ws[0]
  sc = w.abs().amax(dim=1) / 7.0
  ms_q = timeit(lambda: device.gptq_quantize(w, hinv, sc, None, 0, 4, True), 2)
  out["gptq_int4_4096x4096"] = {
      "value": ROWS * COLS * 4 / (ms_inv + ms_q) / 1e6, "ms_hessian_16384_tokens": ms_h,
      "hessian_tflops": 2.0 * 16384 * COLS * COLS / ms_h / 1e9,
      "ms_hessian_inverse": ms_inv, "ms_obs_loop": ms_q,
      "obs_loop_tflops": 1.0 * ROWS * COLS * COLS / ms_q / 1e9,
      "note": "value = fp32 weight bytes / (inverse + OBS loop) for one layer, Hessian given"}
  del x, h, hinv
  llama = ctx.weights(LLAMA_LAYER * 4, 4242)
  n_b = sum(t.numel() for t in llama) * 4
  for bits, bpw in ((8, 5.0), (4, 4.5)):
    st = {}

    def rows():
      st["o"] = device.requant_rows_batch(llama, bits, True, want_q=(bits == 8), want_packed=(bits == 4),
                                           outs=st.get("o"))
    ms = timeit(rows, reps)
    out[f"llama7b_shapes_int{bits}_perchannel"] = {
        "value": n_b / ms / 1e6, "ms_per_step": ms, "tensors": len(llama), "fp32_bytes": n_b,
        "bytes_per_weight": bpw, "roofline_frac": (n_b / 4) * bpw / ms / 1e6 / peak}
    del st
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


def guarded(fn, *args):
  try:
    return fn(*args)
  except Exception as e:  # pylint: disable=broad-except
    import traceback
    traceback.print_exc(file=sys.stderr)
    return {"error": f"{type(e).__name__}: {e}"[:300]}


def main():
  a = parse()
  quiet_stdout()
  if a.impl == "reference":
    run_reference(a)
    return
  ctx = Ctx(a)
  world, rank = ctx.world, ctx.rank
  warm = max(a.warmup, 3)
  sampler = ClockSampler(ctx.local)
  if rank == 0:
    sampler.start()
  if os.environ.get("AEQB_BENCH_ONLY") == "llama7b_gptq":  # diagnosis: this mode alone, not a bench line
    res = guarded(run_gptq, ctx, max(2, min(a.steps, 10)), 1)
    if rank == 0:
      sampler.stop()
      emit({"only": "llama7b_gptq", "n_gpus": world, "result": res})
    if world > 1:
      ctx.dist.destroy_process_group()
    return
  fc, ws, r8 = run_fc4096(ctx, a.steps, warm)
  clocks = sampler.stop() if rank == 0 else None
  e2e = run_e2e(ctx, ws, r8, a.steps)
  if world == 1 and a.modes != "headline":
    e2e["quantizer"] = guarded(run_e2e_quantizer, ctx, ws)
  modes = {
      "int8_perchannel": {"value": fc["value"], "ms_per_step": fc["ms_per_step"],
                          "launches_per_step": fc["launches"] / a.steps, "per_rank_ms": fc["per_rank_ms"],
                          "ms_per_step_without_exchange": fc["ms_per_step_without_exchange"],
                          "parity_checked": fc["parity_checked"]},
      "int4_block32_packed": fc["int4"],
      "int8_perchannel_64_tensor_stack": fc["stack64"],
  }
  if a.modes != "headline":
    modes["calib512"] = guarded(run_calib, ctx, ws, max(3, min(a.steps, 10)), 2)
  extra = {}
  if a.modes == "all" and world == 1:
    extra = guarded(extra_modes, ctx, ws, a.steps)
  del ws, r8
  ctx.free()
  if a.modes != "headline":
    modes["gemma2b_int4b32"] = guarded(run_gemma, ctx, max(5, min(a.steps, 50)), 3)
    ctx.free()
    modes["llama7b_gptq"] = guarded(run_gptq, ctx, 2, 1)
    ctx.free()
  modes.update(extra if isinstance(extra, dict) else {"extra_error": extra})

  cpu = None
  if rank == 0 and not a.no_cpu_baseline:
    c = cpu_baseline(a.cpu_sample)
    cpu = {"value": c["int8"], "unit": UNIT, "cores": c["threads"], "kind": spec_kind(c["kind"]),
           "source": kind_source(c["kind"]),
           "sample": f"{a.cpu_sample} of the workload's [{ROWS},{COLS}] tensors, one pass,"
                     f" naive_min_max_quantize.get_tensor_quant_params of {kind_text(c['kind'])}, {c['threads']}"
                     " threads over independent tensors",
           "int4_block32_packed": c["int4"], "single_thread_value": c["int8_single_thread"],
           "host_cores": os.cpu_count()}

  # which config is the line's value
  head = {"fc4096_int8": None, "gemma2b_int4b32": "gemma2b_int4b32", "calib512": "calib512",
          "llama7b_gptq": "llama7b_gptq"}[a.workload]
  value, ms = fc["value"], fc["ms_per_step"]
  unit = UNIT
  if head and "value" in modes.get(head, {}):
    value, ms, unit = modes[head]["value"], modes[head]["ms_per_step"], modes[head].get("unit", UNIT)
  if rank == 0:
    cfg = config_of(a)
    if head:
      cfg["workload"] = f"{a.workload} (see modes.{head})"
    emit(({
        "metric": METRIC, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "exchange": {"kind": fc["exchange"],
                     "scale_exchange_matches_nccl_all_gather": fc["scale_exchange_matches_nccl_all_gather"],
                     "peer_mapping_error": fc["peer_mapping_error"]},
        "parity_checked": bool(fc["parity_checked"] and fc["int4"]["parity_checked"]),
        "per_rank_ms": fc["per_rank_ms"],
        "modes": modes,
        "roofline": fc["roofline"],
        "e2e": e2e,
        "cpu_baseline": cpu,
        "gpu_launches": int(fc["launches"]),
        "clocks": clocks,
    }))
  if world > 1:
    ctx.dist.destroy_process_group()


if __name__ == "__main__":
  main()
