"""Host-buffer pipeline at the boundary the reference presents (SURVEY.md §8b "Ownership"):
inputs are read-only NumPy views on an mmap'd file, outputs fresh NumPy arrays.

  python tools/host_bench.py [--tensors 16] [--threads 8] [--devices all]

Prints GB/s of fp32 weight bytes for: pinned in/out (DMA in place), pageable NumPy in/out,
np.memmap views in / fresh NumPy out, and the same fanned out over every visible GPU; beside
the box's own pinned H2D rate.  One JSON line at the end.
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")):
  if p not in sys.path:
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--tensors", type=int, default=16)
  ap.add_argument("--threads", type=int, default=0)
  ap.add_argument("--devices", default="")
  ap.add_argument("--reps", type=int, default=3)
  a = ap.parse_args()
  if a.threads:
    os.environ["AEQB_HOST_THREADS"] = str(a.threads)
  import torch
  from aeq_b200 import _lib, host
  R = C = 4096
  T = a.tensors
  nbytes = T * R * C * 4
  rng = np.random.default_rng(0)
  base = (rng.standard_normal((R, C), dtype=np.float32) * 0.02)
  out = {"tensors": T, "fp32_bytes": nbytes, "host_cores": os.cpu_count(),
         "affinity": len(os.sched_getaffinity(0))}

  def timeit(fn):
    fn()
    best = 1e9
    for _ in range(a.reps):
      t0 = time.perf_counter()
      fn()
      best = min(best, time.perf_counter() - t0)
    return nbytes / best / 1e9

  # pinned in / out
  pin = [host.pinned_empty((R, C), np.float32) for _ in range(T)]
  for p in pin:
    p[...] = base
  outs = host.requant_rows(pin, 8, True, alloc=host.pinned_empty)
  out["pinned"] = timeit(lambda: host.requant_rows(pin, 8, True, outs=outs))
  want_q = outs[0][0].copy()
  # pageable NumPy in / out (outputs reused)
  pag = [base.copy() for _ in range(T)]
  outs_p = host.requant_rows(pag, 8, True)
  out["pageable"] = timeit(lambda: host.requant_rows(pag, 8, True, outs=outs_p))
  assert np.array_equal(outs_p[0][0], want_q)
  # pageable in, FRESH outputs every call (what get_tensor_quant_params returns)
  out["pageable_fresh_outputs"] = timeit(lambda: host.requant_rows(pag, 8, True))
  # memmap views in (read-only), fresh outputs
  with tempfile.NamedTemporaryFile(suffix=".bin", dir=os.environ.get("TMPDIR", "/tmp")) as f:
    for _ in range(T):
      f.write(base.tobytes())
    f.flush()
    mm = np.memmap(f.name, dtype=np.float32, mode="r", shape=(T, R, C))
    views = [mm[i] for i in range(T)]
    got = host.requant_rows(views, 8, True)
    assert np.array_equal(got[0][0], want_q) and np.array_equal(got[-1][0], want_q)
    out["memmap_fresh_outputs"] = timeit(lambda: host.requant_rows(views, 8, True))
    out["memmap_int4_block32_packed"] = timeit(
        lambda: host.requant_blocks(views, 32, 4, want_q=False, want_packed=True, want_scale=False,
                                    want_scale_f16=True))
    if a.devices:
      devs = host.set_devices("all" if a.devices == "all" else [int(x) for x in a.devices.split(",")])
      out["devices"] = devs
      got = host.requant_rows(views, 8, True)
      assert np.array_equal(got[0][0], want_q) and np.array_equal(got[-1][0], want_q)
      out["memmap_fresh_outputs_multi_gpu"] = timeit(lambda: host.requant_rows(views, 8, True))
      out["pinned_multi_gpu"] = timeit(lambda: host.requant_rows(pin, 8, True, outs=outs))
      host.set_devices(None)
    del mm, views
  out["worker_threads"] = _lib.load().aeqb_host_worker_threads()
  # staged single copies (hostio): pageable -> device -> pageable
  from aeq_b200 import hostio
  big = np.concatenate([base] * 4)  # 256 MiB
  hostio.to_device(big)
  t0 = time.perf_counter()
  d = hostio.to_device(big)
  torch.cuda.synchronize()
  out["copy_in_gbs"] = big.nbytes / (time.perf_counter() - t0) / 1e9
  hostio.to_host(d)
  t0 = time.perf_counter()
  back = hostio.to_host(d)
  out["copy_out_gbs"] = big.nbytes / (time.perf_counter() - t0) / 1e9
  assert np.array_equal(back, big)
  t0 = time.perf_counter()
  d2 = torch.from_numpy(big).to("cuda")
  torch.cuda.synchronize()
  out["torch_pageable_to_gbs"] = big.nbytes / (time.perf_counter() - t0) / 1e9
  # the link itself
  pin_in = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
  d_in = torch.empty_like(pin_in, device="cuda")
  best = 0.0
  for _ in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d_in.copy_(pin_in, non_blocking=True)
    torch.cuda.synchronize()
    best = max(best, pin_in.numel() / (time.perf_counter() - t0) / 1e9)
  out["h2d_pinned_gbs"] = best
  # one-thread memcpy rate of this host (what the old staging path was bound by)
  dst = np.empty_like(big)
  t0 = time.perf_counter()
  np.copyto(dst, big)
  out["memcpy_one_thread_gbs"] = big.nbytes / (time.perf_counter() - t0) / 1e9
  print(json.dumps(out))


if __name__ == "__main__":
  main()
