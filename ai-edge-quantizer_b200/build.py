"""Builds libaeqb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Usage: python ai-edge-quantizer_b200/build.py [--force] [--verbose]
The output lands next to the Python package (aeq_b200/libaeqb200.so) so that it
travels with the repo snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "aeq_b200", "libaeqb200.so")
STAMP = OUT + ".stamp"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-cudart", "static",
    "--expt-relaxed-constexpr",
]


def sources() -> list[str]:
  return sorted(
      os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
  h = hashlib.sha256()
  for p in sources() + sorted(
      os.path.join(CSRC, f) for f in os.listdir(CSRC)
      if f.endswith((".h", ".cuh"))) + [
          os.path.join(HERE, "..", "include", "aeqb200.h"), __file__]:
    with open(p, "rb") as f:
      h.update(os.path.basename(p).encode() + b"\0" + f.read())  # location-independent
  h.update(" ".join(FLAGS).encode())
  return h.hexdigest()


COMPILE_FLAGS = [f for f in FLAGS if f not in ("--shared", "-cudart", "static")]
OBJ_DIR = os.path.join(HERE, "build")


def _file_digest(src: str) -> str:
  h = hashlib.sha256()
  deps = [src] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))) + [
      os.path.join(HERE, "..", "include", "aeqb200.h")]
  for p in deps:
    with open(p, "rb") as f:
      h.update(os.path.basename(p).encode() + b"\0" + f.read())  # location-independent
  h.update(" ".join(COMPILE_FLAGS).encode())
  return h.hexdigest()


def _compile_one(src: str, force: bool, verbose: bool) -> str:
  """One translation unit -> build/<name>.o, skipped when source, headers and flags are unchanged."""
  obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
  stamp = obj + ".stamp"
  dig = _file_digest(src)
  if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
    return obj
  cmd = [NVCC, *COMPILE_FLAGS, "-c", "-o", obj, src]
  if verbose:
    cmd[1:1] = ["-Xptxas", "-v"]
    print(" ".join(cmd), file=sys.stderr)
  subprocess.run(cmd, check=True)
  with open(stamp, "w") as f:
    f.write(dig)
  return obj


def build(force: bool = False, verbose: bool = False) -> str:
  """Compiles the translation units in parallel (one nvcc each) and links the shared library."""
  from concurrent.futures import ThreadPoolExecutor
  dig = _digest()
  if (not force and os.path.exists(OUT) and os.path.exists(STAMP)
      and open(STAMP).read().strip() == dig):
    return OUT
  os.makedirs(OBJ_DIR, exist_ok=True)
  with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
    objs = list(pool.map(lambda s: _compile_one(s, force, verbose), sources()))
  cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler",
         "-fPIC,-fvisibility=hidden", "-cudart", "static", "-o", OUT, *objs]
  subprocess.run(cmd, check=True)
  with open(STAMP, "w") as f:
    f.write(dig)
  return OUT


if __name__ == "__main__":
  ap = argparse.ArgumentParser()
  ap.add_argument("--force", action="store_true")
  ap.add_argument("--verbose", action="store_true")
  a = ap.parse_args()
  print(build(a.force, a.verbose))
