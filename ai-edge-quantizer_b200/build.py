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
      h.update(p.encode() + b"\0" + f.read())
  h.update(" ".join(FLAGS).encode())
  return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
  dig = _digest()
  if (not force and os.path.exists(OUT) and os.path.exists(STAMP)
      and open(STAMP).read().strip() == dig):
    return OUT
  cmd = [NVCC, *FLAGS, "-o", OUT, *sources()]
  if verbose:
    cmd.insert(1, "-Xptxas")
    cmd.insert(2, "-v")
    print(" ".join(cmd), file=sys.stderr)
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
