"""Makes the UNMODIFIED reference travel: copies its importable Python sources into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (the header oracle/aeq_oracle.py carries applies here too): nothing
under aeq_b200/ imports oracle/_ref; only tests/, __graft_entry__.smoke() and bench.py's
`--impl reference` / `cpu_baseline` legs do, as the checker and the timed CPU baseline.

The reference is pure Python, so "building" it is a file copy: every non-test `.py` (and the
recipe / policy `.json` files the package loads) of /root/reference/ai_edge_quantizer lands in
oracle/_ref/ai_edge_quantizer/, byte for byte, with a MANIFEST.json of sha256 digests.
oracle/_ref/ is git-ignored (the copies never enter this repo's history) but not
gpurun-ignored, so it ships to the GPU box with the snapshot exactly like the built `.so`.
There `oracle/refshim` finds it (no /root/reference on that box) and the reference's own
`naive_min_max_quantize.get_tensor_quant_params`, `transformation_utils.pack_data`,
`algorithm_manager`, ... run unmodified next to the device path.

  python oracle/make_ref.py [--reference /root/reference] [--check]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
PKG = "ai_edge_quantizer"
SKIP_DIRS = {"tests", "examples", "experimental", "__pycache__"}


def _wanted(rel: str) -> bool:
  base = os.path.basename(rel)
  if base.endswith("_test.py") or base in ("conftest.py", "test_utils.py"):
    return False
  return base.endswith((".py", ".json")) or base in ("VERSION", "LICENSE")


def files(reference_root: str) -> list[str]:
  src = os.path.join(reference_root, PKG)
  out = []
  for d, dirs, names in os.walk(src):
    dirs[:] = sorted(x for x in dirs if x not in SKIP_DIRS)
    for n in sorted(names):
      rel = os.path.relpath(os.path.join(d, n), reference_root)
      if _wanted(rel):
        out.append(rel)
  for extra in ("VERSION", "LICENSE"):
    if os.path.exists(os.path.join(reference_root, extra)):
      out.append(extra)
  return out


def _sha(path: str) -> str:
  with open(path, "rb") as f:
    return hashlib.sha256(f.read()).hexdigest()


def make(reference_root: str = "/root/reference", out: str = OUT) -> dict:
  """Copies the sources; returns the manifest {relative path: sha256}."""
  if not os.path.isdir(os.path.join(reference_root, PKG)):
    raise FileNotFoundError(f"no {PKG}/ under {reference_root}")
  manifest = {}
  tmp = out + ".tmp"
  shutil.rmtree(tmp, ignore_errors=True)
  for rel in files(reference_root):
    dst = os.path.join(tmp, rel)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(os.path.join(reference_root, rel), dst)
    manifest[rel] = _sha(dst)
  version = ""
  if os.path.exists(os.path.join(reference_root, "VERSION")):
    version = open(os.path.join(reference_root, "VERSION")).read().strip()
  with open(os.path.join(tmp, "MANIFEST.json"), "w") as f:
    json.dump({"source": reference_root, "version": version, "files": manifest}, f, indent=1, sort_keys=True)
  shutil.rmtree(out, ignore_errors=True)
  os.replace(tmp, out)
  return manifest


def check(out: str = OUT) -> bool:
  """True when oracle/_ref exists and every file still matches its manifest digest."""
  try:
    m = json.load(open(os.path.join(out, "MANIFEST.json")))["files"]
  except Exception:
    return False
  return all(os.path.exists(os.path.join(out, rel)) and _sha(os.path.join(out, rel)) == d
             for rel, d in m.items())


if __name__ == "__main__":
  ap = argparse.ArgumentParser()
  ap.add_argument("--reference", default="/root/reference")
  ap.add_argument("--check", action="store_true")
  a = ap.parse_args()
  if a.check:
    sys.exit(0 if check() else 1)
  m = make(a.reference)
  print(f"{len(m)} files -> {OUT}")
