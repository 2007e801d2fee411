"""The C-ABI shared library loads and exports exactly what include/aeqb200.h declares.
No compute calls (CPU box)."""
import ctypes
import os
import re
import subprocess

import pytest

from aeq_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
  if not os.path.exists(_lib.LIB_PATH):
    import __graft_entry__
    __graft_entry__.build()
  return _lib.load()


def test_header_and_binding_table_agree(lib):
  declared = set(_lib.header_functions())
  assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
  assert len(declared) >= 13


def test_every_declared_symbol_is_exported(lib):
  for name in _lib.header_functions():
    assert hasattr(lib, name), name


def test_only_c_abi_is_exported():
  out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True,
                       text=True, check=True).stdout
  exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
  ours = {s for s in exported if s.startswith("aeqb_")}
  assert ours == set(_lib.header_functions())
  assert not [s for s in exported if s.startswith("_ZN4aeqb")], "C++ internals leaked"


def test_version_and_error_string(lib):
  assert lib.aeqb_version() == 100
  assert isinstance(lib.aeqb_last_error(), bytes)
  assert lib.aeqb_minmax_workspace_bytes() >= 20


def test_argument_validation_needs_no_gpu(lib):
  """Bad arguments are rejected before any CUDA call, with the reference's wording."""
  rc = lib.aeqb_requant_blocks_f32(None, 4, 48, 32, 4, None, None, None, None, None, None)
  assert rc != 0
  assert b"is not divisible by block size 32" in lib.aeqb_last_error()
  rc = lib.aeqb_requant_rows_f32(None, 4, 8, 5, 1, None, None, None, None, None, None)
  assert rc != 0 and b"num_bits" in lib.aeqb_last_error()
  rc = lib.aeqb_pack_bits(None, 8, 3, None, None)
  assert rc != 0


def test_signatures_have_no_torch_types():
  text = open(_lib.HEADER_PATH).read()
  assert "torch" not in re.sub(r"/\*.*?\*/", "", text, flags=re.S).lower()
  assert 'extern "C"' in text


def test_sass_contains_tma_bulk_copies():
  """The tile-stream kernels really use the TMA bulk-copy engine (UBLKCP) + mbarriers."""
  cuobjdump = "/usr/local/cuda/bin/cuobjdump"
  if not os.path.exists(cuobjdump):
    pytest.skip("cuobjdump not available")
  sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
  assert "UBLKCP" in sass and "SYNCS" in sass
  assert "sm_100a" in sass
