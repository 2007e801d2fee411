"""Import shim that makes /root/reference's numeric modules importable here.

TEST INFRASTRUCTURE ONLY. It exists so that (a) tests/golden/ fixtures can be
generated from the *unmodified* reference and (b) CPU-side tests in the build
container can cross-check `oracle/` against the reference when
`/root/reference` is mounted. The GPU box has no /root/reference: there the
shim resolves to `oracle/_ref/` (the reference's own sources copied byte for
byte by the committed `oracle/make_ref.py`; git-ignored, shipped with the
snapshot like the built `.so`), so `-m gpu` tests and `bench.py --impl
reference` can run the UNMODIFIED reference beside the device path. Nothing
under `aeq_b200/` (the product) imports this.

Recipe: put stub packages for the absent wheels
(`immutabledict`, `ai_edge_litert.tools.*`) on sys.path and pre-register an
empty `ai_edge_quantizer` package object whose `__path__` points at the
reference tree, so `ai_edge_quantizer/__init__.py:19` (which drags in the
LiteRT interpreter and `flatbuffers`) never runs.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")
_TRAVELLING = os.path.normpath(os.path.join(_HERE, "..", "_ref"))


def _find_root() -> str:
  """/root/reference (build container) first, then the travelling copy oracle/_ref."""
  env = os.environ.get("AEQ_REFERENCE_ROOT")
  for cand in ([env] if env else []) + ["/root/reference", _TRAVELLING]:
    if os.path.isdir(os.path.join(cand, "ai_edge_quantizer")):
      return cand
  return env or "/root/reference"


REFERENCE_ROOT = _find_root()


def available() -> bool:
  return os.path.isdir(os.path.join(REFERENCE_ROOT, "ai_edge_quantizer"))


def kind() -> str:
  """"reference" (the mounted tree) or "_ref" (oracle/_ref, the travelling byte-for-byte copy)."""
  return "_ref" if os.path.normpath(REFERENCE_ROOT) == _TRAVELLING else "reference"


def install() -> None:
  """Idempotently wires the stubs and the package bypass."""
  if not available():
    raise RuntimeError(f"reference tree not mounted at {REFERENCE_ROOT}")
  if _STUBS not in sys.path:
    sys.path.insert(0, _STUBS)
  if "ai_edge_quantizer" not in sys.modules:
    pkg = types.ModuleType("ai_edge_quantizer")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "ai_edge_quantizer")]
    sys.modules["ai_edge_quantizer"] = pkg
  try:
    from absl import flags
    if not flags.FLAGS.is_parsed():
      flags.FLAGS.mark_as_parsed()
  except Exception:  # absl is optional for the numeric modules
    pass


def ref(module: str):
  """Imports `ai_edge_quantizer.<module>` from the reference tree."""
  install()
  return importlib.import_module("ai_edge_quantizer." + module)


def fc_op_info(weight_cfg, op_name="FULLY_CONNECTED", **op_cfg_kw):
  """Synthetic OpInfo for a constant FC/EMBEDDING weight (no .tflite)."""
  q = ref("qtyping")
  op = types.SimpleNamespace(inputs=[0, 1, -1], outputs=[2])
  return q.OpInfo(
      op,
      q.TFLOperationName(op_name),
      0,
      q.OpQuantizationConfig(weight_tensor_config=weight_cfg, **op_cfg_kw),
  )
