"""hessian_inverse at the Llama-7B orders: time and residual of every Cholesky variant.
  python tools/hinv_time.py [K ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")]
import torch  # noqa: E402

from aeq_b200 import device  # noqa: E402

dev = torch.device("cuda:0")
VARIANTS = {
    "dmma+lookahead": {"AEQB_CHOL_DMMA_MIN_K": "0"},
    "dmma serial": {"AEQB_CHOL_DMMA_MIN_K": "0", "AEQB_CHOL_NO_LOOKAHEAD": "1"},
    "two-level simt": {"AEQB_CHOL_DMMA_MIN_K": "1000000", "AEQB_CHOL_TWO_LEVEL_MIN_K": "0"},
}
only = os.environ.get("HINV_VARIANTS")
for k in [int(a) for a in sys.argv[1:]] or [4096, 11008]:
  x = torch.randn(16384, k, device=dev)
  h = device.xtx(x, 2.0 / 8)
  del x
  for name, env in VARIANTS.items():
    if only and name not in only.split(","):
      continue
    for key in ("AEQB_CHOL_DMMA_MIN_K", "AEQB_CHOL_NO_LOOKAHEAD", "AEQB_CHOL_TWO_LEVEL_MIN_K"):
      os.environ.pop(key, None)
    os.environ.update(env)
    for _ in range(2):
      device.hessian_inverse(h, 0.01)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hd = h.clone()
    hinv = device.hessian_inverse(hd, 0.01, keep_damped_diagonal=True)
    e1.record()
    torch.cuda.synchronize()
    resid = hd @ hinv.double()
    resid.diagonal().sub_(1.0)
    print(f"K={k:6d} {name:16s} {e0.elapsed_time(e1):8.2f} ms   max|H Hinv - I| = {float(resid.abs().max()):.2e}", flush=True)
    del hd, hinv, resid
