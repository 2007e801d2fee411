"""Single-GPU replay of the N = 2 llama7b_gptq bench step (both ranks' token shards generated here, the partial
Hessians added in float64 as the NCCL reduce does), to tell a numeric failure from a timing one.
  python tools/gptq_n2_repro.py [world=2] [repeats=3]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ai-edge-quantizer_b200")):
  sys.path.insert(0, p)
import torch  # noqa: E402
from aeq_b200 import device  # noqa: E402
from aeq_b200.algorithms.uniform_quantize import hadamard_gptq  # noqa: E402

LLAMA_LAYER = [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)]
LLAMA_FEEDS = ["attn_in"] * 3 + ["attn_out"] + ["mlp_in"] * 2 + ["mlp_mid"]
LLAMA_INPUTS = [("attn_in", 4096), ("attn_out", 4096), ("mlp_in", 4096), ("mlp_mid", 11008)]
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
repeats = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
t_local = 262144 // world
for owner in range(world):
  hs = {}
  for rank in range(world):  # rank's shard of layer `owner`'s inputs (bench.py generates every layer's in order)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    for l in range(world):
      for name, k in LLAMA_INPUTS:
        x = torch.randn(t_local, k, device=dev, generator=gen)
        if l == owner:
          h = device.xtx(x, 2.0 / 128)
          hs[name] = h if name not in hs else hs[name] + h
        del x
  gen = torch.Generator(device=dev).manual_seed(777 + owner)
  layer = [torch.randn(r, c, device=dev, generator=gen) * 0.02 for r, c in LLAMA_LAYER]
  for name, h in hs.items():
    d = torch.diagonal(h)
    print(f"owner {owner} {name}: dtype {h.dtype} diag min {float(d.min()):.4g} max {float(d.max()):.4g} "
          f"finite {bool(torch.isfinite(h).all())} sym_err {float((h - h.T).abs().max()):.3g}")
  for rep in range(repeats):
    for conc in (True, False):
      try:
        hadamard_gptq.quantize_layer_device(layer, LLAMA_FEEDS, hs, 4, True, 4096, 0.01, concurrent=conc)
        torch.cuda.synchronize()
        print(f"owner {owner} rep {rep} concurrent={conc}: ok")
      except Exception as e:  # noqa: BLE001
        print(f"owner {owner} rep {rep} concurrent={conc}: {type(e).__name__}: {e}")
  # each Hessian on its own, every Cholesky variant the environment selects
  for name, h in hs.items():
    try:
      device.hessian_inverse(h, 0.01)
      print(f"owner {owner} {name}: single inverse ok")
    except Exception as e:  # noqa: BLE001
      print(f"owner {owner} {name}: single inverse {type(e).__name__}: {e}")
