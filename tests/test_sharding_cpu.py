"""Host-side logic of the multi-GPU path on CPU: LPT assignment and the one all-gather of scale
vectors, run at world size 2 over gloo (the N > 1 plumbing; the arithmetic itself is GPU-only,
so the per-rank `compute` here is the oracle standing in as the checker's reference)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aeq_b200 import sharding
from oracle import aeq_oracle as O


def test_assignment_is_balanced_and_deterministic():
  sizes = [4096 * 4096 * 4] * 7 + [16384 * 2048 * 4] * 4 + [256 * 2048 * 4] * 9
  for world in (1, 2, 4, 8):
    owner = sharding.assign_tensors(sizes, world)
    assert owner == sharding.assign_tensors(sizes, world)
    assert set(owner) <= set(range(world))
    assert sorted(sum((sharding.owned(owner, r) for r in range(world)), [])) == list(range(len(sizes)))
    assert sharding.imbalance(sizes, owner, world) < 1.15
  # Gemma-2B-shaped layer set x 18 layers over 8 GPUs (BASELINE configs[2])
  layer = [2048 * 2048] * 2 + [256 * 2048] * 2 + [16384 * 2048] * 2 + [2048 * 16384]
  sizes = [4 * s for s in layer] * 18
  assert sharding.imbalance(sizes, sharding.assign_tensors(sizes, 8), 8) < 1.02
  with pytest.raises(ValueError):
    sharding.assign_tensors([1], 0)


def test_single_process_path_needs_no_group():
  ws = [O.synthetic_weight(8, 64, i) for i in range(3)]
  compute = lambda arrs: [(r["q"], None, r["scale"], r["zero_point"])
                          for r in (O.minmax_requant(a, 8, True) for a in arrs)]
  owner, mine, scales = sharding.requantize_sharded(ws, compute, lambda w: w.shape[0])
  assert owner == [0, 0, 0] and sorted(mine) == [0, 1, 2]
  for w, s in zip(ws, scales):
    np.testing.assert_array_equal(s.numpy(), O.minmax_requant(w, 8, True)["scale"].reshape(-1))


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, shapes, block, out_dir):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    ws = [O.synthetic_weight(r, c, i) for i, (r, c) in enumerate(shapes)]
    calls = []

    def compute(arrs):
      calls.append(len(arrs))
      out = []
      for a in arrs:
        r = O.minmax_requant(a, 4 if block else 8, True, block=block)
        out.append((r["q"], None, r["scale"], None))
      return out

    length = (lambda w: w.size // block) if block else (lambda w: w.shape[0])
    owner, mine, scales = sharding.requantize_sharded(ws, compute, length)
    assert sorted(mine) == sharding.owned(owner, rank)
    assert calls == [len(mine)]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), owner=np.array(owner),
             **{f"s{i}": s.numpy() for i, s in enumerate(scales)})
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("block", [0, 32])
def test_world2_gloo_allgather_of_scales(tmp_path, block):
  """Ragged tensor set over 2 ranks: both ranks end with every tensor's scale vector."""
  shapes = [(64, 256), (8, 1024), (130, 64), (16, 512), (3, 2048)]
  port = _free_port()
  mp.start_processes(_worker, args=(2, port, shapes, block, str(tmp_path)), nprocs=2, join=True,
                     start_method="spawn")
  got = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
  assert list(got[0]["owner"]) == list(got[1]["owner"])
  assert set(got[0]["owner"]) == {0, 1}
  for i, (r, c) in enumerate(shapes):
    want = O.minmax_requant(O.synthetic_weight(r, c, i), 4 if block else 8, True, block=block)["scale"]
    for g in got:
      np.testing.assert_array_equal(g[f"s{i}"], want.reshape(-1))


def test_allgather_validates_ownership():
  with pytest.raises(ValueError, match="must supply exactly the tensors it owns"):
    sharding.allgather_vectors({1: torch.zeros(3)}, [3, 3], [0, 0])
  with pytest.raises(ValueError, match="expected 3 values"):
    sharding.allgather_vectors({0: torch.zeros(2)}, [3], [0])


def _peer_worker(rank, world, port, out_dir, fake_success_on):
  """PeerScales set-up on a box without CUDA: the allocation fails (really, or not at all on the
  rank that fakes it) and every rank must leave the constructor with the same RuntimeError."""
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from aeq_b200 import _lib, peer
    calls = []
    if rank in fake_success_on:
      real = _lib.call

      def fake(name, *args):
        calls.append(name)
        if name == "aeqb_peer_alloc":
          args[1]._obj.value = 0x10000  # ctypes.byref(ptr)
          return None
        if name == "aeqb_peer_free":
          return None
        return real(name, *args)
      peer._lib.call = fake
    try:
      peer.PeerScales(16, torch.device("cpu"))
      outcome = "constructed"
    except RuntimeError as e:
      outcome = str(e)
    with open(os.path.join(out_dir, f"peer{rank}.txt"), "w") as f:
      f.write(outcome + "\n" + ",".join(calls))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize("fake_success_on", [(), (1,)])
def test_world2_peer_setup_fails_together(tmp_path, fake_success_on):
  """A rank whose peer-visible allocation fails must not leave the others in a collective: the
  vote makes every rank raise (and the rank that did allocate frees its buffer)."""
  if torch.cuda.is_available():
    pytest.skip("the failure path needs a box without CUDA")
  port = _free_port()
  mp.start_processes(_peer_worker, args=(2, port, str(tmp_path), fake_success_on), nprocs=2,
                     join=True, start_method="spawn")
  for r in range(2):
    outcome, calls = (tmp_path / f"peer{r}.txt").read_text().split("\n")
    assert "allocating the peer-visible buffer failed on at least one rank" in outcome
    if r in fake_success_on:
      assert calls.split(",") == ["aeqb_peer_alloc", "aeqb_peer_free"]


def test_split_units_cut_heavy_tensors_by_rows():
  """One [16384, 2048] among small tensors: whole-tensor LPT cannot balance 2 ranks, row pieces can."""
  shapes = [(16384, 2048)] + [(256, 2048)] * 6
  sizes = [r * c * 4 for r, c in shapes]
  assert sharding.imbalance(sizes, sharding.assign_tensors(sizes, 2), 2) > 1.6
  units = sharding.split_units(shapes, 2)
  assert units == sharding.split_units(shapes, 2)
  us = sharding.unit_sizes(units, shapes)
  assert sum(us) == sum(sizes)
  assert sharding.imbalance(us, sharding.assign_tensors(us, 2), 2) < 1.1
  rows = sorted((r0, r1) for i, r0, r1 in units if i == 0)
  assert rows[0][0] == 0 and rows[-1][1] == 16384 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
  assert all(u in units for u in [(k, 0, 256) for k in range(1, 7)])
  assert sharding.split_units(shapes, 1) == [(i, 0, s[0]) for i, s in enumerate(shapes)]
  assert all((r1 - r0) % 32 == 0 or r1 == 1000 for _, r0, r1 in sharding.split_units([(1000, 64)], 4, row_align=32))
  with pytest.raises(ValueError):
    sharding.split_units(shapes, 0)


def _split_worker(rank, world, port, out_dir):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    shapes = [(512, 128), (16, 64), (8, 256)]
    ws = [O.synthetic_weight(r, c, i) for i, (r, c) in enumerate(shapes)]
    compute = lambda arrs: [(r["q"], None, r["scale"], None) for r in (O.minmax_requant(a, 8, True) for a in arrs)]
    units, owner, mine, scales = sharding.requantize_row_sharded(ws, compute, lambda w: w.shape[0])
    held = [units[u] for u in mine]
    # per-tensor granularity: each rank reduces the rows it holds, one all-reduce of 2 floats joins them
    mm = torch.tensor([[np.inf, -np.inf]] * len(ws), dtype=torch.float32)
    for i, r0, r1 in held:
      mm[i, 0] = min(float(mm[i, 0]), float(ws[i][r0:r1].min()))
      mm[i, 1] = max(float(mm[i, 1]), float(ws[i][r0:r1].max()))
    mm = sharding.allreduce_minmax(mm)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), units=np.array(units), owner=np.array(owner),
             mm=mm.numpy(), held=np.array(held), **{f"s{i}": s.numpy() for i, s in enumerate(scales)})
  finally:
    dist.destroy_process_group()


def test_world2_gloo_row_split_and_minmax_allreduce(tmp_path):
  """The heavy tensor is cut by rows over both ranks; every rank ends with the whole scale vector
  (bit-exact vs the unsplit oracle) and the tensor-wise (min, max) of the split tensor."""
  port = _free_port()
  mp.start_processes(_split_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
  got = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
  assert np.array_equal(got[0]["units"], got[1]["units"]) and np.array_equal(got[0]["owner"], got[1]["owner"])
  assert {tuple(h)[0] for h in got[0]["held"]} & {tuple(h)[0] for h in got[1]["held"]} == {0}  # both hold rows of tensor 0
  for i, (r, c) in enumerate([(512, 128), (16, 64), (8, 256)]):
    w = O.synthetic_weight(r, c, i)
    want = O.minmax_requant(w, 8, True)["scale"].reshape(-1)
    for g in got:
      np.testing.assert_array_equal(g[f"s{i}"], want)
      assert g["mm"][i, 0] == w.min() and g["mm"][i, 1] == w.max()
