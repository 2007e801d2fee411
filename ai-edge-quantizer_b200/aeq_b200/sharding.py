"""Multi-GPU sharding of the requantisation path: one process per GPU, tensors as units.

Every weight tensor is an independent unit of work (the reference walks them one op
at a time with no cross-tensor state, params_generator.py:110-183), so the path
shards with NO data-path collective: `assign_tensors` bin-packs tensors onto ranks by
bytes (longest-processing-time first), each rank requantises the tensors it owns, and
quantised payloads stay with their owner.  The one real exchange is small: every rank
needs every tensor's scale vector (4 B per row per-channel, 2-4 B per block) to write
the model's quantisation parameters, so `allgather_vectors` does ONE all-gather of a
flat, offset-indexed buffer (NCCL for CUDA tensors, gloo for CPU tensors).

Works unchanged at world size 1 (no process group needed).
"""
from __future__ import annotations

import heapq
from typing import Callable, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist


def assign_tensors(sizes: Sequence[int], world: int) -> list[int]:
  """Rank that owns each tensor: LPT bin packing by size, ties broken by index (deterministic,
  so every rank computes the same table without communicating)."""
  if world < 1:
    raise ValueError("world size must be >= 1")
  order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
  heap = [(0, r) for r in range(world)]
  heapq.heapify(heap)
  owner = [0] * len(sizes)
  for i in order:
    load, r = heapq.heappop(heap)
    owner[i] = r
    heapq.heappush(heap, (load + int(sizes[i]), r))
  return owner


def owned(owner: Sequence[int], rank: int) -> list[int]:
  return [i for i, r in enumerate(owner) if r == rank]


def imbalance(sizes: Sequence[int], owner: Sequence[int], world: int) -> float:
  """max rank load / mean rank load (1.0 = perfect)."""
  loads = [0] * world
  for s, r in zip(sizes, owner):
    loads[r] += int(s)
  mean = sum(loads) / world
  return max(loads) / mean if mean else 1.0


def _world(group) -> tuple[int, int]:
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(group), dist.get_world_size(group)
  return 0, 1


def _vote(err, group, device) -> None:
  """All ranks raise together or none does: a one-sided raise before a collective would leave
  the other ranks waiting in it (same rule as peer.PeerScales._vote)."""
  rank, world = _world(group)
  if world > 1:
    ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0 and err is None:
      err = RuntimeError("allgather_vectors: another rank reported an error; no data was exchanged")
  if err is not None:
    raise err


def allgather_vectors(local: dict[int, torch.Tensor], lengths: Sequence[int], owner: Sequence[int],
                      group=None, dtype: Optional[torch.dtype] = None,
                      device: Optional[torch.device] = None) -> list[torch.Tensor]:
  """All ranks end up with every tensor's vector.

  local:   {tensor index: 1-D tensor} for the tensors this rank owns (all one dtype / device).
  lengths: vector length of EVERY tensor (known everywhere: it follows from the shapes).
  dtype / device: needed only by a rank that owns no tensor (fewer tensors than ranks): it then
           contributes a zero-filled row instead of guessing.
  One all_gather_into_tensor of a [world, max_rank_len] buffer; rank r's slots are its
  tensors in index order, so offsets need no exchange.  Argument errors are voted on first, so
  that every rank raises or none does.
  """
  rank, world = _world(group)
  mine = owned(owner, rank)
  per_rank = [sum(int(lengths[i]) for i in owned(owner, r)) for r in range(world)]
  width = max(per_rank) if per_rank else 0
  proto = next(iter(local.values())) if local else None
  dtype = proto.dtype if proto is not None else dtype
  device = proto.device if proto is not None else device
  err = None
  if sorted(local) != mine:
    err = ValueError(f"rank {rank} must supply exactly the tensors it owns: {mine}")
  elif world > 1 and width > 0 and (dtype is None or device is None):
    err = ValueError("a rank without tensors must be given dtype= and device= to join the all-gather")
  else:
    for i in mine:
      if local[i].numel() != int(lengths[i]):
        err = ValueError(f"tensor {i}: expected {lengths[i]} values, got {local[i].numel()}")
        break
  if device is None:  # nothing to infer from: the vote itself runs on the backend's default
    device = torch.device("cuda", torch.cuda.current_device()) if (
        world > 1 and dist.get_backend(group) == "nccl") else torch.device("cpu")
  _vote(err, group, device)
  if width == 0:
    return [torch.empty(0, dtype=dtype or torch.float32) for _ in lengths]
  flat = torch.zeros(width, dtype=dtype, device=device)
  off = 0
  for i in mine:
    v = local[i].reshape(-1)
    flat[off:off + v.numel()] = v
    off += v.numel()
  if world == 1:
    gathered = flat.reshape(1, width)
  else:
    buf = torch.empty(world * width, dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(buf, flat, group=group)
    gathered = buf.reshape(world, width)
  out: list[Optional[torch.Tensor]] = [None] * len(lengths)
  for r in range(world):
    off = 0
    for i in owned(owner, r):
      n = int(lengths[i])
      out[i] = gathered[r, off:off + n]
      off += n
  return out  # type: ignore[return-value]


def requantize_sharded(weights: Sequence[np.ndarray], compute: Callable[[list[np.ndarray]], list],
                       scale_length: Callable[[np.ndarray], int], group=None,
                       device: Optional[torch.device] = None):
  """Shards `weights` over the process group, runs `compute` on the owned ones, all-gathers scales.

  weights:      the model's weight arrays (every rank sees the same list; only owned ones are read).
  compute:      owned arrays -> list of (q, packed, scale, extra) tuples, e.g.
                `functools.partial(aeq_b200.host.requant_rows, bits=8)`.
  scale_length: array -> number of scale entries (rows, or rows * cols / block).
  Returns (owner table, {index: result tuple} for owned tensors, [scale vector of every tensor]).
  """
  rank, world = _world(group)
  sizes = [int(w.size) * w.dtype.itemsize for w in weights]
  owner = assign_tensors(sizes, world)
  mine = owned(owner, rank)
  results = compute([weights[i] for i in mine]) if mine else []
  local = {}
  for i, res in zip(mine, results):
    s = torch.from_numpy(np.ascontiguousarray(res[2]).reshape(-1))
    local[i] = s.to(device) if device is not None else s
  scales = allgather_vectors(local, [scale_length(w) for w in weights], owner, group,
                             dtype=torch.float32, device=device or torch.device("cpu"))
  return owner, dict(zip(mine, results)), scales


# ------------------------------------------------------------------ row-split units (SURVEY.md §8e "Partitioning")
def split_units(shapes: Sequence[Sequence[int]], world: int, itemsize: int = 4,
                threshold: float = 0.5, row_align: int = 1) -> list[tuple[int, int, int]]:
  """Units of work [(tensor index, first row, end row)] in tensor order.

  A tensor is one unit unless its bytes exceed `threshold` x (total bytes / world) — the balance
  threshold: one such tensor alone would tip a rank over its fair share.  It is then cut BY ROWS
  into the fewest near-equal pieces that fit under the threshold.  Rows are the separable axis of
  the path: a per-channel row and every 32..256-block lie inside one row of the [rows, cols] view
  (common_quantize.py:1337-1344, uqt:246-262), so a piece is requantised exactly as the whole
  tensor would be.  Only TENSORWISE granularity couples the pieces, through one (min, max) pair:
  `allreduce_minmax`.  `row_align` keeps piece boundaries on a multiple of rows (e.g. 2 when INT4
  rows of odd length are packed across the row boundary).  Deterministic: every rank computes the
  same list without communicating.
  """
  if world < 1:
    raise ValueError("world size must be >= 1")
  sizes = [int(np.prod(s)) * itemsize for s in shapes]
  fair = sum(sizes) / world if sizes else 0.0
  limit = threshold * fair
  units = []
  for i, (shape, nbytes) in enumerate(zip(shapes, sizes)):
    rows = int(shape[0]) if len(shape) else 1
    pieces = 1
    if world > 1 and limit > 0 and nbytes > limit:
      pieces = min(int(-(-nbytes // limit)), max(rows // max(row_align, 1), 1))
    if pieces <= 1:
      units.append((i, 0, rows))
      continue
    per = -(-rows // pieces)
    per = -(-per // row_align) * row_align
    r0 = 0
    while r0 < rows:
      units.append((i, r0, min(r0 + per, rows)))
      r0 += per
  return units


def unit_sizes(units: Sequence[tuple[int, int, int]], shapes: Sequence[Sequence[int]],
               itemsize: int = 4) -> list[int]:
  return [(r1 - r0) * int(np.prod(shapes[i][1:])) * itemsize for i, r0, r1 in units]


def allreduce_minmax(local: torch.Tensor, group=None) -> torch.Tensor:
  """Per-tensor (min, max) pairs combined over the ranks that hold rows of the tensor: ONE
  all-reduce(MAX) of [-min, max] (2 floats per split tensor).  A rank holding no row of a tensor
  passes (+inf, -inf).  NaNs follow the local reduction's rule before they get here."""
  rank, world = _world(group)
  if world == 1:
    return local
  packed = torch.stack([-local[:, 0], local[:, 1]], dim=1).contiguous()
  dist.all_reduce(packed, op=dist.ReduceOp.MAX, group=group)
  return torch.stack([-packed[:, 0], packed[:, 1]], dim=1)


def requantize_row_sharded(weights: Sequence[np.ndarray], compute: Callable[[list[np.ndarray]], list],
                           scale_length: Callable[[np.ndarray], int], group=None,
                           device: Optional[torch.device] = None, threshold: float = 0.5,
                           row_align: int = 1):
  """`requantize_sharded` with tensors above the balance threshold cut by rows.

  compute / scale_length see row slices `w[r0:r1]` (views, no copy) of the split tensors.
  Returns (units, owner of each unit, {unit index: result tuple} for owned units,
  [scale vector of every TENSOR], pieces concatenated in row order).
  """
  rank, world = _world(group)
  shapes = [w.shape for w in weights]
  units = split_units(shapes, world, weights[0].dtype.itemsize if len(weights) else 4, threshold, row_align)
  owner = assign_tensors(unit_sizes(units, shapes, weights[0].dtype.itemsize if len(weights) else 4), world)
  mine = owned(owner, rank)
  views = [weights[i][r0:r1] for i, r0, r1 in (units[u] for u in mine)]
  results = compute(views) if mine else []
  local = {}
  for u, res in zip(mine, results):
    s = torch.from_numpy(np.ascontiguousarray(res[2]).reshape(-1))
    local[u] = s.to(device) if device is not None else s
  lengths = [scale_length(weights[i][r0:r1]) for i, r0, r1 in units]
  per_unit = allgather_vectors(local, lengths, owner, group, dtype=torch.float32,
                               device=device or torch.device("cpu"))
  scales: list[list[torch.Tensor]] = [[] for _ in weights]
  for (i, _, _), v in zip(units, per_unit):
    scales[i].append(v)
  return units, owner, dict(zip(mine, results)), [torch.cat(p) if len(p) > 1 else p[0] for p in scales]
