"""Gathered per-channel scales over NVLink peer memory (one process per GPU, one node).

The sharded requantisation path has one exchange: every rank needs every tensor's scale vector
(sharding.py).  With NCCL that is one all-gather launch per step, whose latency is what the
multi-GPU bench loses against linear scaling.  Here every rank owns a `[world, slots]` fp32
buffer that all the other ranks map (CUDA IPC, `aeqb_peer_alloc` / `aeqb_peer_open`); a rank's
scale outputs are views into ITS row of ITS buffer, and the requantisation kernel stores each row's
scale into the same slot of every peer's buffer from its own epilogue
(`aeqb_requant_rows_batch_mirror_f32`).  Nothing else is launched; remote rows are complete after
`sync()` (stream synchronise + barrier).

Setup needs `torch.distributed` only to pass the 64-byte handles around.  There is no fallback in
here: if peer mapping fails the constructor raises and the caller decides (bench.py falls back to
the NCCL all-gather and says so in its JSON line).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from aeq_b200 import _lib


class _DevMem:
  """__cuda_array_interface__ holder so torch can view a raw device allocation."""

  def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
    self.__cuda_array_interface__ = {
        "shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3, "strides": None}


_TYPESTR = {torch.float32: ("<f4", 4), torch.float16: ("<f2", 2)}


class PeerScales:
  """`[world, slots]` fp32 on every rank; row r is written by rank r's kernels on all ranks."""

  def __init__(self, slots: int, device: torch.device, group=None, dtype: torch.dtype = torch.float32):
    """dtype: torch.float32 (per-channel scales) or torch.float16 (blockwise scale tensors)."""
    if not (dist.is_available() and dist.is_initialized()):
      raise RuntimeError("PeerScales needs an initialised process group")
    self.group = group
    self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
    if self.world - 1 > 15:
      raise ValueError("at most 16 ranks (kMaxPeers = 15)")
    typestr, itemsize = _TYPESTR[dtype]
    self.slots = (int(slots) * itemsize + 15) // 16 * 16 // itemsize  # every rank's row starts 16-byte aligned
    self.device = device
    nbytes = self.world * self.slots * itemsize
    # Every step below ends in a vote, so that a rank whose allocation or mapping failed does not
    # leave the others waiting in a collective: either all ranks get a PeerScales or all raise.
    self._base = None
    self._peer_ptrs = []
    handle = (ctypes.c_ubyte * 64)()
    err = None
    try:
      ptr = ctypes.c_void_p()
      _lib.call("aeqb_peer_alloc", nbytes, ctypes.byref(ptr), ctypes.cast(handle, ctypes.c_void_p))
      self._base = int(ptr.value)
    except Exception as e:  # pylint: disable=broad-except
      err = e
    self._vote(err, "allocating the peer-visible buffer")
    mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=device)
    every = torch.empty(self.world * 64, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(every, mine, group=group)
    every = every.cpu().numpy().reshape(self.world, 64)
    deltas = []
    try:
      for r in range(self.world):
        if r == self.rank:
          continue
        h = (ctypes.c_ubyte * 64)(*every[r].tolist())
        p = ctypes.c_void_p()
        _lib.call("aeqb_peer_open", ctypes.cast(h, ctypes.c_void_p), ctypes.byref(p))
        self._peer_ptrs.append(int(p.value))
        deltas.append(int(p.value) - self._base)
    except Exception as e:  # pylint: disable=broad-except
      err = e
    self._vote(err, "mapping the peers' buffers")
    self.n_peers = len(deltas)
    self._deltas = (ctypes.c_int64 * max(1, self.n_peers))(*deltas)
    self.gathered = torch.as_tensor(_DevMem(self._base, self.world * self.slots, typestr),
                                    device=device).view(self.world, self.slots)
    self.local = self.gathered[self.rank]  # this rank's scale outputs are views into this row
    self.sync()

  def _vote(self, err, what: str):
    ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=self.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
    if int(ok.item()) == 1:
      return
    for p in self._peer_ptrs:
      try:
        _lib.call("aeqb_peer_close", p)
      except Exception:  # pylint: disable=broad-except
        pass
    self._peer_ptrs = []
    if self._base is not None:
      self._device_sync()
      dist.barrier(group=self.group)
      _lib.call("aeqb_peer_free", self._base)
      self._base = None
    else:
      dist.barrier(group=self.group)
    raise RuntimeError(f"PeerScales: {what} failed on at least one rank"
                       + (f" (here: {err})" if err else ""))

  @property
  def deltas_ptr(self):
    return ctypes.cast(self._deltas, ctypes.c_void_p)

  def sync(self):
    """After this, every rank's `gathered` holds all rows written before the call."""
    self._device_sync()
    dist.barrier(group=self.group)

  def _device_sync(self):
    if self.device.type == "cuda":
      torch.cuda.synchronize(self.device)

  def close(self):
    if self._base is None:
      return
    self.sync()
    for p in self._peer_ptrs:
      _lib.call("aeqb_peer_close", p)
    self.sync()  # nobody maps the buffer any more
    self.gathered = self.local = None
    _lib.call("aeqb_peer_free", self._base)
    self._base = None
