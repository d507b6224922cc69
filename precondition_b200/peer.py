"""All-gather of the block-sharded roots over NVLink peer memory (``pc_peer_all_gather``).

The exchange of DS:2876-2877 for one process per GPU on an NVSwitch box: every rank pushes its
payload into the peers' receive buffers with copy engines (CUDA IPC mappings) and signals with
4-byte epoch flags; no SM is involved, so it overlaps the persistent GEMM kernels that own the
whole GPU.  ``torch.distributed`` is only the launcher-side plumbing that carries the IPC
handles once; if peer mappings are impossible (different nodes, no P2P, an allocator without
IPC support) every rank falls back to ``torch.distributed.all_gather_into_tensor`` (NCCL)."""
from __future__ import annotations

import ctypes
import os

import torch

from precondition_b200 import _lib


class _NcclGather:
  """Same interface over NCCL."""
  kind = "nccl"

  def __init__(self, nbytes, world, group, device):
    self.group = group
    self.recv = torch.empty((world, nbytes), dtype=torch.uint8, device=device)

  def all_gather(self, send):
    import torch.distributed as dist
    dist.all_gather_into_tensor(self.recv.view(-1), send.view(torch.uint8).view(-1),
                                group=self.group)
    return self.recv

  def release(self):
    pass


class _PeerGather:
  kind = "peer_copy_engine"

  def __init__(self, nbytes, world, rank, device, recv, flags, peer_recv, peer_flags):
    self.recv, self.flags = recv, flags
    self.nbytes, self.epoch = nbytes, 0
    g = _lib.PeerGroup()
    g.world, g.rank, g.slot_bytes = world, rank, nbytes
    for p in range(world):
      g.recv[p] = recv.data_ptr() if p == rank else peer_recv[p]
      g.flags[p] = flags.data_ptr() if p == rank else peer_flags[p]
    self.group = g
    self.device = device

  def all_gather(self, send):
    assert send.is_cuda and send.is_contiguous() and send.numel() * send.element_size() <= self.nbytes
    self.epoch += 1
    lib = _lib.load()
    with torch.cuda.device(self.device):
      _lib.check(lib.pc_peer_all_gather(
          ctypes.byref(self.group), ctypes.c_void_p(send.data_ptr()),
          send.numel() * send.element_size(), self.epoch,
          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return self.recv

  def release(self):
    lib = _lib.load()
    with torch.cuda.device(self.device):
      _lib.check(lib.pc_peer_release(ctypes.byref(self.group), self.epoch,
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))


def make_all_gather(nbytes: int, group=None, device=None):
  """Collective constructor (call on every rank of ``group`` in the same order): a gather
  object with ``recv`` ([world, nbytes] uint8), ``all_gather(send)`` (enqueue on the current
  stream; returns ``recv``) and ``release()`` (enqueue after the last reader of ``recv``)."""
  import torch.distributed as dist
  world, rank = dist.get_world_size(group), dist.get_rank(group)
  device = device or torch.device("cuda", torch.cuda.current_device())
  if os.environ.get("PC_GATHER", "").lower() == "nccl" or dist.get_backend(group) != "nccl":
    return _NcclGather(nbytes, world, group, device)
  lib = _lib.load()
  recv = torch.empty((world, nbytes), dtype=torch.uint8, device=device)
  flags = torch.zeros(_lib.PC_PEER_FLAG_WORDS, dtype=torch.int32, device=device)
  ok, mine = 1, None
  try:
    hs = []
    for t in (recv, flags):
      h = _lib.IpcHandle()
      _lib.check(lib.pc_ipc_export(ctypes.c_void_p(t.data_ptr()), ctypes.byref(h)))
      hs.append(bytes(h))
    mine = (hs[0], hs[1], os.uname().nodename)
  except Exception:  # pylint: disable=broad-except
    ok = 0
  everyone = [None] * world
  dist.all_gather_object(everyone, mine, group=group)
  peer_recv, peer_flags = {}, {}
  if ok and all(e is not None and e[2] == mine[2] for e in everyone) and world <= _lib.PC_MAX_PEERS:
    try:
      for p, e in enumerate(everyone):
        if p == rank:
          continue
        for blob, table in ((e[0], peer_recv), (e[1], peer_flags)):
          h = _lib.IpcHandle.from_buffer_copy(blob)
          out = ctypes.c_void_p()
          _lib.check(lib.pc_ipc_open(ctypes.byref(h), ctypes.byref(out)))
          table[p] = out.value
    except Exception:  # pylint: disable=broad-except
      ok = 0
  else:
    ok = 0
  agree = torch.tensor([ok], dtype=torch.int32, device=device)
  dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)
  torch.cuda.synchronize(device)  # the flags are zero everywhere before anyone pushes
  dist.barrier(group=group)
  if int(agree.item()) == 0:
    return _NcclGather(nbytes, world, group, device)
  return _PeerGather(nbytes, world, rank, device, recv, flags, peer_recv, peer_flags)
