"""world_size-2 gloo test (CPU) of the block-sharding + all-gather exchange that
wraps the root solver (DS:2841-2879).  The root function is a CPU stand-in; the
partition, filler padding and order restoration are the code under test."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _fake_root(xs, ps, pads, **kw):
  """Deterministic per-matrix function so order mistakes are visible; fillers
  (padding 0) give zeros like DS:930-937."""
  scale = ps.to(torch.float32)[:, None, None] * (pads > 0)[:, None, None]
  roots = xs * scale
  metrics = torch.stack([xs.sum((1, 2)), ps.float(), pads.float(), xs[:, 0, 0],
                         torch.ones(len(ps))], 1)
  return roots, metrics


def _worker(rank, world, port, n_stats, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from precondition_b200.distributed_shampoo import sharded_inverse_pth_roots
  g = torch.Generator().manual_seed(0)
  stats = torch.randn((n_stats, 4, 4), generator=g)
  exps = torch.arange(1, n_stats + 1, dtype=torch.int32)
  roots, metrics = sharded_inverse_pth_roots(stats, exps, world, rank, None, root_fn=_fake_root)
  want = stats * exps.float()[:, None, None]
  ok = torch.equal(roots, want) and roots.shape[0] == n_stats
  ok = ok and torch.equal(metrics[:, 1], exps.float())
  ok = ok and bool((metrics[:, 2] == 4).all())
  out[rank] = bool(ok)
  dist.destroy_process_group()


def _fake_fd(grams, prevs, exps, r, pads, **kw):
  """Stand-in for the sketch update: fillers (padding 0) give zeros (DS:1265-1268)."""
  new = (prevs + grams[:, :, :r + 2] * exps.float()[:, None, None]) * (pads > 0)[:, None, None]
  return new, torch.zeros((len(exps), 5))


def _fd_worker(rank, world, port, n_stats, out):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from precondition_b200.distributed_shampoo import sharded_fd_updates
  g = torch.Generator().manual_seed(1)
  d, r = 6, 2
  grams = torch.randn((n_stats, d, d), generator=g)
  prevs = torch.randn((n_stats, d, r + 2), generator=g)
  exps = torch.arange(1, n_stats + 1, dtype=torch.int32)
  pads = torch.full((n_stats,), d, dtype=torch.int32)
  new = sharded_fd_updates(grams, prevs, exps, pads, r, world, rank, None, fd_fn=_fake_fd)
  want = prevs + grams[:, :, :r + 2] * exps.float()[:, None, None]
  out[rank] = bool(torch.equal(new, want) and new.shape[0] == n_stats)
  dist.destroy_process_group()


def _run(world, n_stats, worker=None):
  worker = worker or _worker
  port = _free_port()
  ctx = mp.get_context("spawn")
  with ctx.Manager() as mgr:
    out = mgr.dict()
    procs = [ctx.Process(target=worker, args=(r, world, port, n_stats, out))
             for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert all(out[r] for r in range(world)), dict(out)


def test_two_ranks_uneven_count():
  _run(2, 5)  # 5 statistics over 2 ranks: one (I, exponent 1, padding 0) filler


def test_two_ranks_even_count():
  _run(2, 6)


def test_more_ranks_than_statistics():
  _run(3, 2)


def test_sketch_updates_two_ranks_uneven_count():
  """Same partition / filler / gather logic around the Sketchy update (DS:2706-2738)."""
  _run(2, 3, _fd_worker)
