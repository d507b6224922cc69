"""numpy restatement of precondition/tearfree (TEST INFRASTRUCTURE -- never imported by the
product): blocked Shampoo direction, grafting (NONE / SGD / RMSPROP), momentum, weight decay and
the learning rate, on lists of arrays.  Reference: /root/reference/precondition/tearfree/
{optimizer.py:61-99, second_order.py:55-75, reshaper.py:52-133, shampoo.py:150-547,
grafting.py:89-300, momentum.py:81-139} ("TF" below).  Pinned by tests/golden/tearfree.npz, which
oracle/gen_golden.py records from the unmodified reference sources."""
from __future__ import annotations

import math

import numpy as np

from oracle import optimizer as O

f32 = np.float32


# --- reshaper (TF/reshaper.py:52-133) ---------------------------------------------------------
def derive_shapes(shape, merge_dims, block_size):
  merged = O.merge_small_dims(list(shape), merge_dims)
  if merged == [1]:
    return [], []
  if block_size == 0:
    return merged, merged
  padded = [((s + block_size - 1) // block_size) * block_size if s >= block_size else s
            for s in merged]
  return merged, padded


def merge(x, merge_dims, block_size):
  merged, padded = derive_shapes(x.shape, merge_dims, block_size)
  y = x.reshape(merged)
  if block_size > 0 and merged:
    y = np.pad(y, [(0, p - m) for p, m in zip(padded, merged)])
  return y


def unmerge(y, shape, merge_dims, block_size):
  merged, _ = derive_shapes(shape, merge_dims, block_size)
  if block_size > 0:
    y = y[tuple(slice(0, m) for m in merged)]
  return y.reshape(shape)


# --- blocked shampoo (TF/shampoo.py) --------------------------------------------------------------
def blocks_of(x, block_size):
  """List of contiguous blocks of x (every axis >= block_size is cut), in the order of the
  reference's N axis (left large axis major)."""
  large = [i for i, d in enumerate(x.shape) if d >= block_size]
  ranges = [range(x.shape[i] // block_size) for i in large]
  out = []
  for idx in np.ndindex(*[len(r) for r in ranges]):
    sl = [slice(None)] * x.ndim
    for ax, k in zip(large, idx):
      sl[ax] = slice(k * block_size, (k + 1) * block_size)
    out.append(tuple(sl))
  return out


def pth_inv_root(p, cov):  # TF/shampoo.py:440-448
  eps = 1e-6
  w, v = np.linalg.eigh(cov.astype(f32))
  mask = w <= f32(eps) * np.max(w)
  half = np.where(mask, f32(1.0), w) ** f32(-0.5 / p)
  half = np.where(mask, f32(0.0), half).astype(f32)
  hv = (half[None, :] * v).astype(f32)
  return (hv @ hv.T).astype(f32)


class ShampooLeaf:
  def __init__(self, shape, block_size):
    self.shape, self.block_size = list(shape), block_size
    if any(d == 1 for d in shape):
      raise ValueError("unit dimensions")
    self.slices = blocks_of(np.zeros(shape, np.int8), block_size)
    dims = [min(d, block_size) for d in shape]
    self.stats = [[np.zeros((d, d), f32) for d in dims] for _ in self.slices]
    self.roots = [[np.eye(d, dtype=f32) for d in dims] for _ in self.slices]

  def update(self, g, count, decay, stat_freq, precond_freq):
    r = g.ndim
    out = np.zeros_like(g)
    for n, sl in enumerate(self.slices):
      blk = g[sl]
      if count % stat_freq == 0:  # TF/shampoo.py:278-281, 409-430
        for a in range(r):
          rest = [i for i in range(r) if i != a]
          cov = np.tensordot(blk, blk, axes=(rest, rest)).astype(f32)
          old = self.stats[n][a]
          self.stats[n][a] = (old + cov if decay == 1.0
                              else old * f32(decay) + cov * f32(1 - decay)).astype(f32)
      if count % precond_freq == 0:  # TF/shampoo.py:291-296, 451-458
        for a in range(r):
          self.roots[n][a] = pth_inv_root(2 * r, self.stats[n][a])
      y = blk
      for a in range(r):  # TF/shampoo.py:461-491: contract the inner axis of every root
        y = np.moveaxis(np.tensordot(self.roots[n][a], y, axes=([1], [a])), 0, a).astype(f32)
      out[sl] = y
    return out


# --- the optimizer (TF/optimizer.py:61-99) ---------------------------------------------------------
class Tearfree:
  """lists of arrays in, lists of arrays out.  graft in {'none', 'sgd', 'rmsprop'}."""

  def __init__(self, params, learning_rate, graft="rmsprop", graft_decay=0.999,
               graft_epsilon=1e-23, start_preconditioning_step=0,
               skip_preconditioning_any_dim_gt=4096, skip_preconditioning_rank1=True,
               merge_dims=1024, block_size=1024, update_preconditioners_freq=1,
               update_statistics_freq=1, second_moment_decay=0.999, ema=False, nesterov=True,
               momentum_decay=0.9, weight_decay=0.0, weight_decay_after_momentum=True):
    self.__dict__.update(locals())
    self.count = 0
    self.masked = []
    for p in params:  # TF/grafting.py:324-336
      skip = graft != "none" and ((skip_preconditioning_rank1 and p.ndim <= 1) or
                                  any(s > skip_preconditioning_any_dim_gt for s in p.shape))
      self.masked.append(skip)
    self.leaves = [None if m else
                   ShampooLeaf(derive_shapes(p.shape, merge_dims, block_size)[1], block_size)
                   for p, m in zip(params, self.masked)]
    self.acc = [np.zeros_like(p, dtype=f32) for p in params]
    self.trace = [np.zeros_like(p, dtype=f32) for p in params]

  def direction(self, g, leaf):
    merged = merge(g, self.merge_dims, self.block_size)
    y = leaf.update(merged, self.count, self.second_moment_decay, self.update_statistics_freq,
                    self.update_preconditioners_freq)
    return unmerge(y, g.shape, self.merge_dims, self.block_size)

  def update(self, grads, params):
    lr = self.learning_rate(self.count) if callable(self.learning_rate) else self.learning_rate
    outs = []
    for i, (g, p) in enumerate(zip(grads, params)):
      g = g.astype(f32)
      base = None if self.masked[i] else self.direction(g, self.leaves[i])
      if self.graft == "none":
        x = base
      else:
        if self.graft == "sgd":
          u = g
        else:  # TF/grafting.py:205-222
          d = self.graft_decay
          sq = np.square(g)
          self.acc[i] = (sq + self.acc[i] if d == 1.0
                         else sq * f32(1 - d) + f32(d) * self.acc[i]).astype(f32)
          u = (g * (f32(1.0) / np.sqrt(self.acc[i] + f32(self.graft_epsilon)))).astype(f32)
        if base is None:
          x = u
        else:  # TF/grafting.py:256-272
          bn = np.linalg.norm(base)
          mult = np.linalg.norm(u) / bn if bn > 0 else f32(0.0)
          x = (base * f32(mult)).astype(f32) if self.count >= self.start_preconditioning_step else u
      wd = lambda t: (t + f32(self.weight_decay) * p).astype(f32)
      if self.weight_decay > 0 and not self.weight_decay_after_momentum:
        x = wd(x)
      if self.momentum_decay:  # TF/momentum.py:84-91
        m = f32(self.momentum_decay)
        if self.ema:
          x = (x * f32(1 - self.momentum_decay)).astype(f32)
        self.trace[i] = (x + m * self.trace[i]).astype(f32)
        x = (x + m * self.trace[i]).astype(f32) if self.nesterov else self.trace[i]
      if self.weight_decay > 0 and self.weight_decay_after_momentum:
        x = wd(x)
      outs.append((f32(-1.0 * lr) * x).astype(f32))
    self.count += 1
    return outs
