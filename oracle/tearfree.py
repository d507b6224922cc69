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


# --- sketchy (TF/sketchy.py) ---------------------------------------------------------------------
class SketchyLeaf:
  """Per axis: eigvecs [d, k], eigvals [k] (singular values), inv_eigvals [k], tail, inv_tail."""

  def __init__(self, shape, rank):
    if any(d == 1 for d in shape):
      raise ValueError("unit dimensions")
    self.shape = list(shape)
    self.axes = []
    for d in shape:
      k = min(d, rank)
      self.axes.append(dict(eigvecs=np.zeros((d, k), f32), eigvals=np.zeros(k, f32),
                            inv_eigvals=np.zeros(k, f32), tail=f32(0), inv_tail=f32(0)))

  def update_axis(self, g, dim, st, smd, epsilon, relative):  # TF/sketchy.py:380-470
    d, k = st["eigvecs"].shape
    sketch = st["eigvecs"] * st["eigvals"][None, :]
    rest = [i for i in range(g.ndim) if i != dim]
    g_dm = g.transpose([dim] + rest).reshape(d, -1)
    decay = np.sqrt(f32(smd))
    updated = np.concatenate([sketch * decay, g_dm], axis=1).astype(f32)
    updated = np.linalg.qr(updated.T, mode="r").T
    if np.isfinite(updated).all():
      u, s, _ = np.linalg.svd(updated, full_matrices=False)
    else:
      m = min(updated.shape)
      u, s = np.full((d, m), np.nan, f32), np.full((m,), np.nan, f32)
    cutoff = max(s[k], 0.0) if k < len(s) else f32(0.0)
    top = np.maximum(s[:k], 0.0)
    deflated = np.sqrt(np.maximum(0.0, top - cutoff)) * np.sqrt(top + cutoff)
    tail = st["tail"] * decay + cutoff**2
    undeflated = np.square(np.maximum(top, 0.0)) + st["tail"] * decay
    mask = deflated > 0
    alpha = f32(-1.0 / (2 * g.ndim))
    eps = np.max(undeflated) * epsilon if (relative and epsilon > 0) else epsilon
    st["eigvecs"] = (u[:, :k] * mask).astype(f32)
    st["inv_eigvals"] = np.where(mask, (undeflated + eps) ** alpha, 0.0).astype(f32)
    st["eigvals"] = (deflated * mask).astype(f32)
    st["inv_tail"] = f32((tail + eps) ** alpha if tail > 0 else 0.0)
    st["tail"] = f32(tail)

  def update(self, g, count, smd, epsilon, relative, update_freq):
    if count % update_freq == 0:
      for dim, st in enumerate(self.axes):
        self.update_axis(g, dim, st, smd, epsilon, relative)
    roll = tuple(range(1, g.ndim)) + (0,)
    for st in self.axes:  # TF/sketchy.py:322-361
      v = st["eigvecs"]
      basis = np.tensordot(g, v, axes=[[0], [0]])
      low = np.tensordot(basis, v, axes=[[g.ndim - 1], [1]])
      g = np.transpose(g, roll)
      complement = g - low
      scaled = np.tensordot(basis * st["inv_eigvals"], v, axes=[[g.ndim - 1], [1]])
      g = (scaled + st["inv_tail"] * complement).astype(f32)
    return g


# --- the optimizer (TF/optimizer.py:61-99) ---------------------------------------------------------
class Tearfree:
  """lists of arrays in, lists of arrays out.  graft in {'none', 'sgd', 'rmsprop'}."""

  def __init__(self, params, learning_rate, graft="rmsprop", graft_decay=0.999,
               graft_epsilon=1e-23, start_preconditioning_step=0,
               skip_preconditioning_any_dim_gt=4096, skip_preconditioning_rank1=True,
               merge_dims=1024, block_size=1024, update_preconditioners_freq=1,
               update_statistics_freq=1, second_moment_decay=0.999, ema=False, nesterov=True,
               momentum_decay=0.9, weight_decay=0.0, weight_decay_after_momentum=True,
               second_order="shampoo", sketchy_rank=128, sketchy_epsilon=1e-7,
               sketchy_relative_epsilon=True, sketchy_decay=0.999, sketchy_update_freq=1):
    self.__dict__.update(locals())
    self.count = 0
    self.masked = []
    for p in params:  # TF/grafting.py:324-336
      skip = graft != "none" and ((skip_preconditioning_rank1 and p.ndim <= 1) or
                                  any(s > skip_preconditioning_any_dim_gt for s in p.shape))
      self.masked.append(skip)
    if second_order == "sketchy":  # TF/second_order.py:84-85: no padding
      self.block_size = 0
      self.leaves = [None if m else SketchyLeaf(derive_shapes(p.shape, merge_dims, 0)[1],
                                                sketchy_rank)
                     for p, m in zip(params, self.masked)]
    else:
      self.leaves = [None if m else
                     ShampooLeaf(derive_shapes(p.shape, merge_dims, block_size)[1], block_size)
                     for p, m in zip(params, self.masked)]
    self.acc = [np.zeros_like(p, dtype=f32) for p in params]
    self.trace = [np.zeros_like(p, dtype=f32) for p in params]

  def direction(self, g, leaf):
    merged = merge(g, self.merge_dims, self.block_size)
    if self.second_order == "sketchy":
      y = leaf.update(merged, self.count, self.sketchy_decay, self.sketchy_epsilon,
                      self.sketchy_relative_epsilon, self.sketchy_update_freq)
    else:
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
