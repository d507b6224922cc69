"""Mirror of precondition/tearfree/shampoo.py (TF/shampoo.py:30-547): blocked Shampoo second-order
direction.  Per tensor (already merged / padded by ``reshaper``) and per axis: block-diagonal
Gram statistics [N, B, B], their pseudo-inverse 2*rank-th roots, and the mode products of the
blocked gradient with those roots.

Device side (no torch math on the path):
  * statistics:  grouped GEMM descriptors over strided unfoldings of the blocked gradient, the
    EMA folded into the epilogue (``pc_grouped_gemm_tc`` when the block is a multiple of 128,
    ``pc_grouped_gemm`` otherwise) -- TF/shampoo.py:409-430
  * roots:       ``pc_pinv_pth_root_eigh_batched`` over all statistics of one size at once --
    TF/shampoo.py:433-458
  * direction:   one grouped GEMM launch list per axis (mode products) -- TF/shampoo.py:461-491
Statistics and roots of equal size live in one buffer per size so the root solver sees one batch."""
import dataclasses
import functools
import math
import os
from typing import Any, NamedTuple, Optional, Sequence

import torch

from precondition_b200 import _lib
from precondition_b200 import ops
from precondition_b200.tearfree import _tree
from precondition_b200.tearfree import praxis_shim

EIGH_CUTOFF = 1e-6  # `eps` of _pth_inv_root, TF/shampoo.py:442
WARM_START = os.environ.get("PC_TF_EIGH_WARM", "1") != "0"  # previous eigenvectors as the basis


@dataclasses.dataclass
class Options:
  """Shampoo covariance approximation options (TF/shampoo.py:30-51)."""
  block_size: int = 1024
  update_preconditioners_freq: int = 1
  update_statistics_freq: int = 1
  second_moment_decay: float = 0.999


class _AxesBlocks(NamedTuple):
  """Statistics and their inverse roots for one tensor: per axis [N, B, B] (TF/shampoo.py:70-109)."""
  stats: list
  roots: list


class _ShampooState(NamedTuple):
  count: torch.Tensor  # scalar int32 (host)
  blocks: Any          # tree of _AxesBlocks


@dataclasses.dataclass(frozen=True)
class _BlocksMetadata:
  """Indexing information of one blocked tensor (TF/shampoo.py:119-147)."""
  block_sizes: list
  num_blocks: int
  debug_name: str
  large_block_size: int
  param_shape: list
  large_axes: list
  blocks_per_large_axis: list
  blocks_axis: int


def _blocks_metadata(options: Options, param_shape: Sequence[int], debug: str) -> _BlocksMetadata:
  dims = [min(dim, options.block_size) for dim in param_shape]  # TF/shampoo.py:150-171
  large_axes = [i for i, d in enumerate(param_shape) if d >= options.block_size]
  blocks_per_large_axis = [param_shape[i] // options.block_size for i in large_axes]
  num_blocks = math.prod(blocks_per_large_axis + [1])
  return _BlocksMetadata(block_sizes=dims, num_blocks=num_blocks, debug_name=debug,
                         large_block_size=options.block_size, large_axes=large_axes,
                         param_shape=list(param_shape),
                         blocks_per_large_axis=blocks_per_large_axis,
                         blocks_axis=min(large_axes, default=0))


def _validate(options: Options) -> None:  # TF/shampoo.py:174-199
  if options.block_size <= 1:
    raise ValueError(f"block_size ({options.block_size}) must be >1")
  if options.update_preconditioners_freq <= 0:
    raise ValueError("update_preconditioners_freq ({}) must be positive".format(
        options.update_preconditioners_freq))
  if options.update_statistics_freq <= 0:
    raise ValueError("update_statistics_freq ({}) must be positive".format(
        options.update_statistics_freq))
  if not (0 <= options.second_moment_decay <= 1):
    raise ValueError(f"second_moment_decay ({options.second_moment_decay}) "
                     "should be in [0, 1]")
  if options.block_size > ops.EIGH_MAX_DIM:
    raise ValueError(f"block_size ({options.block_size}) is above the eigensolver's limit "
                     f"({ops.EIGH_MAX_DIM})")


def _check_param(options: Options, path, shape):  # TF/shampoo.py:205-229
  if any(dim == 1 for dim in shape):
    raise ValueError("param {} shape ({}) has unit dimensions".format(path, tuple(shape)))
  if sum(dim >= options.block_size for dim in shape) > 2:
    raise ValueError("param {} shape ({}) has >2 large dims for block size {}".format(
        path, tuple(shape), options.block_size))
  if any(dim % options.block_size != 0 for dim in shape if dim >= options.block_size):
    raise ValueError("param {} shape ({}) has large dims indivisible by block size {}".format(
        path, tuple(shape), options.block_size))


# ---------------------------------------------------------------------------
# blocking: the N axis goes first here (one contiguous block per n) -- a fixed permutation of
# the reference's layout (TF/shampoo.py:302-406), which only ever contracts within a block
# ---------------------------------------------------------------------------
def _split_shape(meta: _BlocksMetadata):
  """Shape with every large axis split into (blocks, block) and the permutation that brings
  the block-count axes to the front."""
  shape, counts, inner = [], [], []
  for i, d in enumerate(meta.param_shape):
    if i in meta.large_axes:
      counts.append(len(shape))
      shape += [d // meta.large_block_size, meta.large_block_size]
      inner.append(len(shape) - 1)
    else:
      inner.append(len(shape))
      shape.append(d)
  return shape, counts + inner


def _blockify(x: torch.Tensor, meta: _BlocksMetadata) -> torch.Tensor:
  """[N, b_0, .., b_{r-1}] view (a copy only when two axes are blocked)."""
  shape, perm = _split_shape(meta)
  return x.reshape(shape).permute(perm).reshape([meta.num_blocks] + meta.block_sizes)


def _deblockify(xb: torch.Tensor, meta: _BlocksMetadata) -> torch.Tensor:
  shape, perm = _split_shape(meta)
  inv = [0] * len(perm)
  for i, p in enumerate(perm):
    inv[p] = i
  return xb.reshape([shape[p] for p in perm]).permute(inv).reshape(meta.param_shape)


# ---------------------------------------------------------------------------
# engine: buffers and static launch lists for one parameter tree
# ---------------------------------------------------------------------------
class _Leaf:
  def __init__(self, path, meta):
    self.path, self.meta = path, meta
    self.slots = []  # per axis: (dim, first index in the bucket of that dim)


class _Engine:
  def __init__(self, options: Options, params):
    self.options = options
    self.leaves = []
    self.device = None
    counts = {}

    def plan(path, p):
      if not isinstance(p, torch.Tensor):
        raise TypeError(f"tearfree shampoo: parameter {path} is not a tensor")
      if not p.is_cuda:
        raise RuntimeError("tearfree shampoo needs CUDA tensors: there is no CPU fallback")
      _check_param(self.options, path, p.shape)
      self.device = self.device or p.device
      leaf = _Leaf(path, _blocks_metadata(self.options, list(p.shape), str(path)))
      for d in leaf.meta.block_sizes:
        leaf.slots.append((d, counts.get(d, 0)))
        counts[d] = counts.get(d, 0) + leaf.meta.num_blocks
      self.leaves.append(leaf)
      return leaf

    self.plan_tree = _tree.tree_map_with_path(plan, params)
    dev = self.device
    self.stats, self.roots, self.ps, self.eigvecs, self.eig_valid = {}, {}, {}, {}, {}
    for d, cnt in counts.items():
      self.stats[d] = torch.zeros((cnt, d, d), dtype=torch.float32, device=dev)
      self.roots[d] = torch.eye(d, dtype=torch.float32, device=dev).repeat(cnt, 1, 1)
      self.ps[d] = torch.zeros(cnt, dtype=torch.int32)
      # eigenvectors of the last solve, the warm start of the next one (not part of the state:
      # after a restore the first solve simply starts cold)
      self.eigvecs[d] = torch.zeros((cnt, d, d), dtype=torch.float32, device=dev)
      self.eig_valid[d] = False
    for leaf in self.leaves:
      m = leaf.meta
      for d, first in leaf.slots:
        self.ps[d][first:first + m.num_blocks] = 2 * len(m.param_shape)  # TF/shampoo.py:454
      leaf.xb = torch.empty([m.num_blocks] + m.block_sizes, dtype=torch.float32, device=dev)
      leaf.y = [torch.empty_like(leaf.xb) for _ in range(min(2, len(m.block_sizes)))]
    self.ps = {d: t.to(dev) for d, t in self.ps.items()} if dev is not None else {}
    self._build_lists()

  # views of the state in the reference's layout
  def blocks_tree(self):
    def view(leaf):
      n = leaf.meta.num_blocks
      return _AxesBlocks(stats=[self.stats[d][f:f + n] for d, f in leaf.slots],
                         roots=[self.roots[d][f:f + n] for d, f in leaf.slots])
    return _tree.tree_map(view, self.plan_tree, is_leaf=lambda x: isinstance(x, _Leaf))

  def owns(self, blocks) -> bool:
    mine = _tree.tree_leaves(self.blocks_tree(), is_leaf=lambda x: isinstance(x, _AxesBlocks))
    theirs = _tree.tree_leaves(blocks, is_leaf=lambda x: isinstance(x, _AxesBlocks))
    if len(mine) != len(theirs):
      raise ValueError("tearfree shampoo: state does not match the parameters it was built for")
    return all(a.data_ptr() == b.data_ptr()
               for m, t in zip(mine, theirs) for a, b in zip(m.stats + m.roots, t.stats + t.roots))

  def adopt(self, blocks):
    """Copies a foreign state (a restored checkpoint) into the engine's buffers."""
    mine = _tree.tree_leaves(self.blocks_tree(), is_leaf=lambda x: isinstance(x, _AxesBlocks))
    theirs = _tree.tree_leaves(blocks, is_leaf=lambda x: isinstance(x, _AxesBlocks))
    for m, t in zip(mine, theirs):
      for a, b in zip(m.stats + m.roots, t.stats + t.roots):
        a.copy_(b.to(device=a.device, dtype=torch.float32))
    self.eig_valid = {d: False for d in self.eig_valid}

  def _build_lists(self):
    decay = self.options.second_moment_decay
    w_old, w_new = (1.0, 1.0) if decay == 1.0 else (decay, 1 - decay)  # TF/shampoo.py:544-547
    stat_descs, apply_descs = [], {}
    for leaf in self.leaves:
      m = leaf.meta
      bs, r = m.block_sizes, len(m.block_sizes)
      blk = math.prod(bs)
      for a in range(r):
        d, first = leaf.slots[a]
        pre, suf = math.prod(bs[:a]), math.prod(bs[a + 1:])
        src = leaf.xb if a == 0 else leaf.y[(a - 1) % 2]
        dst = leaf.y[a % 2]
        for n in range(m.num_blocks):
          x0 = leaf.xb.data_ptr() + 4 * n * blk
          s = self.stats[d][first + n]
          # statistics: C(i, j) = sum over everything but axis a, X(i, k) = xb[n][p, i, q],
          # k = p * suf + q
          g = _lib.GemmDesc()
          g.a = g.b = x0
          g.c = g.c_in = s.data_ptr()
          g.a_iinner, g.a_sio, g.a_si = d, 0, suf
          g.a_kinner, g.a_sko, g.a_ski = suf, d * suf, 1
          g.b_sj, g.b_kinner, g.b_sko, g.b_ski = suf, suf, d * suf, 1
          g.c_iinner, g.c_sio, g.c_sii = d, 0, d
          g.m, g.n, g.k, g.alpha, g.beta = d, d, pre * suf, w_new, w_old
          stat_descs.append(g)
          # direction: mode product of axis a with the (symmetric) root
          root = self.roots[d][first + n].data_ptr()
          i0, o0 = src.data_ptr() + 4 * n * blk, dst.data_ptr() + 4 * n * blk
          lst = apply_descs.setdefault(a, [])
          if suf == 1 and r > 1:
            # last axis: out[p, j] = sum_k in[p, k] root[j, k]
            g = _lib.GemmDesc()
            g.a, g.b, g.c, g.c_in = i0, root, o0, None
            g.a_iinner, g.a_sio, g.a_si = pre, 0, d
            g.a_kinner, g.a_sko, g.a_ski = d, 0, 1
            g.b_sj, g.b_kinner, g.b_sko, g.b_ski = d, d, 0, 1
            g.c_iinner, g.c_sio, g.c_sii = pre, 0, d
            g.m, g.n, g.k, g.alpha, g.beta = pre, d, d, 1.0, 0.0
            lst.append(g)
          else:
            # out[p, i, q] = sum_k root[i, k] in[p, k, q], one product per p
            for p in range(pre):
              g = _lib.GemmDesc()
              g.a, g.c_in = root, None
              g.b = i0 + 4 * p * d * suf
              g.c = o0 + 4 * p * d * suf
              g.a_iinner, g.a_sio, g.a_si = d, 0, d
              g.a_kinner, g.a_sko, g.a_ski = d, 0, 1
              g.b_sj, g.b_kinner, g.b_sko, g.b_ski = 1, d, 0, suf
              g.c_iinner, g.c_sio, g.c_sii = d, 0, suf
              g.m, g.n, g.k, g.alpha, g.beta = d, suf, d, 1.0, 0.0
              lst.append(g)
    tc_ok = self.device is not None and bool(_lib.load().pc_device_supports_tcgen05())

    def lists(descs):
      tc = [g for g in descs if tc_ok and ops.tc_gemm_eligible(g)]
      simt = [g for g in descs if not (tc_ok and ops.tc_gemm_eligible(g))]
      out = []
      if tc:
        out.append(ops.TcGemmList(tc, self.device))
      if simt:
        out.append(ops.SimtGemmLists(simt, self.device))
      return out

    self.stat_lists = lists(stat_descs)
    # axis a + 1 reads what axis a wrote: one launch group per axis, in order
    self.apply_lists = [lists(apply_descs[a]) for a in sorted(apply_descs)]

  def step(self, updates, count: int, alias_outputs: bool = False):
    """Statistics / roots as scheduled, then the direction for every leaf.  Returns the
    direction tree (tensors shaped like the updates)."""
    o = self.options

    def load(leaf, u):
      if not isinstance(u, torch.Tensor) or not u.is_cuda:
        raise RuntimeError("tearfree shampoo needs CUDA tensors: there is no CPU fallback")
      if not u.is_floating_point():
        raise TypeError(f"tearfree shampoo: updates must be floating point, got {u.dtype}")
      assert list(u.shape) == leaf.meta.param_shape, (u.shape, leaf.meta.param_shape)
      shape, perm = _split_shape(leaf.meta)
      leaf.xb.view([shape[p] for p in perm]).copy_(u.reshape(shape).permute(perm))
      return leaf

    is_leaf = lambda x: isinstance(x, _Leaf)
    _tree.tree_map(load, self.plan_tree, updates, is_leaf=is_leaf)
    if count % o.update_statistics_freq == 0:      # TF/shampoo.py:278-281
      for lst in self.stat_lists:
        lst.run()
    if count % o.update_preconditioners_freq == 0:  # TF/shampoo.py:291-296
      for d in self.stats:
        ops.pinv_pth_root_eigh_batched(self.stats[d], self.ps[d], EIGH_CUTOFF, out=self.roots[d],
                                       eigvecs=self.eigvecs[d], eigvecs_valid=self.eig_valid[d])
        self.eig_valid[d] = WARM_START
    for group in self.apply_lists:
      for lst in group:
        lst.run()

    def result(leaf):
      r = len(leaf.meta.block_sizes)
      out = _deblockify(leaf.y[(r - 1) % 2], leaf.meta)
      if alias_outputs:
        return out  # possibly a view of the engine's buffer: valid until the next step
      return out.clone() if out.data_ptr() == leaf.y[(r - 1) % 2].data_ptr() else out.contiguous()

    return _tree.tree_map(result, self.plan_tree, is_leaf=is_leaf)


def apply(options: Options, _alias_outputs: bool = False
          ) -> praxis_shim.ShardedGradientTransformation:
  """Gradient transform for (blocked) shampoo preconditioning (TF/shampoo.py:54-67).
  `_alias_outputs` (private): the returned directions may be views of the engine's buffers, valid
  until the next update -- for callers that consume them at once (``optimizer.tearfree``)."""
  _validate(options)
  holder = {}

  def init_fn(params) -> _ShampooState:
    eng = holder["engine"] = _Engine(options, params)
    return _ShampooState(count=torch.zeros([], dtype=torch.int32), blocks=eng.blocks_tree())

  def update_fn(updates, state: _ShampooState, params=None):
    del params
    eng = holder.get("engine")
    if eng is None:
      eng = holder["engine"] = _Engine(options, updates)
    if not eng.owns(state.blocks):
      eng.adopt(state.blocks)
    new_updates = eng.step(updates, int(state.count), _alias_outputs)
    return new_updates, _ShampooState(count=state.count + 1, blocks=eng.blocks_tree())

  return praxis_shim.ShardedGradientTransformation(init_fn, update_fn,
                                                   functools.partial(_pspec, options))


def _pspec(options: Options, params: praxis_shim.NestedHParams) -> praxis_shim.NestedHParams:
  """Sharding specification of the shampoo state: everything replicated (TF/shampoo.py:232-266)."""
  count_pspec = praxis_shim.WeightHParams(shape=[], init=None, dtype=torch.int32,
                                          collections=None, tensor_split_dims_mapping=[])

  def make_blocks_pspec(path, param):
    meta = _blocks_metadata(options, param.shape, str(path))
    replicated = functools.partial(praxis_shim.WeightHParams, init=None, dtype=torch.float32,
                                   collections=None, tensor_split_dims_mapping=[-1, -1, -1])
    stats = [replicated((meta.num_blocks, d, d)) for d in meta.block_sizes]
    return dict(stats=stats, roots=stats)

  return dict(count=count_pspec,
              blocks=_tree.tree_map_with_path(make_blocks_pspec, params,
                                              is_leaf=lambda x: hasattr(x, "shape")))
