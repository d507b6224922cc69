"""Mirror of precondition/tearfree/sketchy.py (TF/sketchy.py:29-515): Sketchy second-order
direction -- per tensor axis a rank-k frequent-directions sketch of the gradient covariance, the
direction is  V diag(inv_eigvals) V^T g + inv_tail (g - V V^T g)  along every axis.

Device side: the Gram matrix of every mode unfolding comes from the grouped GEMM; the sketch
update is ``pc_fd_update_batched`` in its tearfree mode (the top k + 1 eigenpairs of
decay * V diag(s^2) V^T + G G^T stand in for the QR + SVD of the stacked factor,
TF/sketchy.py:386-400), batched over all axes of equal (padded) size; the operator
inv_tail I + V diag(inv_eigvals - inv_tail) V^T is formed once per update
(``pc_low_rank_to_dense``) and applied as mode products by the grouped GEMM.

Axes shorter than rank + 3 are embedded in a zero-padded problem of that size (the packed
sketch layout needs rank + 2 < d); their state views show the true [d, k] / [k] shapes.
Not built: memory_alloc, ekfac_svd, linear_approx_tail, add_ggt (NotImplementedError)."""
import dataclasses
import functools
import math
from typing import Any, NamedTuple, Optional

import torch

from precondition_b200 import _lib
from precondition_b200 import ops
from precondition_b200.tearfree import _tree
from precondition_b200.tearfree import praxis_shim


@dataclasses.dataclass
class Options:
  """Sketchy covariance approximation options (TF/sketchy.py:29-62); same fields and defaults."""
  epsilon: float = 1e-7
  rank: int = 128
  relative_epsilon: bool = True
  second_moment_decay: float = 0.999
  update_freq: int = 1
  add_ggt: bool = False
  memory_alloc: Optional[dict] = None
  ekfac_svd: bool = False
  linear_approx_tail: bool = False


class _AxisState(NamedTuple):
  """The covariance sketch of one tensor axis (TF/sketchy.py:78-90)."""
  eigvecs: torch.Tensor
  eigvals: torch.Tensor
  inv_eigvals: torch.Tensor
  tail: torch.Tensor
  inv_tail: torch.Tensor
  ema_ggt: Any
  svd_result_u: Any
  svd_result_s: Any
  inv_prev_tail: Any


class _TensorState(NamedTuple):
  axes: list


class _SketchyState(NamedTuple):
  count: torch.Tensor
  sketches: Any


def _validate(options: Options) -> None:  # TF/sketchy.py:104-118
  if options.update_freq <= 0:
    raise ValueError("update_freq ({}) must be positive".format(options.update_freq))
  if not (0 <= options.second_moment_decay <= 1):
    raise ValueError(f"second_moment_decay ({options.second_moment_decay}) "
                     "should be in [0, 1]")
  if options.rank <= 0:
    raise ValueError(f"rank ({options.rank}) must be at least 1")
  for name in ("add_ggt", "memory_alloc", "ekfac_svd", "linear_approx_tail"):
    if getattr(options, name):
      raise NotImplementedError(f"tearfree sketchy option {name} is not built")
  if options.rank + 1 + 32 > 512:
    raise ValueError(f"rank ({options.rank}) above the eigen-solver's limit of 479")


class _Axis:
  def __init__(self, d, k, slot):
    self.d, self.k, self.slot = d, k, slot  # slot: (bucket key, index)


class _Leaf:
  def __init__(self, path, shape):
    self.path, self.shape, self.axes = path, list(shape), []


class _Engine:
  def __init__(self, options: Options, params):
    self.options, self.leaves, self.device = options, [], None
    counts = {}

    def plan(path, p):
      if not isinstance(p, torch.Tensor) or not p.is_cuda:
        raise RuntimeError("tearfree sketchy needs CUDA tensors: there is no CPU fallback")
      if any(dim == 1 for dim in p.shape):  # TF/sketchy.py:160-163
        raise ValueError("param {} shape ({}) has unit dimensions".format(path, tuple(p.shape)))
      self.device = self.device or p.device
      leaf = _Leaf(path, p.shape)
      for d in p.shape:
        k = min(d, options.rank)     # TF/sketchy.py:176
        key = (max(d, k + 3), k)     # (problem size, sketch rank)
        leaf.axes.append(_Axis(d, k, (key, counts.get(key, 0))))
        counts[key] = counts.get(key, 0) + 1
      self.leaves.append(leaf)
      return leaf

    self.plan_tree = _tree.tree_map_with_path(plan, params)
    dev = self.device
    self.packed, self.gram, self.dense, self.ps, self.pads, self.scratch = {}, {}, {}, {}, {}, {}
    for (D, k), cnt in counts.items():
      self.packed[(D, k)] = torch.zeros((cnt, D, k + 2), dtype=torch.float32, device=dev)
      self.scratch[(D, k)] = torch.empty_like(self.packed[(D, k)])
      self.gram[(D, k)] = torch.zeros((cnt, D, D), dtype=torch.float32, device=dev)
      self.dense[(D, k)] = torch.zeros((cnt, D, D), dtype=torch.float32, device=dev)
      self.ps[(D, k)] = torch.zeros(cnt, dtype=torch.int32)
      self.pads[(D, k)] = torch.zeros(cnt, dtype=torch.int32)
    for leaf in self.leaves:
      for ax in leaf.axes:
        key, i = ax.slot
        self.ps[key][i] = 2 * len(leaf.shape)  # alpha = -1 / (2 ndim), TF/sketchy.py:440
        self.pads[key][i] = ax.d
      leaf.xb = torch.empty(leaf.shape, dtype=torch.float32, device=dev)
      leaf.y = [torch.empty_like(leaf.xb) for _ in range(min(2, len(leaf.shape)))]
    if dev is not None:
      self.ps = {k: v.to(dev) for k, v in self.ps.items()}
      self.pads = {k: v.to(dev) for k, v in self.pads.items()}
    self._build_lists()

  def sketches_tree(self):
    def view(leaf):
      axes = []
      for ax in leaf.axes:
        (D, k), i = ax.slot
        P = self.packed[(D, k)][i]
        axes.append(_AxisState(
            eigvecs=P[:ax.d, :k], eigvals=P[D - k:, k + 1], inv_eigvals=P[:k, k],
            tail=P[1, k + 1], inv_tail=P[0, k + 1], ema_ggt=praxis_shim.MaskedNode(),
            svd_result_u=praxis_shim.MaskedNode(), svd_result_s=praxis_shim.MaskedNode(),
            inv_prev_tail=praxis_shim.MaskedNode()))
      return _TensorState(axes)
    return _tree.tree_map(view, self.plan_tree, is_leaf=lambda x: isinstance(x, _Leaf))

  def _axis_states(self, sketches):
    return [a for t in _tree.tree_leaves(sketches, is_leaf=lambda x: isinstance(x, _TensorState))
            for a in t.axes]

  def owns(self, sketches) -> bool:
    mine, theirs = self._axis_states(self.sketches_tree()), self._axis_states(sketches)
    if len(mine) != len(theirs):
      raise ValueError("tearfree sketchy: state does not match the parameters it was built for")
    return all(m.eigvecs.data_ptr() == t.eigvecs.data_ptr() for m, t in zip(mine, theirs))

  def adopt(self, sketches):
    """Copies a foreign state (a restored checkpoint) into the packed buffers."""
    for m, t in zip(self._axis_states(self.sketches_tree()), self._axis_states(sketches)):
      for name in ("eigvecs", "eigvals", "inv_eigvals", "tail", "inv_tail"):
        getattr(m, name).copy_(getattr(t, name).to(device=self.device, dtype=torch.float32))
    for key in self.packed:
      ops.low_rank_to_dense(self.packed[key], key[1], out=self.dense[key])

  def _build_lists(self):
    gram_descs, apply_descs = [], {}
    for leaf in self.leaves:
      bs, r = leaf.shape, len(leaf.shape)
      for a, ax in enumerate(leaf.axes):
        (D, k), i = ax.slot
        d = ax.d
        pre, suf = math.prod(bs[:a]), math.prod(bs[a + 1:])
        # Gram of the mode-a unfolding into the top-left d x d of the padded problem
        g = _lib.GemmDesc()
        g.a = g.b = leaf.xb.data_ptr()
        g.c, g.c_in = self.gram[(D, k)][i].data_ptr(), None
        g.a_iinner, g.a_sio, g.a_si = d, 0, suf
        g.a_kinner, g.a_sko, g.a_ski = suf, d * suf, 1
        g.b_sj, g.b_kinner, g.b_sko, g.b_ski = suf, suf, d * suf, 1
        g.c_iinner, g.c_sio, g.c_sii = d, 0, D
        g.m, g.n, g.k, g.alpha, g.beta = d, d, pre * suf, 1.0, 0.0
        gram_descs.append(g)
        # mode product with the dense operator (symmetric, row stride D)
        op = self.dense[(D, k)][i].data_ptr()
        src = leaf.xb if a == 0 else leaf.y[(a - 1) % 2]
        dst = leaf.y[a % 2]
        lst = apply_descs.setdefault(a, [])
        if suf == 1 and r > 1:
          g = _lib.GemmDesc()
          g.a, g.b, g.c, g.c_in = src.data_ptr(), op, dst.data_ptr(), None
          g.a_iinner, g.a_sio, g.a_si = pre, 0, d
          g.a_kinner, g.a_sko, g.a_ski = d, 0, 1
          g.b_sj, g.b_kinner, g.b_sko, g.b_ski = D, d, 0, 1
          g.c_iinner, g.c_sio, g.c_sii = pre, 0, d
          g.m, g.n, g.k, g.alpha, g.beta = pre, d, d, 1.0, 0.0
          lst.append(g)
        else:
          for p in range(pre):
            g = _lib.GemmDesc()
            g.a, g.c_in = op, None
            g.b = src.data_ptr() + 4 * p * d * suf
            g.c = dst.data_ptr() + 4 * p * d * suf
            g.a_iinner, g.a_sio, g.a_si = d, 0, D
            g.a_kinner, g.a_sko, g.a_ski = d, 0, 1
            g.b_sj, g.b_kinner, g.b_sko, g.b_ski = 1, d, 0, suf
            g.c_iinner, g.c_sio, g.c_sii = d, 0, suf
            g.m, g.n, g.k, g.alpha, g.beta = d, suf, d, 1.0, 0.0
            lst.append(g)
    tc_ok = self.device is not None and bool(_lib.load().pc_device_supports_tcgen05())

    def lists(descs):
      tc = [g for g in descs if tc_ok and ops.tc_gemm_eligible(g)]
      simt = [g for g in descs if not (tc_ok and ops.tc_gemm_eligible(g))]
      return ([ops.TcGemmList(tc, self.device)] if tc else []) + \
             ([ops.SimtGemmLists(simt, self.device)] if simt else [])

    self.gram_lists = lists(gram_descs)
    self.apply_lists = [lists(apply_descs[a]) for a in sorted(apply_descs)]

  def step(self, updates, count: int, alias_outputs: bool = False):
    o = self.options
    is_leaf = lambda x: isinstance(x, _Leaf)

    def load(leaf, u):
      if not isinstance(u, torch.Tensor) or not u.is_cuda:
        raise RuntimeError("tearfree sketchy needs CUDA tensors: there is no CPU fallback")
      if not u.is_floating_point():
        raise TypeError(f"tearfree sketchy: updates must be floating point, got {u.dtype}")
      leaf.xb.copy_(u)
      return leaf

    _tree.tree_map(load, self.plan_tree, updates, is_leaf=is_leaf)
    if count % o.update_freq == 0:  # TF/sketchy.py:272-286
      for lst in self.gram_lists:
        lst.run()
      for key, packed in self.packed.items():
        D, k = key
        ops.fd_update_root_batched(
            self.gram[key], packed, self.ps[key], k, padding_starts=self.pads[key],
            decay=o.second_moment_decay, input_is_gram=True, out=self.scratch[key],
            tearfree_epsilon=o.epsilon, tearfree_relative_epsilon=o.relative_epsilon)
        packed.copy_(self.scratch[key])
        ops.low_rank_to_dense(packed, k, out=self.dense[key])
    for group in self.apply_lists:
      for lst in group:
        lst.run()

    def result(leaf):
      out = leaf.y[(len(leaf.shape) - 1) % 2]
      return out if alias_outputs else out.clone()

    return _tree.tree_map(result, self.plan_tree, is_leaf=is_leaf)


def apply(options: Options, _alias_outputs: bool = False
          ) -> praxis_shim.ShardedGradientTransformation:
  """Gradient transform for Sketchy preconditioning (TF/sketchy.py:65-75)."""
  _validate(options)
  holder = {}

  def init_fn(params) -> _SketchyState:
    eng = holder["engine"] = _Engine(options, params)
    return _SketchyState(count=torch.zeros([], dtype=torch.int32), sketches=eng.sketches_tree())

  def update_fn(updates, state: _SketchyState, params=None):
    del params
    eng = holder.get("engine")
    if eng is None:
      eng = holder["engine"] = _Engine(options, updates)
    if not eng.owns(state.sketches):
      eng.adopt(state.sketches)
    new_updates = eng.step(updates, int(state.count), _alias_outputs)
    return new_updates, _SketchyState(count=state.count + 1, sketches=eng.sketches_tree())

  return praxis_shim.ShardedGradientTransformation(init_fn, update_fn,
                                                   functools.partial(_pspec, options))


def _pspec(options: Options, params: praxis_shim.NestedHParams) -> praxis_shim.NestedHParams:
  """Sharding specification of the sketchy state: everything replicated (TF/sketchy.py:194-260)."""
  count_pspec = praxis_shim.WeightHParams(shape=[], init=None, dtype=torch.int32,
                                          collections=None, tensor_split_dims_mapping=[])

  def _replicated(shape):
    return praxis_shim.WeightHParams(shape=list(shape), init=None, dtype=torch.float32,
                                     collections=None,
                                     tensor_split_dims_mapping=[-1] * len(shape))

  def _tensor_pspec(path, param):
    def axis(d):
      k = min(d, options.rank)
      return dict(eigvecs=_replicated((d, k)), eigvals=_replicated((k,)),
                  inv_eigvals=_replicated((k,)), tail=_replicated(()), inv_tail=_replicated(()),
                  ema_ggt=praxis_shim.MaskedNode(), svd_result_u=praxis_shim.MaskedNode(),
                  svd_result_s=praxis_shim.MaskedNode(), inv_prev_tail=praxis_shim.MaskedNode())
    return dict(axes=[axis(d) for d in param.shape])

  return dict(count=count_pspec,
              sketches=_tree.tree_map_with_path(_tensor_pspec, params,
                                                is_leaf=lambda x: hasattr(x, "shape")))
