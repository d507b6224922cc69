"""numpy restatement of ``distributed_shampoo(...).init/update`` (TEST INFRASTRUCTURE).

Follows DS:1293-1321 (merge_small_dims), DS:1387-1437 (BlockPartitioner),
DS:1508-1708 (Preconditioner), DS:2585-2675 (init / statistics),
DS:2816-3010 and DS:3012-3281 (preconditioner computation, f32 and quantised,
including pad-to-max, contiguous device chunks and the failure fallback),
DS:3442-3659 (orchestration, grafting, momentum).  Params/grads are flat lists
(or dicts) of numpy arrays; state is plain Python objects.

Not restated (rejected with ValueError, as in the product): LOBPCG, eigh=True,
pjit/shard_optimizer_states, FD diagnostics.
"""
from __future__ import annotations

import enum
import itertools
from typing import Any, List, Optional

import numpy as np

from oracle import numerics as N


class GraftingType(enum.IntEnum):  # DS:499-506
  NONE = 0
  SGD = 1
  ADAGRAD = 2
  RMSPROP = 3
  RMSPROP_NORMALIZED = 4
  SQRT_N = 5
  ADAGRAD_NORMALIZED = 6


class PreconditionerType(enum.IntEnum):  # DS:509-517
  ALL = 1
  INPUT = 2
  OUTPUT = 3


def merge_small_dims(shape_to_merge, max_dim):
  """DS:1293-1321."""
  if shape_to_merge and np.all(np.array(shape_to_merge) == 1):
    return [1]
  out, product = [], 1
  for d in shape_to_merge:
    if product * d <= max_dim:
      product *= d
    else:
      if product > 1:
        out.append(product)
      product = d
  if product > 1:
    out.append(product)
  return out


class BlockPartitioner:
  """DS:1387-1437."""

  def __init__(self, shape, block_size):
    self._shape = tuple(shape)
    self._splits = []
    self._split_sizes = []
    for i, d in enumerate(self._shape):
      if 0 < block_size < d:
        nsplit = (d - 1) // block_size
        indices = (np.arange(nsplit, dtype=np.int32) + 1) * block_size
        sizes = np.ones(nsplit + 1, dtype=np.int32) * block_size
        sizes[-1] = d - indices[-1]
        self._splits.append((i, indices))
        self._split_sizes.append(sizes)
      else:
        self._split_sizes.append(np.array([d], dtype=np.int32))

  def split_sizes(self):
    return self._split_sizes

  def partition(self, tensor):
    assert tuple(tensor.shape) == self._shape
    tensors = [tensor]
    for i, indices in self._splits:
      nxt = []
      for t in tensors:
        nxt.extend(np.split(t, indices, axis=i))
      tensors = nxt
    return tensors

  def merge_partitions(self, partitions):
    for i, indices in reversed(self._splits):
      n = len(indices) + 1
      merged, ind = [], 0
      while ind < len(partitions):
        merged.append(np.concatenate(partitions[ind:ind + n], axis=i))
        ind += n
      partitions = merged
    assert len(partitions) == 1
    return partitions[0]


class Preconditioner:
  """DS:1508-1708."""

  def __init__(self, param_shape, block_size, merge_small_dims_block_size,
               best_effort_shape_interpretation,
               preconditioner_type=PreconditionerType.ALL, compression_rank=0):
    self._original_shape = tuple(param_shape)
    self._transformed_shape = tuple(param_shape)
    if best_effort_shape_interpretation:
      self._transformed_shape = tuple(
          merge_small_dims(self._original_shape, merge_small_dims_block_size))
    self._partitioner = BlockPartitioner(self._transformed_shape, block_size)
    self._preconditioner_type = preconditioner_type
    self._compression_rank = compression_rank

  def should_precondition_dims(self):
    rank = len(self._partitioner.split_sizes())
    t = self._preconditioner_type
    if t == PreconditionerType.ALL or rank <= 1:
      return [True] * rank
    if t == PreconditionerType.INPUT:
      return [True] * (rank - 1) + [False]
    return [False] * (rank - 1) + [True]

  def _preconditioner_shape(self, dim):
    dim = int(dim)
    if self._compression_rank:
      return [dim, N.precond_dim(self._compression_rank, dim)]
    return [dim, dim]

  def shapes_for_preconditioners(self):
    split_sizes = self._partitioner.split_sizes()
    rank = len(split_sizes)
    shapes = []
    for t in itertools.product(*split_sizes):
      if self._preconditioner_type == PreconditionerType.ALL or rank <= 1:
        shapes.extend(map(self._preconditioner_shape, t))
      elif self._preconditioner_type == PreconditionerType.INPUT:
        shapes.extend(map(self._preconditioner_shape, t[:-1]))
      else:
        shapes.extend(map(self._preconditioner_shape, t[-1:]))
    return shapes

  def exponent_for_preconditioner(self):
    return 2 * sum(self.should_precondition_dims())

  def updated_statistics_from_grad(self, stats, grad, w1, w2, to_float=None,
                                   from_float=None, frequent_directions=False):
    to_float = to_float or (lambda x: x)
    from_float = from_float or (lambda x: x)
    g_all = np.reshape(grad, self._transformed_shape)
    dims = [i for i, p in enumerate(self.should_precondition_dims()) if p]
    new_stats, index = [], 0
    for g in self._partitioner.partition(g_all):
      for axis in dims:
        update = N.gram_weighted_update
        if frequent_directions and N.should_compress(self._compression_rank,
                                                     g.shape[axis]):
          update = N.frequent_directions_update
        new_stats.append(from_float(update(to_float(stats[index]), g, axis, w1, w2)))
        index += 1
    return new_stats

  def _preconds_for_grad(self, preconditioners, rank, start, end):
    sel = preconditioners[start:end]
    if self._preconditioner_type == PreconditionerType.INPUT:
      sel = sel + [None]
    elif self._preconditioner_type == PreconditionerType.OUTPUT:
      sel = [None] * (rank - 1) + sel
    assert len(sel) == rank
    return sel

  def preconditioned_grad(self, grad, preconditioners):
    g_all = np.reshape(grad, self._transformed_shape)
    flags = self.should_precondition_dims()
    npre = sum(flags)
    out = []
    for i, g in enumerate(self._partitioner.partition(g_all)):
      ps = self._preconds_for_grad(preconditioners, len(flags), i * npre,
                                   (i + 1) * npre)
      out.append(self._precondition_block(g, flags, ps))
    return np.reshape(self._partitioner.merge_partitions(out), self._original_shape)

  def _precondition_block(self, g, flags, preconditioners):
    """DS:1676-1708."""
    for j, flag in enumerate(flags):
      rank = g.ndim
      roll = tuple(range(1, rank)) + (0,)
      if not flag:
        g = np.transpose(g, roll)
        continue
      dim, app_dim = preconditioners[j].shape
      if app_dim != dim:  # low-rank branch, DS:1690-1705
        vecs, vals, const, skip = N.low_rank_unpack(
            preconditioners[j], abs(self._compression_rank))
        basis = np.tensordot(g, vecs, axes=[[0], [0]])
        lowrank = np.tensordot(basis, vecs, axes=[[rank - 1], [1]])
        g = np.transpose(g, roll)
        complement = g - lowrank
        scaled = np.tensordot(basis * vals, vecs, axes=[[rank - 1], [1]])
        new_g = const * complement + scaled
        g = g if skip else new_g
        continue
      g = np.tensordot(g, preconditioners[j], axes=[[0], [0]])  # DS:1707
    return g


class ParameterStats:
  """DS:367-375."""

  def __init__(self, diagonal_statistics, statistics, preconditioners,
               diagonal_momentum, momentum, avg_grad, training_metrics):
    self.diagonal_statistics = diagonal_statistics
    self.statistics = statistics
    self.preconditioners = preconditioners
    self.diagonal_momentum = diagonal_momentum
    self.momentum = momentum
    self.avg_grad = avg_grad
    self.training_metrics = training_metrics  # [num_stats, 5] float32 or None


class ShampooState:
  """DS:488-490."""

  def __init__(self, count, stats):
    self.count = count
    self.stats = stats


def _batch_chunks(n_items, num_devices):
  """Contiguous chunks of DS:1827-1831 as index lists."""
  b = n_items // num_devices
  return [list(range(i, i + b)) for i in range(0, n_items, b)]


class _Shampoo:

  def __init__(self, learning_rate, block_size, beta1=0.9, beta2=0.999,
               diagonal_epsilon=1e-10, matrix_epsilon=1e-6, weight_decay=0.0,
               start_preconditioning_step=5, preconditioning_compute_steps=1,
               statistics_compute_steps=1, best_effort_shape_interpretation=True,
               graft_type=GraftingType.SGD, nesterov=True, exponent_override=0,
               batch_axis_name=None, num_devices=1,
               best_effort_memory_usage_reduction=False,
               inverse_failure_threshold=0.1, moving_average_for_momentum=False,
               skip_preconditioning_dim_size_gt=4096,
               clip_by_scaled_gradient_norm=None, relative_matrix_epsilon=True,
               merge_small_dims_block_size=4096,
               precondtioner_type=PreconditionerType.ALL, compression_rank=0,
               frequent_directions=False, reset_preconditioner=False,
               average_grad=False, skip_preconditioning_rank_lt=1,
               decoupled_learning_rate=True, decoupled_weight_decay=False,
               generate_training_metrics=True, reuse_preconditioner=False,
               stats_quantized_dtype=None, root_fn=None):
    self.reset_frequency = None
    if reset_preconditioner and not frequent_directions:  # DS:2019-2020
      raise ValueError("reset_preconditioner=True requries frequent_directions")
    if reset_preconditioner:  # DS:2022-2024
      self.reset_frequency = int(np.round(1 / (1 - beta2))) if beta2 != 1 else None
      beta2 = 1.0
    if frequent_directions and compression_rank <= 0:  # DS:2028-2030
      raise ValueError("frequent_directions=True requires compression_rank > 0,"
                       f" found {compression_rank}")
    if average_grad and not frequent_directions:  # DS:2032-2033
      raise ValueError("average_grad requested but frequent_directions is False")
    if frequent_directions and statistics_compute_steps != preconditioning_compute_steps:
      raise ValueError("frequent_directions=True requires statistics_compute_steps"
                       " to equal preconditioning_compute_steps")
    self.__dict__.update(locals())
    del self.__dict__["self"]
    self.beta2 = beta2
    # DS:2051-2054
    self.quantize_second_moment = bool(
        best_effort_memory_usage_reduction and not compression_rank and
        not frequent_directions and batch_axis_name)
    # DS:2056-2064: int16; ``stats_quantized_dtype`` lets tests exercise int8
    self.qdt_second = (stats_quantized_dtype or np.int16
                       ) if self.quantize_second_moment else np.float32
    self.root_fn = root_fn or N.matrix_inverse_pth_root

  # ---- helpers --------------------------------------------------------
  def _graft_has_diag(self):  # DS:2042-2045
    return self.graft_type not in (GraftingType.SGD, GraftingType.SQRT_N,
                                   GraftingType.NONE)

  def _momentum_dtype(self, var):  # DS:2047-2049
    return np.int8 if (self.best_effort_memory_usage_reduction and
                       var.ndim > 1) else np.float32

  def _quantize_momentum(self, m):  # DS:2111-2114
    return N.QuantizedValue.from_float_value(m, self._momentum_dtype(m))

  def _maybe_quantize_matrices(self, mats):  # DS:2087-2095
    if self.qdt_second != np.float32:
      return [N.QuantizedValue.from_float_value(s, self.qdt_second, True) for s in mats]
    return mats

  @staticmethod
  def _to_float(v):  # DS:2072-2076
    return v.to_float() if isinstance(v, N.QuantizedValue) else v

  def _preconditioner(self, param):  # DS:2116-2125
    return Preconditioner(param.shape, self.block_size,
                          self.merge_small_dims_block_size,
                          self.best_effort_shape_interpretation,
                          self.precondtioner_type, self.compression_rank)

  def _skip_preconditioning(self, param):  # DS:2627-2629
    return param.ndim < self.skip_preconditioning_rank_lt or any(
        s > self.skip_preconditioning_dim_size_gt for s in param.shape)

  # ---- init (DS:2585-2625) --------------------------------------------
  def init(self, params: List[np.ndarray]) -> ShampooState:
    stats = []
    for param in params:
      pre = self._preconditioner(param)
      statistics, preconditioners = [], []
      if not self._skip_preconditioning(param):
        shapes = pre.shapes_for_preconditioners()
        statistics = [np.float32(self.matrix_epsilon) * np.eye(s[0], dtype=np.float32)
                      for s in shapes]
        preconditioners = [np.eye(s[0], s[1], dtype=np.float32) * np.float32(s[0] == s[1])
                           for s in shapes]
      diag = np.zeros_like(param) if self._graft_has_diag() else []
      stats.append(ParameterStats(
          N.QuantizedValue.from_float_value(diag, np.float32),
          self._maybe_quantize_matrices(statistics),
          self._maybe_quantize_matrices(preconditioners),
          self._quantize_momentum(np.zeros_like(param)),
          self._quantize_momentum(np.zeros_like(param)),
          np.zeros_like(param) if (self.frequent_directions and self.average_grad) else None,
          np.zeros((len(statistics), 5), np.float32)
          if self.generate_training_metrics else None))
    return ShampooState(count=0, stats=stats)

  # ---- statistics (DS:2631-2675) --------------------------------------
  def _compute_stats(self, grad, state, param, step):
    pre = self._preconditioner(param)
    new_statistics = [[]] * len(state.statistics)
    w1 = self.beta2
    w2 = self.beta2 if self.beta2 == 1.0 else 1.0 - self.beta2
    new_avg_grad = None
    if not self._skip_preconditioning(param):
      if self.frequent_directions and self.average_grad:  # DS:2640-2645
        if self.statistics_compute_steps == 1 or step % self.statistics_compute_steps == 1:
          new_avg_grad = grad
        else:
          new_avg_grad = state.avg_grad + grad
        grad = new_avg_grad / np.float32(self.statistics_compute_steps)
      if self.statistics_compute_steps > 1 and step % self.statistics_compute_steps != 0:
        new_statistics = state.statistics
      else:
        new_statistics = pre.updated_statistics_from_grad(
            state.statistics, grad, w1, w2, to_float=self._to_float,
            from_float=lambda x: self._maybe_quantize_matrices([x])[0],
            frequent_directions=self.frequent_directions)
    return ParameterStats(state.diagonal_statistics, new_statistics,
                          state.preconditioners, state.diagonal_momentum,
                          state.momentum, new_avg_grad, state.training_metrics)

  # ---- one root with compression dispatch (DS:2688-2740) ---------------
  def _mi_pth_root(self, stat, exponent, padding_start, prev):
    kw = dict(ridge_epsilon=self.matrix_epsilon,
              relative_matrix_epsilon=self.relative_matrix_epsilon)
    if self.compression_rank != 0:
      if N.should_compress(self.compression_rank, padding_start):
        if self.frequent_directions:
          return N.fd_update_root(stat, exponent, rank=self.compression_rank,
                                  decay=self.beta2, padding_start=padding_start,
                                  prev=prev, **kw)
        return N.low_rank_root(stat, exponent,
                               compression_rank=self.compression_rank,
                               padding_start=padding_start, **kw)
      root, m = self.root_fn(stat, exponent, padding_start=padding_start, **kw)
      return root[:, :N.precond_dim(self.compression_rank, stat.shape[0])], m
    return self.root_fn(stat, exponent, padding_start=padding_start, **kw)

  # ---- preconditioners (DS:2816-3010, DS:3012-3281, DS:3442-3494) ------
  def _compute_preconditioners(self, states, params, step):
    statistics, num_per_state, original_shapes, exponents, prev_precs = [], [], [], [], []
    max_size = 0
    for state, param in zip(states, params):
      num_per_state.append(len(state.statistics))
      if state.statistics:
        pre = self._preconditioner(param)
        for s in state.statistics:
          exponents.append(pre.exponent_for_preconditioner()
                           if self.exponent_override == 0 else self.exponent_override)
          shp = tuple(s.shape) if isinstance(s, N.QuantizedValue) else s.shape
          original_shapes.append(shp)
          max_size = max(max_size, shp[0])
        statistics.extend(state.statistics)
        prev_precs.extend(state.preconditioners)
    if not statistics:
      return states
    quantized = self.qdt_second != np.float32
    num_devices = self.num_devices if self.batch_axis_name else 1
    num_statistics = len(statistics)
    perform_step = step % self.preconditioning_compute_steps == 0  # DS:2909

    def pd_of(size):
      return N.precond_dim(self.compression_rank, size)

    new_precs = list(prev_precs)
    metrics_rows = None
    if perform_step:
      # pad to max size and to a multiple of the device count (DS:2841-2850);
      # in the quantised path the *dequantised* padded region is zero
      # (DS:3046-3067), masked by padding_start either way.
      packed = []
      for s in statistics:
        if quantized:
          q = N.pad_square_matrix(s.quantized, max_size)
          dg = N.pad_vector(s.diagonal, max_size)
          bs = N.pad_vector(s.bucket_size, max_size)
          packed.append(N.dequantize(q, dg, bs, self.qdt_second, True))
        else:
          packed.append(N.pad_square_matrix(s, max_size))
      to_pad = -num_statistics % num_devices
      packed.extend([np.eye(max_size, dtype=np.float32)] * to_pad)
      exps = list(exponents) + [1] * to_pad
      paddings = [shp[0] for shp in original_shapes] + [0] * to_pad
      prevs = [None] * len(packed)
      if self.reuse_preconditioner:  # DS:2135-2160
        pd = pd_of(max_size)
        for i, pp in enumerate(prev_precs):
          pp = self._to_float(pp)
          if self.reset_frequency is not None:
            pp = pp * np.float32(0.0 if step % self.reset_frequency == 0 else 1.0)
          prevs[i] = np.pad(pp, ((0, max_size - pp.shape[0]), (0, pd - pp.shape[1])))
        for i in range(num_statistics, len(packed)):
          prevs[i] = np.zeros((max_size, pd), np.float32)
      roots = [None] * len(packed)
      rows = [None] * len(packed)
      # contiguous chunk r is what device r computes (DS:2862-2875); the
      # all-gather + unbatch (DS:2876-2879) restores flat order.
      for chunk in _batch_chunks(len(packed), num_devices):
        for i in chunk:
          r, m = self._mi_pth_root(packed[i], exps[i], paddings[i], prevs[i])
          if quantized:  # DS:2746-2772: requantise root inside the vmap
            qv = N.QuantizedValue.from_float_value(r, self.qdt_second, True)
            r = qv
          roots[i], rows[i] = r, m.as_row()
      metrics_rows = np.stack(rows)
      for i in range(num_statistics):  # DS:2936-2950 / DS:3197-3215
        err = metrics_rows[i, 0]
        if np.isnan(err) or err >= self.inverse_failure_threshold:
          continue
        shp = original_shapes[i]
        pdim = pd_of(shp[0])
        if quantized:
          qv = roots[i]
          new_precs[i] = N.QuantizedValue(
              qv.quantized[:shp[0], :pdim], qv.diagonal[:shp[0]],
              qv.bucket_size[:pdim], self.qdt_second, True, [shp[0], pdim])
        else:
          new_precs[i] = roots[i][:shp[0], :pdim]

    new_states, idx = [], 0
    for n_s, state in zip(num_per_state, states):
      if n_s == 0:
        precs = []
        tm = np.zeros((0, 5), np.float32) if self.generate_training_metrics else None
      else:
        precs = new_precs[idx:idx + n_s]
        tm = None
        if self.generate_training_metrics:
          tm = metrics_rows[idx:idx + n_s] if perform_step else state.training_metrics
        idx += n_s
      new_states.append(ParameterStats(
          state.diagonal_statistics, state.statistics, precs,
          state.diagonal_momentum, state.momentum, state.avg_grad, tm))
    return new_states

  # ---- grafting + momentum (DS:3496-3625) ------------------------------
  def _transform_grad(self, grad, state, param, step):
    f = np.float32
    pre = self._preconditioner(param)
    gt = self.graft_type
    new_diag = state.diagonal_statistics.to_float()
    w1 = f(self.beta2)
    w2 = f(self.beta2 if self.beta2 == 1.0 else 1.0 - self.beta2)
    if gt in (GraftingType.ADAGRAD, GraftingType.ADAGRAD_NORMALIZED):
      sg = grad
      if gt == GraftingType.ADAGRAD_NORMALIZED:
        sg = grad / (np.linalg.norm(grad) + f(N._EPSILON))
      new_diag = new_diag + np.square(sg)
      graft = sg / (np.sqrt(new_diag) + f(self.diagonal_epsilon))
    elif gt in (GraftingType.RMSPROP, GraftingType.RMSPROP_NORMALIZED):
      sg = grad
      if gt == GraftingType.RMSPROP_NORMALIZED:
        sg = grad / (np.linalg.norm(grad) + f(N._EPSILON))
      new_diag = w1 * new_diag + w2 * np.square(sg)
      graft = sg / (np.sqrt(new_diag) + f(self.diagonal_epsilon))
      if self.clip_by_scaled_gradient_norm:
        sgn = np.linalg.norm(graft) / f(np.sqrt(float(graft.size)))
        graft = graft / np.maximum(f(1.0), sgn / f(self.clip_by_scaled_gradient_norm))
    elif gt in (GraftingType.SGD, GraftingType.NONE):
      graft = grad
    else:
      graft = np.ones_like(grad) * np.sign(grad)
    lr = self.learning_rate(step) if callable(self.learning_rate) else self.learning_rate
    lr = f(lr)
    graft = graft * (lr if not self.decoupled_learning_rate else f(1.0))
    if not self._skip_preconditioning(param):
      precond_grad = pre.preconditioned_grad(
          grad, [self._to_float(p) for p in state.preconditioners])
    else:
      precond_grad = graft
    gnorm = np.linalg.norm(graft)
    pnorm = np.linalg.norm(precond_grad)
    mult = gnorm / (pnorm + f(N._EPSILON)) if gt is not GraftingType.NONE else f(1.0)
    shampoo = precond_grad * mult
    shampoo_wd, graft_wd = shampoo, graft
    wd = f(self.weight_decay)
    if self.weight_decay != 0 and not self.decoupled_weight_decay:
      shampoo_wd = shampoo + wd * param
      graft_wd = graft + wd * param
    beta1 = f(self.beta1)
    w = f(1.0 - self.beta1) if self.moving_average_for_momentum else f(1.0)
    shampoo_m = state.momentum.to_float() * beta1 + w * shampoo_wd
    graft_m = state.diagonal_momentum.to_float() * beta1 + w * graft_wd
    run = f(step >= self.start_preconditioning_step)
    mom = run * shampoo_m + (f(1.0) - run) * graft_m
    wdu = run * shampoo_wd + (f(1.0) - run) * graft_wd
    nest = w * wdu + beta1 * mom if self.nesterov else mom
    if self.weight_decay != 0 and self.decoupled_weight_decay:
      wd_lr = f(1.0) if self.decoupled_learning_rate else lr
      nest = nest + wd_lr * wd * param
    out = f(-1.0) * (lr if self.decoupled_learning_rate else f(1.0)) * nest
    new_state = ParameterStats(
        N.QuantizedValue.from_float_value(new_diag, np.float32), state.statistics,
        state.preconditioners, self._quantize_momentum(graft_m),
        self._quantize_momentum(shampoo_m), state.avg_grad, state.training_metrics)
    return out.astype(np.float32), new_state

  # ---- update (DS:3627-3659) -------------------------------------------
  def update(self, grads, state: ShampooState, params):
    step = state.count
    stats = [self._compute_stats(g, s, p, step)
             for g, s, p in zip(grads, state.stats, params)]
    stats = self._compute_preconditioners(stats, params, step)
    outs = [self._transform_grad(g, s, p, step) for g, s, p in zip(grads, stats, params)]
    updates = [o[0] for o in outs]
    return updates, ShampooState(step + 1, [o[1] for o in outs])


def distributed_shampoo(learning_rate, block_size, **kw) -> Any:
  """Oracle twin of DS:1849-3675 (numpy; flat lists of arrays)."""
  for bad in ("lobpcg_topk_precondition", "eigh", "shard_optimizer_states"):
    if kw.pop(bad, 0):
      raise ValueError(f"{bad} is outside the hot path and not restated")
  return _Shampoo(learning_rate, block_size, **kw)
