"""Host-side mirror of ``precondition.distributed_shampoo`` for the B200 hot path.

Same public surface as the reference (DS = precondition/distributed_shampoo.py):
``distributed_shampoo(learning_rate, block_size, **kwargs)`` (DS:1849-1900) returns
a ``GradientTransformation(init, update)``; ``GraftingType``, ``PreconditionerType``,
``merge_small_dims``, ``BlockPartitioner``, ``Preconditioner``,
``pad_square_matrix``, ``matrix_inverse_pth_root``, ``power_iteration`` and
``mat_power`` are importable by name.  Parameters / gradients are pytrees (nested
list / tuple / dict) of CUDA ``torch.Tensor``; all device work goes through the C
ABI in ``include/precond_b200.h`` -- there is no CPU or PyTorch compute fallback.

B200-first layout decisions:
  * statistics / preconditioners of equal size live stacked in one bucket tensor
    ``[N_s, s, s]`` -- the root solver consumes a bucket as one batch with no
    pad-to-max copies (the reference pads every statistic to the largest block,
    DS:2841-2850; the masked maths make both equal);
  * gradient blocks are never materialised: every mode-k unfolding of a block is a
    strided view handed to the grouped GEMM (DS:1412-1422 uses jnp.split copies);
  * the per-step work is a fixed list of launches built once in ``init``.

``shard_optimizer_states=True`` (the pjit path, DS:2162-2583) is the stacked layout: one
padded-to-max array of statistics whose leading axis is sharded over the ranks (each rank
stores and updates only its rows), roots computed where the rows live, preconditioners
all-gathered.  Not built (raise on use): FD diagnostics (``generate_fd_metrics``), merged
shapes of rank > 3.
"""
from __future__ import annotations

import ctypes
import enum
import os
import itertools
from typing import Any, Callable, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np
import torch

from precondition_b200 import _lib, ops, peer
from precondition_b200.quantization_utils import QuantizedValue


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


class GradientTransformation(NamedTuple):
  """optax.GradientTransformation stand-in (DS:3675)."""
  init: Callable
  update: Callable


class ShampooState(NamedTuple):  # DS:488-490
  count: int
  stats: Any


# ---------------------------------------------------------------------------
# shape logic (host only) -- DS:1293-1321, DS:1387-1437, DS:1508-1643
# ---------------------------------------------------------------------------
def merge_small_dims(shape_to_merge, max_dim):
  """Merge small dimensions, e.g. [1, 2, 512, 1, 2048, 1, 3, 4] --> [1024, 2048, 12]
  if max_dim = 1024 (DS:1293-1321)."""
  if shape_to_merge and np.all(np.array(shape_to_merge) == 1):
    return [1]
  resulting_shape, product = [], 1
  for d in shape_to_merge:
    if product * d <= max_dim:
      product *= d
    else:
      if product > 1:
        resulting_shape.append(product)
      product = d
  if product > 1:
    resulting_shape.append(product)
  return resulting_shape


def pad_square_matrix(mat: torch.Tensor, max_size: int) -> torch.Tensor:
  """Given M returns [[M, 0], [0, I]] (DS:1324-1350)."""
  rows, cols = mat.shape
  if rows != cols:
    raise ValueError(f"Must have rows == cols, instead got rows={rows}, cols={cols}")
  if cols > max_size:
    raise ValueError(
        f"Must have cols <= max_size. Instead got cols={cols}, max_size={max_size}.")
  if rows == max_size:
    return mat
  out = torch.eye(max_size, dtype=mat.dtype, device=mat.device)
  out[:rows, :rows] = mat
  return out


class BlockPartitioner:
  """Block metadata of DS:1387-1437.  ``partition`` / ``merge_partitions`` exist for
  API parity (views / copies for inspection); the kernels read blocks in place."""

  def __init__(self, param_or_shape, block_size):
    shape = tuple(getattr(param_or_shape, "shape", param_or_shape))
    self._shape = shape
    self._splits, self._split_sizes = [], []
    for i, d in enumerate(shape):
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

  def block_offsets(self):
    """Row-major list of (offsets, sizes) per block -- itertools.product order of
    DS:1630 == the order ``partition`` yields."""
    per_axis = []
    for sizes in self._split_sizes:
      offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
      per_axis.append(list(zip(offs.tolist(), [int(s) for s in sizes])))
    return [(tuple(o for o, _ in combo), tuple(s for _, s in combo))
            for combo in itertools.product(*per_axis)]

  def partition(self, tensor):
    assert tuple(tensor.shape) == self._shape
    return [tensor[tuple(slice(o, o + s) for o, s in zip(offs, sizes))]
            for offs, sizes in self.block_offsets()]

  def merge_partitions(self, partitions):
    out = torch.empty(self._shape, dtype=partitions[0].dtype, device=partitions[0].device)
    for (offs, sizes), p in zip(self.block_offsets(), partitions):
      out[tuple(slice(o, o + s) for o, s in zip(offs, sizes))] = p
    return out


def _precond_dim(compression_rank, dim):  # DS:520-532
  if not compression_rank:
    return dim
  compressed = abs(compression_rank) + 2
  return dim if compressed >= dim else compressed


class Preconditioner:
  """Shape / exponent metadata of DS:1508-1643 (statistics and application run in
  the library, see ``_Shampoo``)."""

  def __init__(self, param, block_size, merge_small_dims_block_size,
               best_effort_shape_interpretation,
               preconditioner_type=PreconditionerType.ALL, compression_rank=0):
    self._original_shape = tuple(getattr(param, "shape", param))
    self._transformed_shape = self._original_shape
    if best_effort_shape_interpretation:
      self._transformed_shape = tuple(
          merge_small_dims(self._original_shape, merge_small_dims_block_size))
    self._partitioner = BlockPartitioner(self._transformed_shape, block_size)
    self._preconditioner_type = preconditioner_type
    self._compression_rank = compression_rank

  def should_precondition_dims(self):
    rank = len(self._partitioner.split_sizes())
    if self._preconditioner_type == PreconditionerType.ALL or rank <= 1:
      return [True] * rank
    if self._preconditioner_type == PreconditionerType.INPUT:
      return [True] * (rank - 1) + [False]
    return [False] * (rank - 1) + [True]

  def _preconditioner_shape(self, dim):
    dim = int(dim)
    if self._compression_rank:
      return [dim, _precond_dim(self._compression_rank, dim)]
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


# ---------------------------------------------------------------------------
# stand-alone numerical entry points with the reference's signatures
# ---------------------------------------------------------------------------
def matrix_inverse_pth_root(matrix, p, num_iters=100, ridge_epsilon=1e-6,
                            error_tolerance=1e-6, precision=None,
                            relative_matrix_epsilon=True, lobpcg_topk_precondition=0,
                            lobpcg_max_iter=0, padding_start=None, prev=None, eigh=False):
  """DS:702-940 on one matrix -> (root, metrics row [5])."""
  del precision, prev
  if lobpcg_topk_precondition:  # DS:789-812, DS:889-928
    roots, metrics, _ = ops.matrix_inverse_pth_root_lobpcg_batched(
        matrix[None].contiguous(), [int(p)], int(lobpcg_topk_precondition),
        None if padding_start is None else [int(padding_start)], ridge_epsilon=ridge_epsilon,
        error_tolerance=error_tolerance, num_iters=num_iters,
        relative_matrix_epsilon=relative_matrix_epsilon, lobpcg_max_iter=lobpcg_max_iter,
        diagnostics=False)
    return roots[0], metrics[0]
  if eigh:  # matrix_inverse_pth_root_eigh, DS:943-1030
    roots, metrics = ops.matrix_inverse_pth_root_eigh_batched(
        matrix[None].contiguous(), [int(p)],
        None if padding_start is None else [int(padding_start)],
        ridge_epsilon=ridge_epsilon, error_tolerance=error_tolerance,
        relative_matrix_epsilon=relative_matrix_epsilon)
    return roots[0], metrics[0]
  roots, metrics = ops.matrix_inverse_pth_root_batched(
      matrix[None].contiguous(), [int(p)],
      None if padding_start is None else [int(padding_start)],
      ridge_epsilon=ridge_epsilon, error_tolerance=error_tolerance, num_iters=num_iters,
      relative_matrix_epsilon=relative_matrix_epsilon)
  return roots[0], metrics[0]


def power_iteration(matrix, num_iters=100, error_tolerance=1e-6, precision=None,
                    padding_start=None):
  """DS:595-652 -> (None, lambda_max); the eigenvector is not materialised (the
  solver only consumes the eigenvalue, DS:820)."""
  del precision
  lam, _ = ops.power_iteration(matrix[None].contiguous(),
                               None if padding_start is None else [int(padding_start)],
                               num_iters, error_tolerance)
  return None, lam[0]


def _matmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
  """C = A @ B (fp32, CUDA cores) through the grouped-GEMM entry point."""
  m, k = a.shape
  k2, n = b.shape
  assert k == k2
  a, b = a.contiguous(), b.contiguous()
  c = torch.empty((m, n), dtype=torch.float32, device=a.device)
  d = _lib.GemmDesc()
  d.a, d.b, d.c, d.c_in = a.data_ptr(), b.data_ptr(), c.data_ptr(), None
  d.a_iinner, d.a_sio, d.a_si = m, 0, k
  d.a_kinner, d.a_sko, d.a_ski = k, 0, 1
  d.b_sj, d.b_kinner, d.b_sko, d.b_ski = 1, k, 0, n
  d.c_iinner, d.c_sio, d.c_sii = m, 0, n
  d.m, d.n, d.k, d.alpha, d.beta = m, n, k, 1.0, 0.0
  dev = ops.upload_gemm_descs([d], a.device)
  ops.grouped_gemm(dev, 1, m, n)
  return c


def mat_power(mat_m: torch.Tensor, p: int, precision=None) -> torch.Tensor:
  """M^p with the multiply order of DS:655-678 (no-op products skipped)."""
  del precision
  power, mat, i = None, mat_m, int(p)
  while i > 0:
    if i % 2 == 1:
      power = mat if power is None else _matmul(mat, power)
    i //= 2
    if i > 0:
      mat = _matmul(mat, mat)
  if power is None:
    power = torch.eye(mat_m.shape[0], dtype=mat_m.dtype, device=mat_m.device)
  return power


# ---------------------------------------------------------------------------
# pytrees (list / tuple / dict of tensors)
# ---------------------------------------------------------------------------
def _tree_flatten(tree):
  if isinstance(tree, torch.Tensor):
    return [tree], None
  if isinstance(tree, dict):
    keys = sorted(tree.keys())
    parts = [_tree_flatten(tree[k]) for k in keys]
    return [x for p, _ in parts for x in p], ("dict", keys, [d for _, d in parts],
                                               [len(p) for p, _ in parts])
  if isinstance(tree, (list, tuple)):
    parts = [_tree_flatten(v) for v in tree]
    kind = "namedtuple" if hasattr(tree, "_fields") else type(tree).__name__
    return [x for p, _ in parts for x in p], (kind, type(tree), [d for _, d in parts],
                                               [len(p) for p, _ in parts])
  raise TypeError(f"unsupported pytree node {type(tree)}")


def _tree_flatten_any(tree, leaf_type):
  """Leaves of `tree` that are instances of `leaf_type` (dict / list / tuple containers)."""
  if isinstance(tree, leaf_type):
    return [tree]
  if isinstance(tree, dict):
    return [x for k in sorted(tree.keys()) for x in _tree_flatten_any(tree[k], leaf_type)]
  if isinstance(tree, (list, tuple)):
    return [x for v in tree for x in _tree_flatten_any(v, leaf_type)]
  return []


def _tree_unflatten(treedef, leaves):
  if treedef is None:
    return leaves[0]
  kind, meta, subdefs, counts = treedef
  out, i = [], 0
  for d, c in zip(subdefs, counts):
    out.append(_tree_unflatten(d, leaves[i:i + c]))
    i += c
  if kind == "dict":
    return dict(zip(meta, out))
  if kind == "namedtuple":
    return meta(*out)
  return meta(out)


class GlobalShardedParameterStats(NamedTuple):  # DS:382-386
  """Sharded mode: one stacked, padded-to-max array per kind; the leading axis is sharded over
  the devices.  In this implementation ``statistics`` holds THIS rank's rows
  ``[N / D, max, max]`` (row ``i`` is global row ``rank * N / D + i``), ``preconditioners`` all
  ``[N, max, max]`` rows (every rank applies all of them), ``exponents`` all ``[N]``."""
  statistics: Any
  preconditioners: Any
  exponents: Any


class LocalShardedParameterStats(NamedTuple):  # DS:391-402
  diagonal_statistics: Any
  diagonal_momentum: Any
  momentum: Any
  avg_grad: Any
  training_metrics: Any
  index_start: int  # first row of this parameter in the global arrays
  sizes: Any        # true sizes of its statistics


class ShardedShampooStats(NamedTuple):  # DS:482-485
  global_stats: Any
  local_stats: Any


class InitFnState(NamedTuple):  # DS:493-496
  init_fn: Any
  pspec_fn: Any
  shape_and_dtype_fn: Any


class InversePthRootDiagnostics(NamedTuple):
  """DS:109-142: entrywise errors between B^p A and I for an inverse p-th root B."""
  max_diag_error: Any = 0.0
  avg_diag_error: Any = 0.0
  max_off_diag_error: Any = 0.0
  avg_off_diag_error: Any = 0.0
  p: Any = 0.0

  @classmethod
  def create(cls, pth_inverse_root, matrix, p):
    row = ops.root_diagnostics(pth_inverse_root[None].contiguous(), matrix[None].contiguous(),
                               [int(p)])[0]
    return cls(*[row[i] for i in range(5)])


class LOBPCGDiagnostics(NamedTuple):
  """DS:149-195: consistency |A v - l v| / (l + |A v|) and orthogonality of top-k eigenpairs."""
  lobpcg_iters: Any = 0.0
  max_consistency_error: Any = 0.0
  avg_consistency_error: Any = 0.0
  avg_orthogonality_error: Any = 0.0
  max_eigenvalue: Any = 0.0
  min_eigenvalue: Any = 0.0
  num_topk_eigenvectors: Any = 0.0


# ---------------------------------------------------------------------------
# preconditioning_compute_steps schedule (DS:44-76)
# ---------------------------------------------------------------------------
def preconditioning_compute_steps_schedule(lr_fn, start_preconditioning_compute_steps,
                                           end_preconditioning_compute_steps, step):
  """Grows preconditioning_compute_steps along the learning-rate schedule, from the start
  value to start + end as the rate decays to 0, rounded down to a multiple of 10 (DS:44-76)."""
  base_lr = float(lr_fn(0))
  lr = float(lr_fn(step))
  decay_factor = lr / base_lr
  t = start_preconditioning_compute_steps + (1 - decay_factor) * end_preconditioning_compute_steps
  return max((t // 10) * 10, 1)


def newton_gemms_per_iteration(p: int) -> int:
  """G(p) of SURVEY 8(d): necessary products per coupled-Newton iteration."""
  p = int(p)
  return (p.bit_length() - 1) + bin(p).count("1") - 1 + 2


def partition_statistics(buckets, world):
  """Cost-balanced partition of all statistics over ``world`` ranks.

  ``buckets``: list of (key, [cost of statistic 0, 1, ...]).  Buckets are taken in order of
  decreasing per-statistic cost; inside a bucket the statistics, sorted by decreasing cost, are
  dealt round-robin over the ranks visited from the least loaded one upwards, so that per
  bucket the counts differ by at most one (equal-size all-gather payloads, DS:2844-2850
  fillers only for the remainder) while the total cost per rank stays balanced.  Every rank
  computes the same table.  Returns {key: [sorted indices of rank 0, of rank 1, ...]}."""
  load = [0.0] * world
  out = {}
  for key, costs in sorted(buckets, key=lambda kv: (-max(kv[1], default=0.0), str(kv[0]))):
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    rank_order = sorted(range(world), key=lambda r: (load[r], r))
    ranks = [[] for _ in range(world)]
    for k, i in enumerate(order):
      r = rank_order[k % world]
      ranks[r].append(i)
      load[r] += costs[i]
    out[key] = [sorted(r) for r in ranks]
  return out


def gather_layout(owned, cnt, sections):
  """Byte layout of one bucket's packed all-gather and the index arrays of the scatter back
  into state order (``pc_select_scatter``).

  ``owned[r]``: sorted state indices rank r solves; ``cnt``: rows per rank (the shorter ranks
  are padded with fillers, DS:2844-2850); ``sections``: [(name, row_bytes)] with "m" = the
  metrics rows (5 floats).  Every rank sends ``lbytes`` bytes: the sections back to back, each
  16-byte aligned.  Gathered row j = (rank j // cnt, local row j % cnt)."""
  world = len(owned)
  off, sec = 0, {}
  for name, row in sections:
    sec[name] = (off, row)
    off += _align16(cnt * row)
  dst = np.full(world * cnt, -1, dtype=np.int32)
  for r, o in enumerate(owned):
    dst[r * cnt:r * cnt + len(o)] = o
  rr = np.repeat(np.arange(world, dtype=np.int64), cnt)
  ll = np.tile(np.arange(cnt, dtype=np.int64), world)
  src = {name: rr * off + o + ll * row for name, (o, row) in sec.items() if name != "m"}
  return {"lbytes": off, "sections": sec, "dst_index": dst,
          "metrics_offset": (rr * off + sec["m"][0]) // 4 + ll * 5, "src_offset": src}


# ---------------------------------------------------------------------------
# state
# ---------------------------------------------------------------------------
class ParameterStats:
  """Per-parameter optimizer state, field-for-field DS:367-375.  The tensors are VIEWS of
  the optimizer's stacked / flat device buffers (see ``_Shampoo``): ``update`` advances them
  in place.  ``export_state`` / ``import_state`` of the transformation copy them out / in."""

  def __init__(self, diagonal_statistics, statistics, preconditioners, diagonal_momentum,
               momentum, avg_grad, metric_rows, device):
    self.diagonal_statistics = diagonal_statistics
    self.statistics = statistics
    self.preconditioners = preconditioners
    self.diagonal_momentum = diagonal_momentum
    self.momentum = momentum
    self.avg_grad = avg_grad
    self._metric_rows = metric_rows  # [5] rows of the bucket metrics, or None
    self._device = device

  @property
  def training_metrics(self):
    """[num_statistics, 5] rows of TrainingMetrics scalars (DS:338-351)."""
    if self._metric_rows is None:
      return None
    if not self._metric_rows:
      return torch.zeros((0, 5), dtype=torch.float32, device=self._device)
    return torch.stack(list(self._metric_rows))


def _state_tensors(st: ParameterStats):
  """Every tensor of a ParameterStats in a fixed order (export / import / aliasing checks)."""
  out = []

  def add(x):
    if isinstance(x, QuantizedValue):
      for t in (x.quantized, x.diagonal, x.bucket_size):
        if isinstance(t, torch.Tensor):
          out.append(t)
    elif isinstance(x, torch.Tensor):
      out.append(x)

  add(st.diagonal_statistics)
  for x in st.statistics:
    add(x)
  for x in st.preconditioners:
    add(x)
  add(st.diagonal_momentum)
  add(st.momentum)
  add(st.avg_grad)
  if st._metric_rows:
    out.extend(st._metric_rows)
  return out


def _clone_stats(st: ParameterStats) -> ParameterStats:
  def cl(x):
    if isinstance(x, QuantizedValue):
      return QuantizedValue(cl(x.quantized), cl(x.diagonal), cl(x.bucket_size), x.quantized_dtype,
                            x.extract_diagonal, list(x.shape))
    if isinstance(x, torch.Tensor):
      return x.detach().clone()
    return x
  return ParameterStats(cl(st.diagonal_statistics), [cl(x) for x in st.statistics],
                        [cl(x) for x in st.preconditioners], cl(st.diagonal_momentum),
                        cl(st.momentum), cl(st.avg_grad),
                        None if st._metric_rows is None else [cl(x) for x in st._metric_rows],
                        st._device)


class _Bucket:
  """All statistics of one size: stacked storage, one root-solver batch."""

  def __init__(self, size, pdim):
    self.size, self.pdim = size, pdim
    self.exponents: List[int] = []
    self.true_sizes: List[int] = []  # == size except in the stacked (sharded-state) layout
    self.count = 0
    self.job = None


class _ParamPlan:
  pass


class _RootJob:
  """Static buffers of one bucket's (possibly sharded) root solve."""
  pass


def _align16(x):
  return (x + 15) // 16 * 16


class _Shampoo:

  def __init__(self, learning_rate, block_size, beta1, beta2, diagonal_epsilon, matrix_epsilon,
               weight_decay, start_preconditioning_step, preconditioning_compute_steps,
               statistics_compute_steps, best_effort_shape_interpretation, graft_type, nesterov,
               exponent_override, batch_axis_name, best_effort_memory_usage_reduction,
               inverse_failure_threshold, moving_average_for_momentum,
               skip_preconditioning_dim_size_gt, clip_by_scaled_gradient_norm,
               relative_matrix_epsilon, merge_small_dims_block_size, precondtioner_type,
               compression_rank, skip_preconditioning_rank_lt, decoupled_learning_rate,
               decoupled_weight_decay, generate_training_metrics, engine, process_group,
               frequent_directions=False, reuse_preconditioner=False, reset_frequency=None,
               average_grad=False, eigh=False, decay_preconditioning_compute_steps=False,
               end_preconditioning_compute_steps=None, lobpcg_topk_precondition=0,
               lobpcg_max_iter=0, stacked=False, num_devices_for_pjit=None):
    self.__dict__.update({k: v for k, v in locals().items() if k != "self"})
    # DS:2051-2064: second-moment quantisation only with a batch axis
    self.quantize_second_moment = bool(best_effort_memory_usage_reduction and
                                       not compression_rank and batch_axis_name)
    self.qdt_second = torch.int16 if self.quantize_second_moment else torch.float32
    self.device = None
    self.plans: List[_ParamPlan] = []
    self.buckets = {}
    self.metrics = {}
    self._built = False

  # ---- helpers ---------------------------------------------------------
  def _graft_has_diag(self):  # DS:2042-2045
    return self.graft_type not in (GraftingType.SGD, GraftingType.SQRT_N, GraftingType.NONE)

  def _momentum_dtype(self, p):  # DS:2047-2049
    return torch.int8 if (self.best_effort_memory_usage_reduction and p.dim() > 1) \
        else torch.float32

  def _skip_preconditioning(self, shape):  # DS:2627-2629
    return len(shape) < self.skip_preconditioning_rank_lt or any(
        s > self.skip_preconditioning_dim_size_gt for s in shape)

  def _world(self):
    if not self.batch_axis_name:
      return 1, 0
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
      return dist.get_world_size(self.process_group), dist.get_rank(self.process_group)
    return 1, 0

  # ---- init (DS:2585-2625) ----------------------------------------------
  def init(self, params):
    leaves, treedef = _tree_flatten(params)
    self.treedef = treedef
    if not leaves:
      return ShampooState(0, _tree_unflatten(treedef, []))
    self.device = leaves[0].device
    if self.device.type != "cuda":
      raise RuntimeError("precondition_b200 needs CUDA tensors: there is no CPU fallback")
    for p in leaves:
      if not p.is_floating_point():
        raise TypeError(f"parameters must be floating point, got {p.dtype}")
      if p.device != self.device:
        raise RuntimeError("all parameters must live on one CUDA device")
    dev = self.device
    # Flat fp32 buffers, one 128-byte aligned segment per parameter at the same offset in each:
    # gradient, preconditioned gradient, GEMM temporaries, momenta, diagonal statistics.  The
    # optimizer state is float32 whatever the parameter dtype / memory layout is (the kernels
    # read and write 4-byte elements of contiguous row-major segments).
    offsets, total = [], 0
    for p in leaves:
      offsets.append(total)
      total += (p.numel() + 31) // 32 * 32
    self.total = total
    zeros = lambda: torch.zeros(total, dtype=torch.float32, device=dev)
    self.gbuf, self.pgbuf, self.t1buf, self.t2buf = zeros(), zeros(), zeros(), zeros()
    self.mbuf, self.dmbuf = zeros(), zeros()
    self.dsbuf = zeros() if self._graft_has_diag() else None
    self.pbuf = zeros() if self.weight_decay != 0 else None
    # Sketchy with average_grad (DS:2640-2645): statistics are taken from the running
    # gradient sum divided by statistics_compute_steps instead of the raw gradient
    self.use_avg_grad = bool(self.frequent_directions and self.average_grad)
    self.agbuf = zeros() if self.use_avg_grad else None
    self.sgbuf = zeros() if self.use_avg_grad else None

    self.plans, self.buckets = [], {}
    max_size = 0
    if self.stacked:  # DS:2169-2176
      if self.compression_rank or self.best_effort_memory_usage_reduction:
        raise NotImplementedError("shard_optimizer_states: quantised / compressed states are not "
                                  "built in the B200 path")
      for p in leaves:
        if not self._skip_preconditioning(p.shape):
          pre = Preconditioner(p.shape, self.block_size, self.merge_small_dims_block_size,
                               self.best_effort_shape_interpretation, self.precondtioner_type, 0)
          max_size = max([max_size] + [sh[0] for sh in pre.shapes_for_preconditioners()])
    for idx, p in enumerate(leaves):
      plan = _ParamPlan()
      plan.index, plan.offset, plan.shape, plan.numel = idx, offsets[idx], tuple(p.shape), p.numel()
      plan.pre = Preconditioner(p.shape, self.block_size, self.merge_small_dims_block_size,
                                self.best_effort_shape_interpretation, self.precondtioner_type,
                                self.compression_rank)
      plan.skip = self._skip_preconditioning(p.shape)
      plan.tshape = plan.pre._transformed_shape
      plan.flags = plan.pre.should_precondition_dims()
      plan.exponent = (plan.pre.exponent_for_preconditioner()
                       if self.exponent_override == 0 else self.exponent_override)
      plan.mdt = self._momentum_dtype(p)
      plan.stat_refs = []  # (bucket size, index in bucket) in reference order
      plan.stat_sizes = []
      plan.blocks = []
      if not plan.skip:
        if any(plan.flags) and not 1 <= plan.exponent <= 16:
          raise ValueError(f"inverse-root exponent {plan.exponent} outside [1, 16] "
                           "(exponent_override or tensor rank too large for the Newton programs)")
        for offs, sizes in plan.pre._partitioner.block_offsets():
          refs = []
          for axis, flag in enumerate(plan.flags):
            if not flag:
              refs.append(None)
              continue
            s = sizes[axis]
            # sharded-state mode (DS:2162-2240): ONE stacked array, every statistic padded to
            # the largest one; otherwise one bucket per true size
            key = max_size if self.stacked else s
            bk = self.buckets.setdefault(key, _Bucket(key, _precond_dim(self.compression_rank, key)))
            refs.append((key, bk.count))
            plan.stat_refs.append((key, bk.count))
            plan.stat_sizes.append(s)
            bk.exponents.append(plan.exponent)
            bk.true_sizes.append(s)
            bk.count += 1
          plan.blocks.append((offs, sizes, refs))
      self.plans.append(plan)

    # bucket storage: statistics = matrix_epsilon * I, preconditioners = I (DS:2594-2602)
    world, rank = self._world()
    for s, bk in self.buckets.items():
      eye = torch.eye(s, dtype=torch.float32, device=dev)
      bk.lo, bk.n_local = 0, bk.count
      if self.stacked:
        # DS:2229-2256: pad the stack to a multiple of the device count with (I, exponent 1)
        # fillers; rank r keeps the statistics rows [r N / D, (r + 1) N / D)
        ndev = int(self.num_devices_for_pjit or world)
        if world > 1 and ndev != world:
          raise ValueError(f"num_devices_for_pjit={ndev} != world size {world}")
        fill = -bk.count % ndev
        bk.exponents += [1] * fill
        bk.true_sizes += [0] * fill  # padding_start 0: zero root (DS:930-937)
        bk.count += fill
        bk.n_local = bk.count // world
        bk.lo = rank * bk.n_local
      bk.stats = (self.matrix_epsilon * eye).repeat(bk.n_local, 1, 1).contiguous()
      bk.compressed = bk.pdim != s  # DS:535-537: low-rank [s, rank + 2] preconditioners
      if bk.compressed:
        # packed sketches start at zero (DS:2598-2602); `precs` holds the dense operator the
        # packed form applies (pc_low_rank_to_dense) and is what the apply GEMMs read
        bk.packed = torch.zeros((bk.count, s, bk.pdim), dtype=torch.float32, device=dev)
        bk.precs = torch.zeros((bk.count, s, s), dtype=torch.float32, device=dev)
        # the same operator in factored form, c g + (g V) W^T (rank-2 blocks apply it that way)
        bk.loww = torch.zeros((bk.count, s, bk.pdim - 2), dtype=torch.float32, device=dev)
        bk.lowc = torch.zeros((bk.count,), dtype=torch.float32, device=dev)
        bk.needs_dense = False
      else:
        bk.precs = eye.repeat(bk.count, 1, 1).contiguous()
      bk.exps = torch.tensor(bk.exponents, dtype=torch.int32, device=dev)
      bk.exps_host = np.asarray(bk.exponents, dtype=np.int32)
      self.metrics[s] = torch.zeros((bk.count, 5), dtype=torch.float32, device=dev)
      if self.quantize_second_moment:
        q, d, b = ops.quantize(bk.stats, self.qdt_second, True)
        bk.qstats = [q, d, b]
        q, d, b = ops.quantize(bk.precs, self.qdt_second, True)
        bk.qprecs = [q, d, b]
    self._build_launch_lists(leaves)
    # the whole grafting / momentum tail runs as one grouped call over the flat buffers
    self._graft_group = ops.GraftGroup(
        [(pl.offset, pl.numel, not pl.skip) for pl in self.plans], dev)
    self._gviews = [self.gbuf[pl.offset:pl.offset + pl.numel] for pl in self.plans]
    self._pviews = ([self.pbuf[pl.offset:pl.offset + pl.numel] for pl in self.plans]
                    if self.pbuf is not None else None)

    stats = []
    for plan, p in zip(self.plans, leaves):
      seg = slice(plan.offset, plan.offset + plan.numel)
      diag = self.dsbuf[seg].view(p.shape) if self._graft_has_diag() else []
      if plan.mdt == torch.float32:
        mom = QuantizedValue.from_float_value(self.mbuf[seg].view(p.shape), torch.float32)
        dmom = QuantizedValue.from_float_value(self.dmbuf[seg].view(p.shape), torch.float32)
      else:  # int8 momenta (DS:2111-2114): the flat segments are the fp32 scratch of a step
        z = torch.zeros(p.shape, dtype=torch.float32, device=dev)
        mom = QuantizedValue.from_float_value(z, plan.mdt)
        dmom = QuantizedValue.from_float_value(z, plan.mdt)
      st = ParameterStats(
          QuantizedValue.from_float_value(diag, torch.float32),
          [self._stat_view(r) for r in plan.stat_refs],
          [self._prec_view(r) for r in plan.stat_refs],
          dmom, mom,
          (self.agbuf[seg].view(p.shape) if self.use_avg_grad else None),
          ([self.metrics[s][i] for s, i in plan.stat_refs]
           if self.generate_training_metrics else None), dev)
      stats.append(st)
    self._own_leaves = stats
    self._own_ptrs = [[t.data_ptr() for t in _state_tensors(st)] for st in stats]
    self._built = True
    return ShampooState(0, self._wrap_stats(stats))

  def _wrap_stats(self, stats):
    """State tree handed to the caller: per-parameter ParameterStats, or in sharded-state mode
    ShardedShampooStats(global stacked arrays, per-parameter local stats with index_start /
    sizes) -- DS:2243-2256.  The local entries are the same ParameterStats views (``update``
    finds them under ``.local_stats``)."""
    tree = _tree_unflatten(self.treedef, stats)
    if not self.stacked:
      return tree
    if not self.buckets:
      return ShardedShampooStats(GlobalShardedParameterStats(None, None, None), tree)
    (bk,) = self.buckets.values()
    local, start = [], 0
    for plan, st in zip(self.plans, stats):
      local.append(LocalShardedParameterStats(
          st.diagonal_statistics, st.diagonal_momentum, st.momentum, st.avg_grad,
          st.training_metrics, start, list(plan.stat_sizes)))
      start += len(plan.stat_sizes)
    self._sharded_local = local
    return ShardedShampooStats(
        GlobalShardedParameterStats(bk.stats, bk.precs, bk.exps),
        _tree_unflatten(self.treedef, local))

  def _stat_view(self, ref):
    s, i = ref
    bk = self.buckets[s]
    if self.stacked:  # padded row of the stacked array, if this rank holds it
      j = i - bk.lo
      return bk.stats[j] if 0 <= j < bk.n_local else None
    if self.quantize_second_moment:
      q, d, b = bk.qstats
      return QuantizedValue(q[i], d[i], b[i], self.qdt_second, True, [s, s])
    return bk.stats[i]

  def _prec_view(self, ref):
    s, i = ref
    bk = self.buckets[s]
    if self.quantize_second_moment:
      q, d, b = bk.qprecs
      return QuantizedValue(q[i], d[i], b[i], self.qdt_second, True, [s, s])
    return bk.packed[i] if bk.compressed else bk.precs[i]

  # ---- state export / import -------------------------------------------------
  def export_state(self, state):
    """Deep copy of ``state`` that no longer aliases the optimizer's buffers (checkpoint)."""
    leaves = self._flatten_stats(state.stats)
    return ShampooState(int(state.count),
                        _tree_unflatten(self.treedef, [_clone_stats(st) for st in leaves]))

  def import_state(self, state):
    """Copies a foreign state (restored checkpoint, earlier export, deep copy) into the
    optimizer's buffers and returns the equivalent state that aliases them."""
    assert self._built, "call init(params) first"
    leaves = self._flatten_stats(state.stats)
    if len(leaves) != len(self._own_leaves):
      raise ValueError("state does not match the parameters this optimizer was initialised with")
    for own, other in zip(self._own_leaves, leaves):
      if other is own:
        continue
      a, b = _state_tensors(own), _state_tensors(other)
      if len(a) != len(b) or any(x.shape != y.shape or x.dtype != y.dtype for x, y in zip(a, b)):
        raise ValueError("state layout does not match this optimizer's configuration")
      for x, y in zip(a, b):
        if x.data_ptr() != y.data_ptr():
          x.copy_(y)
    r = abs(self.compression_rank)
    for bk in self.buckets.values():
      if bk.compressed:  # the applied operator is derived from the packed sketch
        self._refresh_low_rank(bk)
    return ShampooState(int(state.count), _tree_unflatten(self.treedef, self._own_leaves))

  def _adopt(self, state):
    """``update`` advances the buffers behind the state returned by ``init`` / ``update`` /
    ``import_state``; any other state (checkpoint, copy, another init) is imported first."""
    leaves = self._flatten_stats(state.stats)
    if len(leaves) == len(self._own_leaves):
      same = True
      for own, ptrs, st in zip(self._own_leaves, self._own_ptrs, leaves):
        if st is own:
          continue
        if [t.data_ptr() for t in _state_tensors(st)] != ptrs:
          same = False
          break
      if same:
        return
    self.import_state(state)

  # ---- static launch lists ------------------------------------------------
  def _build_launch_lists(self, leaves):
    """Grouped-GEMM descriptors for the statistics update (DS:1582-1590) and the
    preconditioner application (DS:1676-1708), built once."""
    D = _lib.GemmDesc
    stat_descs, stat_meta, apply_descs = [], [], []
    apply_qsrc = {}  # id(descriptor) -> (bucket, index) of the quantised preconditioner it applies
    # blocks of merged rank > 3: their mode unfoldings are no two-level strided views of the flat
    # gradient buffer, so each block is staged through contiguous copies (in: before the
    # statistics, out: after the last mode product); DS:1676-1708 loops over any rank
    self._staged = []
    self._lowrank_tmp = []  # intermediates of the factored low-rank application (kept alive)
    self._lowrank_apply = os.environ.get("PC_LOWRANK_APPLY", "1") != "0"
    for bk in self.buckets.values():
      if getattr(bk, "compressed", False):
        bk.needs_dense = False
    w1 = float(self.beta2)
    w2 = float(self.beta2 if self.beta2 == 1.0 else 1.0 - self.beta2)  # DS:2635-2636
    f32 = 4
    g0, pg0, t10, t20 = (self.gbuf.data_ptr(), self.pgbuf.data_ptr(), self.t1buf.data_ptr(),
                         self.t2buf.data_ptr())
    sg0 = self.sgbuf.data_ptr() if self.use_avg_grad else g0  # source of the statistics
    for plan in self.plans:
      if plan.skip:
        continue
      dims = list(plan.tshape)
      rank = len(dims)
      strides = [int(np.prod(dims[i + 1:])) for i in range(rank)]
      tmp_off = 0
      for offs, sizes, refs in plan.blocks:
        base_elem = plan.offset + sum(o * st for o, st in zip(offs, strides))
        gbase = g0 + f32 * base_elem
        staged = rank > 3
        if staged:
          st_in = torch.empty(sizes, dtype=torch.float32, device=self.device)
          st_stat = torch.empty_like(st_in) if self.use_avg_grad else st_in
          st_out = torch.empty_like(st_in)
          box = tuple(slice(o, o + z) for o, z in zip(offs, sizes))
          self._staged.append((plan, box, st_in, st_stat, st_out))
          blk_strides = [int(np.prod(sizes[i + 1:])) for i in range(rank)]
        # ---- statistics: one Gram product per preconditioned axis ----
        for axis in range(rank):
          if refs[axis] is None:
            continue
          S, bi = refs[axis]  # S = row stride of the bucket (== s unless stacked)
          s = sizes[axis]
          bk = self.buckets[S]
          if self.stacked and not bk.lo <= bi < bk.lo + bk.n_local:
            continue  # sharded statistics: another rank owns (and updates) this row
          cptr = bk.stats.data_ptr() + f32 * (bi - bk.lo) * S * S
          others = [a for a in range(rank) if a != axis]
          k = int(np.prod([sizes[a] for a in others])) if others else 1
          d = D()
          d.a = d.b = sg0 + f32 * base_elem
          d.c = d.c_in = cptr
          d.a_si = d.b_sj = strides[axis]
          d.a_iinner, d.a_sio = sizes[axis], 0
          if staged:
            # contiguous block: k = (axes before `axis`, flattened) x (axes after it, flattened)
            d.a = d.b = st_stat.data_ptr()
            d.a_si = d.b_sj = blk_strides[axis]
            suf = blk_strides[axis]
            kin, sko, ski = suf, sizes[axis] * suf, 1
          elif len(others) == 0:
            kin, sko, ski = 1, 0, 0
          elif len(others) == 1:
            kin, sko, ski = sizes[others[0]], 0, strides[others[0]]
          else:
            kin, sko, ski = sizes[others[1]], strides[others[0]], strides[others[1]]
          d.a_kinner = d.b_kinner = kin
          d.a_sko = d.b_sko = sko
          d.a_ski = d.b_ski = ski
          d.c_iinner, d.c_sio, d.c_sii = s, 0, S
          d.m = d.n = s
          d.k = k
          d.alpha, d.beta = w2, w1
          if self.frequent_directions and bk.compressed:
            # frequent_directions_update (DS:1473-1505) ignores the old statistic and the
            # weights: R R^T = x x^T.  Only that product enters the sketch update, so the
            # statistic slot holds x x^T itself (no QR on the device).
            d.c_in, d.alpha, d.beta = None, 1.0, 0.0
          stat_descs.append(d)
          stat_meta.append((S, bi))
        # ---- application: contract the leading axis and roll (DS:1678-1707) ----
        bnumel = int(np.prod(sizes))
        cur_ptr, cur_strides, cur_sizes = gbase, list(strides), list(sizes)
        if staged:
          cur_ptr, cur_strides = st_in.data_ptr(), list(blk_strides)
        t_ptrs = [t10 + f32 * (plan.offset + tmp_off), t20 + f32 * (plan.offset + tmp_off)]
        while len(apply_descs) < rank:
          apply_descs.append([])
        comp_axes = [refs[a] is not None and self.buckets[refs[a][0]].compressed
                     for a in range(rank)]
        if rank == 2 and any(comp_axes) and self._lowrank_apply:
          self._lowrank_descs(apply_descs, sizes, refs, comp_axes, gbase, pg0 + f32 * base_elem,
                              strides[0])
          tmp_off += int(np.prod(sizes))
          continue
        for a in range(rank):
          if comp_axes[a]:
            self.buckets[refs[a][0]].needs_dense = True
        for j in range(rank):
          d0 = cur_sizes[0]
          rest_sizes, rest_strides = cur_sizes[1:], cur_strides[1:]
          rest = int(np.prod(rest_sizes)) if rest_sizes else 1
          last = j == rank - 1
          d = D()
          # A(i = rest index, k = leading index)
          d.a = cur_ptr
          d.a_kinner, d.a_sko, d.a_ski = d0, 0, cur_strides[0]
          if len(rest_sizes) == 0:
            d.a_iinner, d.a_sio, d.a_si = 1, 0, 0
          elif len(rest_sizes) == 1:
            d.a_iinner, d.a_sio, d.a_si = rest_sizes[0], 0, rest_strides[0]
          elif staged:  # every intermediate of a staged block is contiguous
            d.a_iinner, d.a_sio, d.a_si = rest, 0, 1
          else:
            d.a_iinner, d.a_sio, d.a_si = rest_sizes[1], rest_strides[0], rest_strides[1]
          # B(j = output column, k) = P[k, j]
          if refs[j] is not None:
            S, bi = refs[j]  # row stride of the stored preconditioner
            pptr = self.buckets[S].precs.data_ptr() + f32 * bi * S * S
            if self.quantize_second_moment and not self.buckets[S].compressed:
              apply_qsrc[id(d)] = (S, bi)  # its QuantizedValue can feed the GEMM directly
          else:  # not preconditioned: pure roll (DS:1684-1686) -> multiply by I
            S = d0
            pptr = self._identity(d0).data_ptr()
          d.b = pptr
          d.b_sj, d.b_kinner, d.b_sko, d.b_ski = 1, d0, 0, S
          d.m, d.n, d.k = rest, d0, d0
          d.alpha, d.beta = 1.0, 0.0
          d.c_in = None
          new_sizes = rest_sizes + [d0]
          if last and staged:
            d.c = st_out.data_ptr()
            d.c_iinner, d.c_sio, d.c_sii = rest, 0, d0
            new_strides = None
          elif last:
            # final layout == original axis order: write into the param-shaped buffer
            d.c = pg0 + f32 * base_elem
            if rank == 1:
              d.c_iinner, d.c_sio, d.c_sii = 1, 0, 0
            elif rank == 2:
              d.c_iinner, d.c_sio, d.c_sii = rest, 0, strides[0]
            else:
              d.c_iinner, d.c_sio, d.c_sii = new_sizes[1], strides[0], strides[1]
            new_strides = None
          else:
            d.c = t_ptrs[j % 2]
            d.c_iinner, d.c_sio, d.c_sii = max(rest, 1), 0, d0
            new_strides = [int(np.prod(new_sizes[i + 1:])) for i in range(rank)]
          apply_descs[j].append(d)
          cur_ptr, cur_sizes, cur_strides = d.c, new_sizes, new_strides
        tmp_off += bnumel
    # Blocks whose GEMM output sizes are multiples of 128 go to the tcgen05 grouped GEMM
    # (scaled-fp16 three-pass products, symmetric rank-k for the Gram update); the rest
    # stay on the fp32 CUDA-core tile.
    use_tc = (self.engine != _lib.PC_ENGINE_SIMT_FP32 and
              bool(_lib.load().pc_device_supports_tcgen05()))

    def split(descs):
      # (outer-product updates of rank-1 parameters stream through a CUDA-core kernel)
      on_tc = lambda d: use_tc and ops.tc_gemm_eligible(d) and not ops.thin_outer_eligible(d)
      tc = [d for d in descs if on_tc(d)]
      simt = [d for d in descs if not on_tc(d)]
      return (ops.TcGemmList(tc, self.device) if tc else None,
              ops.SimtGemmLists(simt, self.device) if simt else None)

    self._stat_count = len(stat_descs)
    self._stat_tc_fused = None
    for bk in self.buckets.values():
      bk.fused_quant = False
    if self.quantize_second_moment and use_tc:
      # Quantised second moments on the tcgen05 path: to_float is fused into the epilogue's
      # C_in read and the column-max reduction of from_float into its write
      # (pc_grouped_gemm_tc_quant); the requantisation is then a single pass.
      fused, fused_ext, rest = [], [], []
      for d, (sz, bi) in zip(stat_descs, stat_meta):
        if not ops.tc_gemm_fused_quant_eligible(d):
          rest.append(d)
          continue
        bk = self.buckets[sz]
        if not bk.fused_quant:
          bk.fused_quant = True
          bk.colmax = torch.zeros((bk.count, sz), dtype=torch.int32, device=self.device)
        q, dg, bs = bk.qstats
        e = _lib.GemmQuant()
        e.q_in = q[bi].data_ptr()
        e.diag_in = dg[bi].data_ptr()
        e.bucket_in = bs[bi].data_ptr()
        e.colmax_out = bk.colmax[bi].data_ptr()
        e.qdtype = ops._QDT[self.qdt_second]
        d.c_in = None
        fused.append(d)
        fused_ext.append(e)
      if fused:
        self._stat_tc_fused = ops.TcGemmList(fused, self.device, quant=fused_ext)
      stat_descs = rest
    self._stat_tc, self._stat_simt = split(stat_descs)
    self._apply_tc, self._apply_simt = [], []
    # quantised preconditioners (DS:3556): products on the tcgen05 path read the int16 / int8
    # QuantizedValue while packing the operand (pc_gemm_quant.b_q); only buckets with a consumer
    # on the CUDA-core path are still dequantised into `precs` before the application
    for bk in self.buckets.values():
      bk.dequant_for_apply = False
    for lst in apply_descs:
      tc = [d for d in lst if use_tc and ops.tc_gemm_eligible(d)]
      simt = [d for d in lst if not (use_tc and ops.tc_gemm_eligible(d))]
      for d in simt:
        if id(d) in apply_qsrc:
          self.buckets[apply_qsrc[id(d)][0]].dequant_for_apply = True
      ext = None
      if tc and any(id(d) in apply_qsrc for d in tc):
        ext = []
        for d in tc:
          e = _lib.GemmQuant()
          if id(d) in apply_qsrc:
            S, bi = apply_qsrc[id(d)]
            q, dg, bs = self.buckets[S].qprecs
            e.b_q, e.b_diag, e.b_bucket = q[bi].data_ptr(), dg[bi].data_ptr(), bs[bi].data_ptr()
            e.b_ld, e.b_qdtype = S, ops._QDT[self.qdt_second]
          ext.append(e)
      self._apply_tc.append(ops.TcGemmList(tc, self.device, quant=ext) if tc else None)
      self._apply_simt.append(ops.SimtGemmLists(simt, self.device) if simt else None)


  def _lowrank_descs(self, apply_descs, sizes, refs, comp_axes, x_ptr, z_ptr, ld):
    """Application descriptors of one rank-2 block [d0, d1] (row stride `ld` in the gradient and
    in the output buffer) with at least one packed low-rank preconditioner (DS:1690-1705):
      axis 0:  Y = P0 X = c0 X + W0 (V0^T X)        axis 1:  Z = Y P1 = c1 Y + (Y V1) W1^T
    two thin products per compressed axis (the `c` term rides on the second one as C_in with a
    device-scalar weight), one dense product per full axis, no roll.  Steps 0-3 of the launch
    lists; Y keeps the gradient's row stride so that C_in and C are addressed alike."""
    D = _lib.GemmDesc
    f32 = 4
    d0, d1 = sizes
    dev = self.device
    while len(apply_descs) < 4:
      apply_descs.append([])
    y = torch.empty((d0 - 1) * ld + d1, dtype=torch.float32, device=dev)
    self._lowrank_tmp.append(y)
    y_ptr = y.data_ptr()

    def prec(a):
      if refs[a] is None:
        n = sizes[a]
        return self._identity(n).data_ptr(), n
      S, bi = refs[a]
      return self.buckets[S].precs.data_ptr() + f32 * bi * S * S, S

    def factors(a):
      S, bi = refs[a]
      bk = self.buckets[S]
      pd = bk.pdim
      return (bk.packed.data_ptr() + f32 * bi * S * pd, pd,
              bk.loww.data_ptr() + f32 * bi * S * (pd - 2), pd - 2,
              bk.lowc.data_ptr() + f32 * bi)

    # ---- axis 0: Y = P0 X ----
    if comp_axes[0]:
      v, pd, w, r, c = factors(0)
      t = torch.empty(r * d1, dtype=torch.float32, device=dev)
      self._lowrank_tmp.append(t)
      g = D()  # T0[q, j] = sum_k V[k, q] X[k, j]
      g.a, g.b, g.c, g.c_in = v, x_ptr, t.data_ptr(), None
      g.a_iinner, g.a_sio, g.a_si, g.a_kinner, g.a_sko, g.a_ski = r, 0, 1, d0, 0, pd
      g.b_sj, g.b_kinner, g.b_sko, g.b_ski = 1, d0, 0, ld
      g.c_iinner, g.c_sio, g.c_sii = r, 0, d1
      g.m, g.n, g.k, g.alpha, g.beta = r, d1, d0, 1.0, 0.0
      apply_descs[0].append(g)
      g = D()  # Y = W T0 + c X
      g.a, g.b, g.c, g.c_in, g.beta_dev = w, t.data_ptr(), y_ptr, x_ptr, c
      g.a_iinner, g.a_sio, g.a_si, g.a_kinner, g.a_sko, g.a_ski = d0, 0, r, r, 0, 1
      g.b_sj, g.b_kinner, g.b_sko, g.b_ski = 1, r, 0, d1
      g.c_iinner, g.c_sio, g.c_sii = d0, 0, ld
      g.m, g.n, g.k, g.alpha, g.beta = d0, d1, r, 1.0, 1.0
      apply_descs[1].append(g)
    else:
      p, S = prec(0)
      g = D()  # Y[i, j] = sum_k P0[i, k] X[k, j]
      g.a, g.b, g.c, g.c_in = p, x_ptr, y_ptr, None
      g.a_iinner, g.a_sio, g.a_si, g.a_kinner, g.a_sko, g.a_ski = d0, 0, S, d0, 0, 1
      g.b_sj, g.b_kinner, g.b_sko, g.b_ski = 1, d0, 0, ld
      g.c_iinner, g.c_sio, g.c_sii = d0, 0, ld
      g.m, g.n, g.k, g.alpha, g.beta = d0, d1, d0, 1.0, 0.0
      apply_descs[1].append(g)
    # ---- axis 1: Z = Y P1 ----
    if comp_axes[1]:
      v, pd, w, r, c = factors(1)
      t = torch.empty(d0 * r, dtype=torch.float32, device=dev)
      self._lowrank_tmp.append(t)
      g = D()  # T1[i, q] = sum_k Y[i, k] V[k, q]
      g.a, g.b, g.c, g.c_in = y_ptr, v, t.data_ptr(), None
      g.a_iinner, g.a_sio, g.a_si, g.a_kinner, g.a_sko, g.a_ski = d0, 0, ld, d1, 0, 1
      g.b_sj, g.b_kinner, g.b_sko, g.b_ski = 1, d1, 0, pd
      g.c_iinner, g.c_sio, g.c_sii = d0, 0, r
      g.m, g.n, g.k, g.alpha, g.beta = d0, r, d1, 1.0, 0.0
      apply_descs[2].append(g)
      g = D()  # Z = T1 W^T + c Y
      g.a, g.b, g.c, g.c_in, g.beta_dev = t.data_ptr(), w, z_ptr, y_ptr, c
      g.a_iinner, g.a_sio, g.a_si, g.a_kinner, g.a_sko, g.a_ski = d0, 0, r, r, 0, 1
      g.b_sj, g.b_kinner, g.b_sko, g.b_ski = r, r, 0, 1
      g.c_iinner, g.c_sio, g.c_sii = d0, 0, ld
      g.m, g.n, g.k, g.alpha, g.beta = d0, d1, r, 1.0, 1.0
      apply_descs[3].append(g)
    else:
      p, S = prec(1)
      g = D()  # Z[i, j] = sum_k Y[i, k] P1[j, k]  (P1 symmetric)
      g.a, g.b, g.c, g.c_in = y_ptr, p, z_ptr, None
      g.a_iinner, g.a_sio, g.a_si, g.a_kinner, g.a_sko, g.a_ski = d0, 0, ld, d1, 0, 1
      g.b_sj, g.b_kinner, g.b_sko, g.b_ski = S, d1, 0, 1
      g.c_iinner, g.c_sio, g.c_sii = d0, 0, ld
      g.m, g.n, g.k, g.alpha, g.beta = d0, d1, d1, 1.0, 0.0
      apply_descs[3].append(g)

  def _identity(self, n):
    cache = self.__dict__.setdefault("_eyes", {})
    if n not in cache:
      cache[n] = torch.eye(n, dtype=torch.float32, device=self.device)
    return cache[n]

  # ---- update (DS:3627-3659) ----------------------------------------------
  def update(self, grads, state, params=None):
    assert self._built, "call init(params) first"
    step = int(state.count)
    g_leaves, _ = _tree_flatten(grads)
    if len(g_leaves) != len(self.plans):
      raise ValueError("gradient tree does not match the parameters given to init")
    self._adopt(state)
    for g, plan in zip(g_leaves, self.plans):
      if tuple(g.shape) != plan.shape:
        raise ValueError(f"gradient shape {tuple(g.shape)} != parameter shape {plan.shape}")
      if not g.is_cuda or not g.is_floating_point():
        raise TypeError("gradients must be floating-point CUDA tensors")
    # (0) stage gradients (any float dtype, any memory layout) into the flat fp32 buffer the
    #     static descriptors point at: one fused pass
    torch._foreach_copy_(self._gviews, [g.reshape(-1) for g in g_leaves])
    if self.weight_decay != 0:
      if params is None:
        raise ValueError("weight_decay needs params")
      p_leaves = _tree_flatten(params)[0]
      torch._foreach_copy_(self._pviews, [p.reshape(-1) for p in p_leaves])
    if self.use_avg_grad:  # DS:2640-2645
      k = self.statistics_compute_steps
      if k == 1 or step % k == 1:
        self.agbuf.copy_(self.gbuf)
      else:
        self.agbuf.add_(self.gbuf)
      torch.div(self.agbuf, float(k), out=self.sgbuf)
    if self._staged:
      self._stage_in()
    # (1) statistics (DS:3644 -> DS:2631-2675)
    if self._stat_count and (self.statistics_compute_steps <= 1 or
                                step % self.statistics_compute_steps == 0):
      self._update_statistics()
    # (2) preconditioners (DS:3648 -> DS:3442-3494), optionally on the schedule of DS:2909-2934
    pcs = self.preconditioning_compute_steps
    if (self.decay_preconditioning_compute_steps and self.end_preconditioning_compute_steps and
        callable(self.learning_rate)):
      pcs = preconditioning_compute_steps_schedule(
          self.learning_rate, self.preconditioning_compute_steps,
          self.end_preconditioning_compute_steps, step)
    if self.buckets and step % pcs == 0:
      self._compute_preconditioners(step)
    # (3) transform (DS:3650 -> DS:3496-3625)
    self._apply_preconditioners()
    lr = self.learning_rate(step) if callable(self.learning_rate) else self.learning_rate
    ubuf = self._transform_all(step, float(lr))
    updates = []
    for plan, g in zip(self.plans, g_leaves):
      u = ubuf[plan.offset:plan.offset + plan.numel].view(plan.shape)
      updates.append(u if g.dtype == torch.float32 else u.to(g.dtype))
    return (_tree_unflatten(self.treedef, updates),
            ShampooState(step + 1, self._wrap_stats(self._own_leaves)))

  def _flatten_stats(self, stats_tree):
    out = []

    def rec(t):
      if isinstance(t, ShardedShampooStats):
        # the local entries mirror the optimizer's own ParameterStats one to one
        n = len(_tree_flatten_any(t.local_stats, LocalShardedParameterStats))
        if getattr(self, "_own_leaves", None) is not None and n == len(self._own_leaves):
          out.extend(self._own_leaves)
          return
        raise ValueError("sharded state does not match this optimizer")
      if isinstance(t, ParameterStats):
        out.append(t)
      elif isinstance(t, dict):
        for k in sorted(t.keys()):
          rec(t[k])
      else:
        for v in t:
          rec(v)

    rec(stats_tree)
    return out

  def _stage_in(self):
    """Contiguous copies of the blocks of rank > 3 tensors (see _build_launch_lists)."""
    for plan, box, st_in, st_stat, _ in self._staged:
      flat = slice(plan.offset, plan.offset + plan.numel)
      st_in.copy_(self.gbuf[flat].view(plan.tshape)[box])
      if st_stat is not st_in:
        st_stat.copy_(self.sgbuf[flat].view(plan.tshape)[box])

  def _stage_out(self):
    for plan, box, _, _, st_out in self._staged:
      self.pgbuf[plan.offset:plan.offset + plan.numel].view(plan.tshape)[box].copy_(st_out)

  def _update_statistics(self):
    if self.quantize_second_moment:  # to_float (DS:1588, QU:97-113)
      for bk in self.buckets.values():
        if bk.fused_quant:
          bk.colmax.zero_()
          continue
        q, d, b = bk.qstats
        ops.dequantize(q, d, b, True, out=bk.stats)
    if self._stat_simt is not None:
      self._stat_simt.run()
    if self._stat_tc is not None:
      self._stat_tc.run()
    if self._stat_tc_fused is not None:
      self._stat_tc_fused.run()
    if self.quantize_second_moment:  # from_float (DS:2654)
      for bk in self.buckets.values():
        q, d, b = bk.qstats
        if bk.fused_quant:
          ops.quantize_from_colmax(bk.stats, bk.colmax, self.qdt_second, q, d, b)
        else:
          ops.quantize(bk.stats, self.qdt_second, True, out=(q, d, b))

  # ---- preconditioners -------------------------------------------------------
  def _compute_preconditioners(self, step=0):
    world, rank = self._world()
    if any(bk.compressed for bk in self.buckets.values()):
      if self.frequent_directions:
        self._fd_update(step, world, rank)
      else:
        self._low_rank_update(world, rank)
    full = [bk for _, bk in sorted(self.buckets.items()) if not bk.compressed]
    if not full:
      return
    if self.eigh:
      for bk in full:
        self._eigh_roots(bk, world, rank)
      return
    if self.lobpcg_topk_precondition > 0:
      # top-k deflation before the Newton iteration (DS:789-812); statistics too small for the
      # requested k take the plain path
      plain = [bk for bk in full if bk.size <= self.lobpcg_topk_precondition + 2]
      for bk in full:
        if bk not in plain:
          self._lobpcg_roots(bk, world, rank)
      if plain:
        self._newton_roots(plain, world, rank)
      return
    self._newton_roots(full, world, rank)

  def _padded_size(self, s):
    """Size of the problem a statistic of size ``s`` is solved in.  The reference pads every
    statistic to the largest block (DS:2841-2843, masked by DS:777-783); here only where it
    pays: 1 x 1 statistics take the iterative path like upstream (DS:850-855 only fires when
    max_size == 1) as [[s, 0], [0, 0]] with padding_start = 1, and sizes such as 1000 or 576
    are embedded in the next multiple of 128 so that they run on the tcgen05 engine."""
    if s == 1 and max(self.buckets) > 1:
      return 2
    if (s >= 256 and s % 128 != 0 and self.engine == _lib.PC_ENGINE_AUTO and
        bool(_lib.load().pc_device_supports_tcgen05())):
      return (s + 127) // 128 * 128
    return s

  def _build_root_jobs(self, full, world, rank):
    """Partition + static buffers of the Newton-root solves (built once per world size)."""
    costs = [(bk.size, [float(self._padded_size(bk.size)) ** 3 * newton_gemms_per_iteration(p)
                        for p in bk.exponents]) for bk in full]
    table = partition_statistics(costs, world)
    dev = self.device
    for bk in full:
      s, sp = bk.size, self._padded_size(bk.size)
      job = _RootJob()
      job.world, job.sp = world, sp
      if self.stacked:  # sharded-state layout: rank r owns the contiguous rows it stores
        owned = [list(range(r * bk.n_local, (r + 1) * bk.n_local)) for r in range(world)]
      else:
        owned = table[s]
      job.mine = owned[rank]
      job.cnt = max(len(o) for o in owned)  # local batch (fillers pad the shorter ranks)
      job.in_place = (world == 1 and sp == s) or self.stacked
      exps = np.ones(job.cnt, dtype=np.int32)
      pads = np.zeros(job.cnt, dtype=np.int32)
      exps[:len(job.mine)] = bk.exps_host[job.mine]
      pads[:len(job.mine)] = np.asarray(bk.true_sizes, dtype=np.int32)[job.mine]
      job.exps_host = exps
      job.exps = torch.from_numpy(exps).to(dev)
      job.pads = torch.from_numpy(pads).to(dev)
      job.mine_idx = torch.tensor(job.mine, dtype=torch.int64, device=dev)
      job.x = bk.stats if job.in_place else torch.zeros((job.cnt, sp, sp), dtype=torch.float32,
                                                        device=dev)
      job.ws = torch.empty(ops.root_workspace_bytes(job.cnt, sp, self.engine) + 256,
                           dtype=torch.uint8, device=dev)
      job.stream = torch.cuda.Stream(dev)
      job.done = torch.cuda.Event()
      qd = self.qdt_second
      qbytes = {torch.int16: 2, torch.int8: 1}.get(qd, 4)
      if world == 1:
        job.roots = torch.empty((job.cnt, sp, sp), dtype=torch.float32, device=dev)
        job.metrics = self.metrics[s] if not self.quantize_second_moment else torch.empty(
            (job.cnt, 5), dtype=torch.float32, device=dev)
        if self.quantize_second_moment:
          job.q = torch.empty((job.cnt, s, s), dtype=qd, device=dev)
          job.qd = torch.empty((job.cnt, s), dtype=torch.float32, device=dev)
          job.qb = torch.empty((job.cnt, s), dtype=torch.float32, device=dev)
          ar = torch.arange(job.cnt, device=dev)
          job.sel = [(job.q, ar * (s * s * qbytes), s * s * qbytes, 0),
                     (job.qd, ar * (s * 4), s * 4, 1), (job.qb, ar * (s * 4), s * 4, 2)]
          job.met_off = ar * 5
          job.dst_idx = ar.to(torch.int32)
      else:
        # one packed payload per rank and bucket: [roots | metrics] in fp32, or
        # [q | diagonal | bucket sizes | metrics] when the state is quantised (DS:3116-3122)
        if self.quantize_second_moment:
          sections = [("q", s * s * qbytes), ("d", s * 4), ("b", s * 4), ("m", 20)]
        else:
          sections = [("r", s * s * 4), ("m", 20)]
        lay = gather_layout(owned, job.cnt, sections)
        off, sec = lay["lbytes"], lay["sections"]
        job.lbytes = off
        job.send = torch.zeros(off, dtype=torch.uint8, device=dev)
        # copy-engine pushes over NVLink peer memory (no SM: overlaps the other buckets'
        # persistent GEMM kernels); NCCL when peer mappings are unavailable
        job.gather = peer.make_all_gather(off, self.process_group, dev)
        job.recv = job.gather.recv
        job.recv_f32 = job.recv.view(-1).view(torch.float32)

        def view(name, dtype, shape):
          o, row = sec[name]
          return job.send[o:o + job.cnt * row].view(dtype).view(shape)

        job.metrics = view("m", torch.float32, (job.cnt, 5))
        if self.quantize_second_moment:
          job.roots = torch.empty((job.cnt, sp, sp), dtype=torch.float32, device=dev)
          job.q = view("q", qd, (job.cnt, s, s))
          job.qd = view("d", torch.float32, (job.cnt, s))
          job.qb = view("b", torch.float32, (job.cnt, s))
        elif sp == s:
          job.roots = view("r", torch.float32, (job.cnt, s, s))
        else:
          job.roots = torch.empty((job.cnt, sp, sp), dtype=torch.float32, device=dev)
          job.send_roots = view("r", torch.float32, (job.cnt, s, s))
        job.dst_idx = torch.from_numpy(lay["dst_index"]).to(dev)
        job.met_off = torch.from_numpy(lay["metrics_offset"]).to(dev)
        job.sel = []
        for k, name in enumerate(("q", "d", "b") if self.quantize_second_moment else ("r",)):
          job.sel.append((job.recv, torch.from_numpy(lay["src_offset"][name]).to(dev),
                          sec[name][1], k))
      bk.job = job

  def _newton_roots(self, full, world, rank):
    """Coupled-Newton roots of every full statistic.  Each bucket (one statistic size) runs on
    its own stream: the solver only enqueues (CUDA graph, device-side convergence), so the
    buckets' power iterations / GEMM chains overlap on the GPU, and with more than one rank
    each bucket's packed all-gather (DS:2876-2877) starts as soon as that bucket is done."""
    if full[0].job is None or full[0].job.world != world:
      self._build_root_jobs(full, world, rank)
    main = torch.cuda.current_stream(self.device)
    fork = torch.cuda.Event()
    fork.record(main)
    kw = dict(ridge_epsilon=self.matrix_epsilon,
              relative_matrix_epsilon=self.relative_matrix_epsilon, engine=self.engine)
    thr = float(self.inverse_failure_threshold)
    # small buckets first: their all-gathers leave the NCCL queue before the big one arrives
    for bk in sorted(full, key=lambda b: b.job.cnt * b.job.sp ** 3):
      job, s = bk.job, bk.size
      with torch.cuda.stream(job.stream):
        job.stream.wait_event(fork)
        if self.quantize_second_moment:
          q, d, b = bk.qstats
          ops.dequantize(q, d, b, True, out=bk.stats)
        if not job.in_place and job.mine:
          if job.sp == s:
            torch.index_select(bk.stats, 0, job.mine_idx, out=job.x[:len(job.mine)])
          else:
            job.x[:len(job.mine), :s, :s].copy_(
                bk.stats if world == 1 else bk.stats.index_select(0, job.mine_idx))
        ops.matrix_inverse_pth_root_batched(
            job.x, job.exps, None if (job.in_place and not self.stacked) else job.pads,
            out=job.roots,
            metrics_out=job.metrics, workspace=job.ws, ps_host=job.exps_host, **kw)
        corner = job.roots if job.sp == s else job.roots[:, :s, :s]
        if self.quantize_second_moment:
          # DS:2746-2772: the new root is requantised by its owner; DS:3183-3208: select
          ops.quantize(corner if job.sp == s else corner.contiguous(), self.qdt_second, True,
                       out=(job.q, job.qd, job.qb))
        elif world > 1 and job.sp != s:
          job.send_roots.copy_(corner)
        if world > 1:
          job.gather.all_gather(job.send)
        if world == 1 and not self.quantize_second_moment:
          lib = _lib.load()
          _lib.check(lib.pc_select_preconditioners(
              ctypes.c_void_p(job.roots.data_ptr()), ctypes.c_void_p(job.metrics.data_ptr()), thr,
              ctypes.c_void_p(bk.precs.data_ptr()), bk.count, job.sp, job.sp, s, s,
              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        else:
          mbase = job.recv_f32 if world > 1 else job.metrics.view(-1)
          dsts = bk.qprecs if self.quantize_second_moment else [bk.precs]
          for src, src_off, row, k in job.sel:
            ops.select_scatter(src, src_off, mbase, job.met_off, job.dst_idx, thr, dsts[k], row,
                               self.metrics[s] if k == 0 else None)
        if world > 1:
          job.gather.release()
        job.done.record(job.stream)
    for bk in full:
      main.wait_event(bk.job.done)

  def _eigh_roots(self, bk, world, rank):
    """`eigh=True`: matrix_inverse_pth_root_eigh for every statistic (DS:2677-2684)."""
    s = bk.size
    kw = dict(ridge_epsilon=self.matrix_epsilon,
              relative_matrix_epsilon=self.relative_matrix_epsilon)
    if s > ops.EIGH_MAX_DIM:
      raise NotImplementedError(
          f"eigh=True supports statistics up to {ops.EIGH_MAX_DIM} x {ops.EIGH_MAX_DIM}, got {s}")
    if self.quantize_second_moment:
      q, d, b = bk.qstats
      ops.dequantize(q, d, b, True, out=bk.stats)
    if world == 1:
      roots, metrics = ops.matrix_inverse_pth_root_eigh_batched(bk.stats, bk.exps, None, **kw)
    else:
      roots, metrics = sharded_inverse_pth_roots(
          bk.stats, bk.exps, world, rank, self.process_group,
          root_fn=ops.matrix_inverse_pth_root_eigh_batched, **kw)
    self.metrics[s].copy_(metrics)
    self._select(bk, roots, metrics)

  def _lobpcg_roots(self, bk, world, rank):
    """`lobpcg_topk_precondition > 0`: deflated roots for a whole bucket (DS:789-812, DS:889-928);
    the per-statistic diagnostics rows are kept in ``bk.diagnostics``."""
    s = bk.size
    kw = dict(ridge_epsilon=self.matrix_epsilon,
              relative_matrix_epsilon=self.relative_matrix_epsilon, engine=self.engine,
              lobpcg_max_iter=self.lobpcg_max_iter)
    k = int(self.lobpcg_topk_precondition)
    if self.quantize_second_moment:
      q, d, b = bk.qstats
      ops.dequantize(q, d, b, True, out=bk.stats)
    if world == 1:
      roots, metrics, diag = ops.matrix_inverse_pth_root_lobpcg_batched(
          bk.stats, bk.exps_host, k, None, diagnostics=self.generate_training_metrics, **kw)
      bk.diagnostics = diag
    else:
      def fn(x, p, pd, **k2):
        r, m, _ = ops.matrix_inverse_pth_root_lobpcg_batched(x, p.cpu(), k, pd, diagnostics=False,
                                                             **k2)
        return r, m
      roots, metrics = sharded_inverse_pth_roots(bk.stats, bk.exps, world, rank,
                                                 self.process_group, root_fn=fn, **kw)
    self.metrics[s].copy_(metrics)
    self._select(bk, roots, metrics)

  def _refresh_low_rank(self, bk):
    """Operator of a packed low-rank preconditioner (DS:1690-1705) after the sketch changed: its
    factors (c, W = V diag(lambda^- - c)) for the two-thin-products application, and the dense
    d x d form only where a block still applies it that way (rank != 2)."""
    r = abs(self.compression_rank)
    ops.low_rank_factors(bk.packed, r, bk.loww, bk.lowc)
    if bk.needs_dense:
      ops.low_rank_to_dense(bk.packed, r, out=bk.precs)

  def _select_rows(self, new, old, metrics):
    """old[i] <- new[i] unless metrics row i reports a failed root (DS:2936-2950), for rows of
    any dtype (fp32 roots, int16 / int8 data, diagonals, bucket sizes): ``pc_select_scatter`` with
    the identity index."""
    n = int(new.shape[0])
    if n == 0:
      return
    row_bytes = new[0].numel() * new.element_size()
    key = (n, row_bytes)
    cache = self.__dict__.setdefault("_select_idx", {})
    if key not in cache:
      ar = torch.arange(n, device=self.device)
      cache[key] = ((ar * row_bytes).to(torch.int64), (ar * metrics.shape[1]).to(torch.int64),
                    ar.to(torch.int32))
    src_off, met_off, dst_idx = cache[key]
    ops.select_scatter(new.contiguous(), src_off, metrics.contiguous(), met_off, dst_idx,
                       self.inverse_failure_threshold, old, row_bytes, None)

  def _select(self, bk, roots, metrics):
    """Failure fallback of DS:2936-2950 (quantised: DS:3183-3208) for a whole bucket."""
    s = bk.size
    if self.quantize_second_moment:
      q, d, b = ops.quantize(roots, self.qdt_second, True)
      for new, old in zip((q, d, b), bk.qprecs):
        self._select_rows(new, old, metrics)
      return
    lib = _lib.load()
    _lib.check(lib.pc_select_preconditioners(
        ctypes.c_void_p(roots.data_ptr()), ctypes.c_void_p(metrics.data_ptr()),
        float(self.inverse_failure_threshold), ctypes.c_void_p(bk.precs.data_ptr()),
        bk.count, s, s, s, s, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

  def _fd_update(self, step, world, rank):
    """Sketchy branch of new_mi_pth_root (DS:2706-2738) for every compressed statistic.

    The reference runs it in the frame padded to the largest statistic (DS:2841-2843)
    and slices the packed result back to [size, rank + 2] (DS:2950): for a statistic
    smaller than that frame the eigenvalue / has_zeros slots that live in the last
    rows of the padded sketch are cut off.  Mirrored exactly by working in the same
    padded frame; when every compressed statistic has the frame size (the usual case)
    the bucket tensors are used in place."""
    r = abs(self.compression_rank)
    comp = [bk for _, bk in sorted(self.buckets.items()) if bk.compressed]
    frame = max(self.buckets)
    dev = self.device
    total = sum(bk.count for bk in comp)
    if len(comp) == 1 and comp[0].size == frame:
      bk = comp[0]
      grams, prevs, exps = bk.stats, bk.packed, bk.exps
      pads = torch.full((total,), frame, dtype=torch.int32, device=dev)
    else:
      grams = torch.zeros((total, frame, frame), dtype=torch.float32, device=dev)
      prevs = torch.zeros((total, frame, r + 2), dtype=torch.float32, device=dev)
      exps = torch.cat([bk.exps for bk in comp])
      pads = torch.cat([torch.full((bk.count,), bk.size, dtype=torch.int32, device=dev)
                        for bk in comp])
      o = 0
      for bk in comp:
        grams[o:o + bk.count, :bk.size, :bk.size] = bk.stats
        prevs[o:o + bk.count, :bk.size] = bk.packed
        o += bk.count
    if self.reset_frequency is not None and step % self.reset_frequency == 0:
      prevs = torch.zeros_like(prevs)  # DS:2140-2143
    kw = dict(ridge_epsilon=self.matrix_epsilon,
              relative_matrix_epsilon=self.relative_matrix_epsilon, decay=float(self.beta2),
              input_is_gram=True)
    if world == 1:
      new, _ = ops.fd_update_root_batched(grams.contiguous(), prevs.contiguous(), exps, r, pads,
                                          **kw)
    else:
      new = sharded_fd_updates(grams, prevs, exps, pads, r, world, rank, self.process_group, **kw)
    o = 0
    for bk in comp:
      bk.packed.copy_(new[o:o + bk.count, :bk.size])
      o += bk.count
      self._refresh_low_rank(bk)
      self.metrics[bk.size].zero_()  # DS:1263-1264: FD reports zero error

  def _low_rank_update(self, world, rank):
    """eigh-based low-rank roots (_low_rank_root, DS:1033-1120) of every compressed bucket,
    with the failure fallback of DS:2936-2950 on the packed preconditioners.  Computed at the
    bucket's own size: the zero eigenvalues the reference's pad-to-max adds are dropped by its
    own roll / flip logic (DS:1088-1098), so the packed result is the same."""
    r = self.compression_rank
    for s, bk in sorted(self.buckets.items()):
      if not bk.compressed:
        continue
      kw = dict(ridge_epsilon=self.matrix_epsilon,
                relative_matrix_epsilon=self.relative_matrix_epsilon)
      if world == 1:
        new, metrics = ops.low_rank_root_batched(bk.stats, bk.exps, r, None, **kw)
      else:
        pads = torch.full((bk.count,), s, dtype=torch.int32, device=self.device)
        new, metrics = sharded_inverse_pth_roots(
            bk.stats, bk.exps, world, rank, self.process_group, pads=pads,
            root_fn=lambda x, p, pd, **k2: ops.low_rank_root_batched(x, p, r, pd, **k2), **kw)
      self.metrics[s].copy_(metrics)
      self._select_rows(new, bk.packed, metrics)
      self._refresh_low_rank(bk)


  def _apply_preconditioners(self):
    if self.quantize_second_moment:  # DS:3556 _maybe_dequantize_preconditioners
      for bk in self.buckets.values():
        if getattr(bk, "dequant_for_apply", True):  # (tcgen05 consumers read the quantised form)
          q, d, b = bk.qprecs
          ops.dequantize(q, d, b, True, out=bk.precs)
    for j, simt in enumerate(self._apply_simt):
      if simt is not None:
        simt.run()
      if self._apply_tc[j] is not None:
        self._apply_tc[j].run()
    if self._staged:
      self._stage_out()

  def _transform_all(self, step, lr):
    """Grafting + momentum tail of every parameter (DS:3496-3625) as ONE grouped call over the
    flat buffers; returns the flat update buffer (a fresh one per step: the returned updates
    stay valid after the next ``update``)."""
    key = (step >= self.start_preconditioning_step, lr)
    if getattr(self, "_graft_opt_key", None) != key:  # the options are the same for every parameter
      self._graft_opt_key = key
      self._graft_opt = ops.make_graft_options(
          beta1=float(self.beta1), beta2=float(self.beta2), graft_type=int(self.graft_type),
          diagonal_epsilon=float(self.diagonal_epsilon), weight_decay=float(self.weight_decay),
          learning_rate=lr, nesterov=int(bool(self.nesterov)),
          moving_average_for_momentum=int(bool(self.moving_average_for_momentum)),
          decoupled_learning_rate=int(bool(self.decoupled_learning_rate)),
          decoupled_weight_decay=int(bool(self.decoupled_weight_decay)),
          run_shampoo=int(step >= self.start_preconditioning_step),
          clip_by_scaled_gradient_norm=float(self.clip_by_scaled_gradient_norm or 0.0))
    quantised = [(pl, st) for pl, st in zip(self.plans, self._own_leaves)
                 if pl.mdt != torch.float32]
    int8 = [(pl, st) for pl, st in quantised if pl.mdt == torch.int8]
    other = [(pl, st) for pl, st in quantised if pl.mdt != torch.int8]
    # int8 momenta (the usual best_effort_memory_usage_reduction case): every to_float /
    # from_float of the model in one grouped call each; the segment table is static
    ptrs = tuple(qv.quantized.data_ptr() for _, st in int8
                 for qv in (st.momentum, st.diagonal_momentum))
    if getattr(self, "_qgroup_key", None) != ptrs:
      self._qgroup_key = ptrs
      items = []
      for pl, st in int8:
        seg = slice(pl.offset, pl.offset + pl.numel)
        for qv, buf in ((st.momentum, self.mbuf), (st.diagonal_momentum, self.dmbuf)):
          items.append((qv.quantized, qv.bucket_size, buf[seg]))
      self._qgroup = ops.QuantGroup(items, self.device) if items else None
    if self._qgroup is not None:
      self._qgroup.dequantize()  # to_float into the flat scratch (DS:3582-3586)
    for pl, st in other:
      seg = slice(pl.offset, pl.offset + pl.numel)
      for qv, buf in ((st.momentum, self.mbuf), (st.diagonal_momentum, self.dmbuf)):
        buf[seg].copy_(qv.to_float().reshape(-1))
    ubuf = torch.empty(self.total, dtype=torch.float32, device=self.device)
    self._graft_group.run(self.gbuf, self.pbuf, self.pgbuf, self.dsbuf, self.dmbuf, self.mbuf,
                          ubuf, self._graft_opt)
    if self._qgroup is not None:
      self._qgroup.quantize()  # requantise (DS:3620-3621)
    for pl, st in other:
      seg = slice(pl.offset, pl.offset + pl.numel)
      for qv, buf in ((st.momentum, self.mbuf), (st.diagonal_momentum, self.dmbuf)):
        q, _, b = QuantizedValue.quantize(buf[seg].view(pl.shape), pl.mdt)
        qv.quantized.copy_(q)
        qv.bucket_size.copy_(b)
    return ubuf


def sharded_inverse_pth_roots(stats, exps, world, rank, group, root_fn=None, pads=None, **kw):
  """Block-sharded roots + all-gather, the device boundary of DS:2841-2879.

  The batch is padded to a multiple of ``world`` with (I, exponent 1, padding 0)
  fillers (DS:2844-2850), rank r computes contiguous chunk r (DS:1827-1831,
  DS:2869-2875), roots and metrics are all-gathered (DS:2876-2877) and the
  fillers dropped (unbatch, DS:1834-1846).  ``root_fn`` is injectable so the
  exchange logic is testable on CPU with gloo."""
  import torch.distributed as dist
  root_fn = root_fn or ops.matrix_inverse_pth_root_batched
  n_stats, s = stats.shape[0], stats.shape[1]
  to_pad = -n_stats % world
  b = (n_stats + to_pad) // world
  lo, hi = rank * b, min((rank + 1) * b, n_stats)
  local = torch.eye(s, dtype=stats.dtype, device=stats.device).repeat(b, 1, 1)
  local_ps = torch.ones(b, dtype=torch.int32, device=stats.device)
  local_pad = torch.zeros(b, dtype=torch.int32, device=stats.device)
  if hi > lo:
    local[:hi - lo] = stats[lo:hi]
    local_ps[:hi - lo] = exps[lo:hi]
    local_pad[:hi - lo] = s if pads is None else pads[lo:hi]
  roots, metrics = root_fn(local.contiguous(), local_ps, local_pad, **kw)
  all_roots = torch.empty((world * b,) + tuple(roots.shape[1:]), dtype=roots.dtype,
                          device=roots.device)  # [.., s, s] roots or [.., s, rank + 2] packed
  all_metrics = torch.empty((world * b, metrics.shape[1]), dtype=metrics.dtype,
                            device=metrics.device)
  dist.all_gather_into_tensor(all_roots, roots.contiguous(), group=group)
  dist.all_gather_into_tensor(all_metrics, metrics.contiguous(), group=group)
  return all_roots[:n_stats], all_metrics[:n_stats]


def sharded_fd_updates(grams, prevs, exps, pads, r, world, rank, group, fd_fn=None, **kw):
  """Block-sharded Sketchy updates + all-gather: the same contiguous partition and filler
  rule as the roots (DS:2844-2850, DS:2862-2877); fillers are (zero, exponent 1, padding 0)
  and produce all-zero sketches (DS:1265-1268)."""
  import torch.distributed as dist
  fd_fn = fd_fn or ops.fd_update_root_batched
  n_stats, d = grams.shape[0], grams.shape[1]
  b = (n_stats + (-n_stats % world)) // world
  lo, hi = rank * b, min((rank + 1) * b, n_stats)
  local_g = torch.zeros((b, d, d), dtype=grams.dtype, device=grams.device)
  local_p = torch.zeros((b, d, r + 2), dtype=prevs.dtype, device=grams.device)
  local_e = torch.ones(b, dtype=torch.int32, device=grams.device)
  local_pad = torch.zeros(b, dtype=torch.int32, device=grams.device)
  if hi > lo:
    local_g[:hi - lo] = grams[lo:hi]
    local_p[:hi - lo] = prevs[lo:hi]
    local_e[:hi - lo] = exps[lo:hi]
    local_pad[:hi - lo] = pads[lo:hi]
  new, _ = fd_fn(local_g, local_p, local_e, r, local_pad, **kw)
  gathered = torch.empty((world * b, d, r + 2), dtype=new.dtype, device=new.device)
  dist.all_gather_into_tensor(gathered, new.contiguous(), group=group)
  return gathered[:n_stats]




class ShampooTransformation(GradientTransformation):
  """``GradientTransformation(init, update)`` plus the state hand-over of this implementation:
  ``export_state(state)`` (detached deep copy, e.g. for a checkpoint) and
  ``import_state(state)`` (copy a foreign state into the optimizer's device buffers)."""
  pass


def distributed_shampoo(
    learning_rate,
    block_size,
    beta1=0.9,
    beta2=0.999,
    diagonal_epsilon=1e-10,
    matrix_epsilon=1e-6,
    weight_decay=0.0,
    start_preconditioning_step=5,
    preconditioning_compute_steps=1,
    decay_preconditioning_compute_steps: bool = False,
    end_preconditioning_compute_steps: Optional[int] = None,
    statistics_compute_steps=1,
    best_effort_shape_interpretation=True,
    graft_type=GraftingType.SGD,
    nesterov=True,
    exponent_override=0,
    batch_axis_name=None,
    statistics_partition_spec=None,
    preconditioner_partition_spec=None,
    num_devices_for_pjit=None,
    shard_optimizer_states=False,
    best_effort_memory_usage_reduction=False,
    inverse_failure_threshold=0.1,
    moving_average_for_momentum=False,
    skip_preconditioning_dim_size_gt=4096,
    clip_by_scaled_gradient_norm=None,
    precision=None,
    tensordot_precision=None,
    relative_matrix_epsilon=True,
    merge_small_dims_block_size=4096,
    lobpcg_topk_precondition: int = 0,
    lobpcg_max_iter: int = 0,
    precondtioner_type=PreconditionerType.ALL,
    generate_fd_metrics: bool = False,
    compression_rank: int = 0,
    frequent_directions: bool = False,
    reset_preconditioner: bool = False,
    average_grad: bool = False,
    skip_preconditioning_rank_lt=1,
    decoupled_learning_rate=True,
    decoupled_weight_decay=False,
    generate_training_metrics=True,
    reuse_preconditioner=False,
    eigh=False,
    # --- B200 extensions (not in the reference) ---
    engine: int = _lib.PC_ENGINE_AUTO,
    process_group=None,
):
  """Distributed Shampoo (keyword surface of DS:1849-1900).

  ``batch_axis_name`` truthy == the reference's pmap mode: preconditioner blocks
  are partitioned over the ranks of ``process_group`` (default world) of an
  initialised ``torch.distributed`` NCCL group and the roots are all-gathered.
  ``precision`` / ``tensordot_precision`` are accepted and ignored: GEMMs are fp32
  (CUDA cores) or fp32-accurate split-fp16 / split-bf16 (tcgen05).

  State semantics: ``init`` returns a ``ShampooState`` whose tensors are views of the
  optimizer's device buffers and ``update`` advances them in place (the state it returns
  aliases the same buffers).  Passing any OTHER state to ``update`` -- a restored checkpoint, an
  ``export_state`` copy, a deep copy -- is supported: it is copied into the buffers first.
  The returned object also carries ``export_state(state)`` and ``import_state(state)``.
  """
  del precision, tensordot_precision
  if generate_fd_metrics and frequent_directions:  # DS:2026: ignored without frequent_directions
    raise NotImplementedError(
        "generate_fd_metrics (FDDiagnostics, DS:197-335) is not built in the B200 hot path")
  if reset_preconditioner and not frequent_directions:  # DS:2019-2020
    raise ValueError("reset_preconditioner=True requries frequent_directions")
  reset_frequency = None
  if reset_preconditioner:  # DS:2022-2024
    reset_frequency = int(np.round(1 / (1 - beta2))) if beta2 != 1 else None
    beta2 = 1.0
  if frequent_directions and compression_rank <= 0:  # DS:2028-2030
    raise ValueError("frequent_directions=True requires compression_rank > 0,"
                     f" found {compression_rank}")
  if average_grad and not frequent_directions:  # DS:2032-2033
    raise ValueError("average_grad requested but frequent_directions is False")
  if frequent_directions and (statistics_compute_steps != preconditioning_compute_steps):
    raise ValueError("frequent_directions=True requires "
                     f"statistics_compute_steps ({statistics_compute_steps}) "
                     "to equal != preconditioning_compute_steps "
                     f"({preconditioning_compute_steps})")
  if exponent_override and not 1 <= int(exponent_override) <= 16:
    raise ValueError(f"exponent_override={exponent_override} is outside [1, 16], the range of "
                     "the Newton step programs (pc_inverse_pth_root_batched)")
  if frequent_directions and not reuse_preconditioner:
    # _fd_update_root asserts that the previous sketch is passed in (DS:1150)
    raise ValueError("frequent_directions=True needs reuse_preconditioner=True")
  if frequent_directions and best_effort_memory_usage_reduction and batch_axis_name:
    pass  # DS:2051-2064: no second-moment quantisation with compression_rank != 0
  opt = _Shampoo(learning_rate, block_size, beta1, beta2, diagonal_epsilon, matrix_epsilon,
                 weight_decay, start_preconditioning_step, preconditioning_compute_steps,
                 statistics_compute_steps, best_effort_shape_interpretation,
                 GraftingType(graft_type), nesterov, exponent_override, batch_axis_name,
                 best_effort_memory_usage_reduction, inverse_failure_threshold,
                 moving_average_for_momentum, skip_preconditioning_dim_size_gt,
                 clip_by_scaled_gradient_norm, relative_matrix_epsilon,
                 merge_small_dims_block_size, PreconditionerType(precondtioner_type),
                 compression_rank, skip_preconditioning_rank_lt, decoupled_learning_rate,
                 decoupled_weight_decay, generate_training_metrics, engine, process_group,
                 frequent_directions, reuse_preconditioner, reset_frequency, average_grad,
                 bool(eigh), bool(decay_preconditioning_compute_steps),
                 end_preconditioning_compute_steps, int(lobpcg_topk_precondition),
                 int(lobpcg_max_iter), bool(shard_optimizer_states), num_devices_for_pjit)
  if shard_optimizer_states:
    # DS:3660-3673: `init` hands back the init / partition-spec / shape functions of the sharded
    # state; `opt.init(params).init_fn(params)` builds ShampooState(count,
    # ShardedShampooStats(global_stats, local_stats)) and `update` is sharded_update_fn.
    if not batch_axis_name:
      opt.batch_axis_name = "pjit"  # the device axis of the sharded state

    def pspec_fn(params, params_partition_spec=None, partition_spec_for_statistics=None):
      """Partition specs of the state (DS:2277-2330): leading axis of the global arrays sharded
      (`statistics_partition_spec`), local statistics like their parameters."""
      del params, params_partition_spec
      stat = partition_spec_for_statistics or statistics_partition_spec
      return ShampooState(count=(), stats=ShardedShampooStats(
          GlobalShardedParameterStats(stat, preconditioner_partition_spec or stat, ()), None))

    def shape_and_dtype_fn(params):
      """Shapes / dtypes of the global arrays without allocating them (DS:2332-2416)."""
      leaves, _ = _tree_flatten(params)
      n, mx = 0, 0
      for p in leaves:
        if not opt._skip_preconditioning(p.shape):
          pre = Preconditioner(p.shape, block_size, merge_small_dims_block_size,
                               best_effort_shape_interpretation,
                               PreconditionerType(precondtioner_type), 0)
          sizes = [sh[0] for sh in pre.shapes_for_preconditioners()]
          n, mx = n + len(sizes), max([mx] + sizes)
      ndev = int(num_devices_for_pjit or 1)
      n += -n % ndev
      return ShampooState(count=((), torch.int32), stats=ShardedShampooStats(
          GlobalShardedParameterStats(((n, mx, mx), torch.float32), ((n, mx, mx), torch.float32),
                                      ((n,), torch.int32)), None))

    def _init_fns(unused_params):
      return InitFnState(init_fn=opt.init, pspec_fn=pspec_fn,
                         shape_and_dtype_fn=shape_and_dtype_fn)

    tx = ShampooTransformation(_init_fns, opt.update)
  else:
    tx = ShampooTransformation(opt.init, opt.update)
  tx.export_state, tx.import_state = opt.export_state, opt.import_state
  return tx
