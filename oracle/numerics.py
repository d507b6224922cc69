"""numpy restatement of the reference's numerical kernels (TEST INFRASTRUCTURE).

See ``oracle/__init__.py`` for scope and pinning status.  ``DS:a-b`` cites lines
a..b of ``precondition/distributed_shampoo.py`` in the reference; ``QU`` is
``precondition/quantization_utils.py``.

All routines take a ``dtype`` (default float32 = what the reference actually
runs in when ``jax_enable_x64`` is off, DS:35-38, DS:773); pass ``np.float64`` for
the ground-truth twin used in residual checks.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Tuple

import numpy as np

_EPSILON = 1e-25  # DS:41


# --------------------------------------------------------------------------
# metrics container (DS:338-351, the five scalars the kernel must emit)
# --------------------------------------------------------------------------
@dataclasses.dataclass
class RootMetrics:
  inverse_pth_root_errors: float = 0.0
  inverse_pth_root_iters: float = 0.0
  final_error_ratio: float = 0.0
  max_eigen_value: float = 0.0
  total_retries: float = 0.0

  def as_row(self) -> np.ndarray:
    return np.array([
        self.inverse_pth_root_errors, self.inverse_pth_root_iters,
        self.final_error_ratio, self.max_eigen_value, self.total_retries
    ], dtype=np.float32)


# --------------------------------------------------------------------------
# power iteration  (DS:595-652)
# --------------------------------------------------------------------------
def power_iteration_start_vector(n: int, dtype=np.float32) -> np.ndarray:
  """Fixed start vector of the reference, DS:642-643."""
  return np.random.RandomState(1729).uniform(-1.0, 1.0, n).astype(dtype)


def power_iteration(matrix: np.ndarray,
                    num_iters: int = 100,
                    error_tolerance: float = 1e-6,
                    padding_start: Optional[int] = None,
                    return_iters: bool = False):
  """Largest eigenvalue by power iteration, DS:595-652.

  Returns (v, s) where ``s`` is the Rayleigh quotient computed in the LAST
  executed step (DS:637-638) and ``v`` the normalised iterate (DS:651).
  """
  dtype = matrix.dtype
  n = matrix.shape[-1]
  v = power_iteration_start_vector(n, dtype)
  if padding_start is not None:
    v = v * (np.arange(n, dtype=np.int32) < padding_start).astype(dtype)
  s = dtype.type(0)
  tol = dtype.type(error_tolerance)
  i = 0
  run_step = True
  while i < num_iters and run_step:  # DS:627-629
    nv = v / np.linalg.norm(v)  # DS:634
    s_v = np.einsum("ij,j->i", matrix, nv)  # DS:636
    s_new = np.einsum("i,i->", nv, s_v)  # DS:637
    run_step = bool(np.abs(s_new - s) > tol)  # DS:639
    v, s, i = s_v, s_new, i + 1
  v_out = v / np.linalg.norm(v)
  if return_iters:
    return v_out, s, i
  return v_out, s


# --------------------------------------------------------------------------
# mat_power  (DS:655-678)
# --------------------------------------------------------------------------
def mat_power(mat_m: np.ndarray, p: int) -> np.ndarray:
  """M^p by LSB-first binary powering, in the reference's multiply order.

  The reference multiplies ``mat @ power`` starting from ``power = I`` and
  always squares once more than needed (DS:667-675); multiplying by the exact
  identity and the discarded final square do not change the result, so they
  are skipped here -- the surviving products are issued in the same order with
  the same operands, hence identical rounding.
  """
  power = None
  mat = mat_m
  i = int(p)
  while i > 0:
    if i % 2 == 1:
      power = mat if power is None else mat @ power  # DS:670-672
    i //= 2
    if i > 0:
      mat = mat @ mat  # DS:674
  if power is None:  # p == 0
    power = np.eye(mat_m.shape[0], dtype=mat_m.dtype)
  return power


def mat_power_literal(mat_m: np.ndarray, p: int) -> np.ndarray:
  """DS:655-678 with every product the reference issues (the multiply by the
  identity and the discarded last squaring included) -- same values as
  ``mat_power``; used when timing the reference algorithm on the CPU."""
  power = np.eye(mat_m.shape[0], dtype=mat_m.dtype)
  mat = mat_m
  i = int(p)
  while i > 0:
    if i % 2 == 1:
      power = mat @ power
    i //= 2
    mat = mat @ mat
  return power


# --------------------------------------------------------------------------
# matrix_inverse_pth_root, coupled Newton branch  (DS:702-940)
# --------------------------------------------------------------------------
def matrix_inverse_pth_root(
    matrix: np.ndarray,
    p: int,
    num_iters: int = 100,
    ridge_epsilon: float = 1e-6,
    error_tolerance: float = 1e-6,
    relative_matrix_epsilon: bool = True,
    padding_start: Optional[int] = None,
    dtype=np.float32,
    trace: Optional[list] = None,
    literal_mat_power: bool = False,
) -> Tuple[np.ndarray, RootMetrics]:
  """(A + eps I)^(-1/p) by the coupled Newton iteration, DS:702-940.

  Default branch only (``lobpcg_topk_precondition=0``, ``eigh=False``).
  ``trace``, if a list, receives (try, iter, error) tuples for diagnostics.
  """
  dtype = np.dtype(dtype)
  f = dtype.type
  assert matrix.shape[0] == matrix.shape[1]
  n = matrix.shape[0]
  orig_dtype = matrix.dtype
  p = int(p)
  a = matrix.astype(dtype)  # DS:773
  alpha = f(-1.0 / p)  # DS:774
  identity = np.eye(n, dtype=dtype)
  if padding_start is not None:  # DS:777-783
    ix = (np.arange(n, dtype=np.int32) < padding_start).astype(dtype)
    a = a * ix[np.newaxis, :]
    a = a * ix[:, np.newaxis]
    identity = identity * ix

  if relative_matrix_epsilon:  # DS:814-828
    _, max_ev = power_iteration(
        a, num_iters=100, error_tolerance=1e-6, padding_start=padding_start)
  else:
    max_ev = f(1.0)
  ridge = f(ridge_epsilon) * np.maximum(max_ev, f(_EPSILON))  # DS:830
  max_error_ratio = f(1.2)  # DS:834
  tol = f(error_tolerance)

  if n == 1:  # DS:850-855 (total_retries treated as 0, see SURVEY 8(a'))
    h = (a + ridge)**alpha
    error, iters, error_ratio, total_retries = f(0), 0, f(0), 0
  else:
    total_retries = 0
    h = identity
    error, iters, error_ratio, failed = f(1000.0), 100, f(1.0), True  # DS:860
    while failed and total_retries < 6:  # DS:862-864
      damped = a + (ridge * f(10**total_retries)) * identity  # DS:869
      z = f(1 + p) / (f(2) * np.linalg.norm(damped).astype(dtype))  # DS:870
      mat_m = damped * z  # DS:871
      err = np.max(np.abs(mat_m - identity))  # DS:872
      mat_h = identity * np.power(z, f(1.0 / p))  # DS:873
      old_h = mat_h
      ratio = f(1.0)
      i = 0
      while i < num_iters and err > tol and ratio < max_error_ratio:  # DS:836-840
        mat_m_i = (f(1) - alpha) * identity + alpha * mat_m  # DS:844
        new_m = (mat_power_literal if literal_mat_power else mat_power)(
            mat_m_i, p) @ mat_m  # DS:845
        new_h = mat_h @ mat_m_i  # DS:846
        new_err = np.max(np.abs(new_m - identity))  # DS:847
        ratio = new_err / err
        mat_m, old_h, mat_h, err = new_m, mat_h, new_h, new_err
        i += 1
        if trace is not None:
          trace.append((total_retries, i, float(err)))
      error = np.max(np.abs(mat_m - identity)).astype(np.float32)  # DS:878
      # DS:879-880 is an arithmetic blend, not a select: a non-finite entry in
      # either matrix poisons the result (0 * inf = nan), exactly as upstream.
      conv = f(1.0) if ratio < max_error_ratio else f(0.0)
      h = conv * mat_h + (f(1) - conv) * old_h
      iters, error_ratio = i, ratio
      failed = bool(error > 0.05)  # DS:858, DS:882
      total_retries += 1

  metrics = RootMetrics(
      inverse_pth_root_errors=float(error),
      inverse_pth_root_iters=float(iters),
      final_error_ratio=float(error_ratio),
      max_eigen_value=float(max_ev),
      total_retries=float(total_retries))
  if padding_start is not None and padding_start == 0:  # DS:930-937
    h = np.zeros_like(h)
    metrics.inverse_pth_root_errors = 0.0
  return np.asarray(h, dtype=orig_dtype), metrics


def matrix_inverse_pth_root_batched(xs, ps, padding_starts=None, **kw):
  """vmap semantics of DS:2742-2744: every matrix behaves as if run alone."""
  roots, rows = [], []
  for b in range(len(xs)):
    pad = None if padding_starts is None else int(padding_starts[b])
    r, m = matrix_inverse_pth_root(xs[b], int(ps[b]), padding_start=pad, **kw)
    roots.append(r)
    rows.append(m.as_row())
  return np.stack(roots), np.stack(rows)


def root_residual(root: np.ndarray, matrix: np.ndarray, p: int,
                  eps: float) -> float:
  """max-abs of X^p (A + eps I) - I evaluated in float64 (ground truth)."""
  x = root.astype(np.float64)
  a = matrix.astype(np.float64) + eps * np.eye(matrix.shape[0])
  return float(np.max(np.abs(np.linalg.matrix_power(x, p) @ a -
                             np.eye(matrix.shape[0]))))


def exact_inverse_pth_root(matrix: np.ndarray, p: int, eps: float) -> np.ndarray:
  """float64 eigh ground truth of (A + eps I)^(-1/p)."""
  a = matrix.astype(np.float64)
  w, v = np.linalg.eigh((a + a.T) / 2 + eps * np.eye(a.shape[0]))
  return (v * np.power(np.maximum(w, 1e-300), -1.0 / p)) @ v.T


# --------------------------------------------------------------------------
# padding helpers  (DS:1324-1369)
# --------------------------------------------------------------------------
def pad_square_matrix(mat: np.ndarray, max_size: int) -> np.ndarray:
  """[[M, 0], [0, I]] of size max_size, DS:1324-1350."""
  rows, cols = mat.shape
  if rows != cols:
    raise ValueError(f"Must have rows == cols, instead got rows={rows}, cols={cols}")
  if cols > max_size:
    raise ValueError(
        f"Must have cols <= max_size. Instead got cols={cols}, max_size={max_size}.")
  if rows == max_size:
    return mat
  out = np.eye(max_size, dtype=mat.dtype)
  out[:rows, :rows] = mat
  return out


def pad_vector(vec: np.ndarray, max_size: int) -> np.ndarray:
  """[V, 0], DS:1353-1369."""
  assert vec.shape[0] <= max_size
  out = np.zeros([max_size], dtype=vec.dtype)
  out[:vec.shape[0]] = vec
  return out


# --------------------------------------------------------------------------
# QuantizedValue  (QU:25-113)
# --------------------------------------------------------------------------
_NUM_BUCKETS = {np.dtype(np.int8): 127.0, np.dtype(np.int16): 32767.0}


def quantize(fvalue: np.ndarray, quantized_dtype, extract_diagonal=False):
  """QU:49-95.  Returns (quantized, diagonal, bucket_size)."""
  qd = "bfloat16" if quantized_dtype == "bfloat16" else np.dtype(quantized_dtype)
  if qd == np.dtype(np.float32):
    return fvalue, None, None  # QU:52-53
  if qd == "bfloat16":  # QU:54-55: plain cast (round-to-nearest-even)
    return to_bfloat16_bits(fvalue), None, None
  if qd not in _NUM_BUCKETS:
    raise ValueError(f"Quantized dtype {quantized_dtype} not supported.")
  fdtype = fvalue.dtype
  num_buckets = fdtype.type(_NUM_BUCKETS[qd])
  if extract_diagonal and fvalue.ndim != 2:
    raise ValueError("Input array must be 2D to work with extract_diagonal.")
  diagonal = None
  if extract_diagonal:  # QU:72-76
    diagonal = np.diag(fvalue).copy()
    fvalue = fvalue - np.diag(diagonal)
  if fvalue.ndim < 1:
    raise ValueError("Input array must have a strictly positive number of dimensions.")
  max_abs = np.max(np.abs(fvalue), axis=0)  # QU:86
  bucket_size = max_abs / num_buckets  # QU:87
  bs = bucket_size[np.newaxis, ...]
  bs_nonzero = np.where(bs > 0.0, bs, np.ones_like(bs))  # QU:90-91
  ratio = fvalue / bs_nonzero
  quantized = np.round(ratio)  # QU:94 (round half to even, like jnp.round)
  return quantized.astype(qd), diagonal, bucket_size


def dequantize(quantized, diagonal, bucket_size, quantized_dtype,
               extract_diagonal=False) -> np.ndarray:
  """QU:97-113."""
  qd = "bfloat16" if quantized_dtype == "bfloat16" else np.dtype(quantized_dtype)
  if qd == np.dtype(np.float32):
    return quantized
  if qd == "bfloat16":
    return from_bfloat16_bits(quantized)
  val = quantized.astype(bucket_size.dtype) * bucket_size[np.newaxis, ...]
  if extract_diagonal:
    val = val + np.diag(diagonal)
  return val


def to_bfloat16_bits(x: np.ndarray) -> np.ndarray:
  """float32 -> bfloat16 (round to nearest even), returned as uint16 bits."""
  u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
  rounding = ((u >> 16) & 1) + np.uint32(0x7FFF)
  out = ((u + rounding) >> 16).astype(np.uint16)
  nan = np.isnan(x)
  if np.any(nan):
    out = np.where(nan, np.uint16(0x7FC0), out)
  return out


def from_bfloat16_bits(b: np.ndarray) -> np.ndarray:
  return (b.astype(np.uint32) << 16).view(np.float32)


@dataclasses.dataclass
class QuantizedValue:
  """Mirror of QU:25-47 (numpy)."""
  quantized: object
  diagonal: object
  bucket_size: object
  quantized_dtype: object
  extract_diagonal: bool
  shape: object

  @classmethod
  def from_float_value(cls, fvalue, quantized_dtype, extract_diagonal=False):
    if isinstance(fvalue, list) and not fvalue:
      return cls([], [], [], quantized_dtype, extract_diagonal, [])
    q, d, b = quantize(fvalue, quantized_dtype, extract_diagonal)
    return cls(q, d, b, quantized_dtype, extract_diagonal, list(q.shape))

  def to_float(self):
    if isinstance(self.quantized, list) and not self.quantized:
      return self.quantized
    return dequantize(self.quantized, self.diagonal, self.bucket_size,
                      self.quantized_dtype, self.extract_diagonal)


# --------------------------------------------------------------------------
# statistics updates  (DS:1440-1505)
# --------------------------------------------------------------------------
def gram_weighted_update(old_stats, g, axis, w1, w2):
  """w1*S + w2*tensordot(g, g) over all axes but ``axis``, DS:1440-1470."""
  axes = [i for i in range(g.ndim) if i != axis]
  gram = np.tensordot(g, g, axes=(axes, axes))
  f = g.dtype.type
  return f(w1) * old_stats + f(w2) * gram


def frequent_directions_update(old_stats_factor, g, axis, w1, w2):
  """Square factor R with R R^T = x x^T, DS:1473-1505 (QR based)."""
  del old_stats_factor, w1, w2
  x = np.reshape(np.moveaxis(g, axis, 0), (g.shape[axis], -1))
  r = np.linalg.qr(x.T, mode="r").T
  assert r.shape == (x.shape[0], min(x.shape))
  return np.pad(r, ((0, 0), (0, x.shape[0] - r.shape[1])))


# --------------------------------------------------------------------------
# low-rank packing  (DS:520-592)
# --------------------------------------------------------------------------
def precond_dim(compression_rank: int, dim: int) -> int:
  """DS:520-532."""
  if not compression_rank:
    return dim
  compressed = abs(compression_rank) + 2
  return dim if compressed >= dim else compressed


def should_compress(compression_rank: int, dim: int) -> bool:
  """DS:535-537."""
  return compression_rank != 0 and abs(compression_rank) + 2 < dim


def fd_low_rank_unpack(precond: np.ndarray, compression_rank: int):
  """DS:555-569 -> (eigvecs, eigvals, inverted_eigvals, const, tail, has_zeros)."""
  r = abs(compression_rank)
  dim, storage = precond.shape
  assert storage < dim and storage == r + 2
  return (precond[:, :r], precond[-r:, -1], precond[:r, -2], precond[0, -1],
          precond[1, -1], bool(precond[-1, -2]))


def fd_low_rank_pack(eigvecs, deflated_eigs, inverted_eigs, new_const, new_tail,
                     has_zeros, rank):
  """DS:572-592 (write order matters where slots overlap)."""
  rank = abs(rank)
  d = eigvecs.shape[0]
  assert eigvecs.shape == (d, rank)
  assert precond_dim(rank, d) == rank + 2 < d
  out = np.zeros((d, rank + 2), dtype=np.float32)
  out[:, :rank] = eigvecs
  out[:rank, -2] = inverted_eigs
  out[0, -1] = new_const
  out[1, -1] = new_tail
  out[-rank:, -1] = deflated_eigs
  out[-1, -2] = np.float32(bool(has_zeros))
  return out


def low_rank_unpack(precond, compression_rank):
  """DS:540-545 -> (eigvecs, inverted_eigvals, const, has_zeros)."""
  vecs, _, inv, const, _, hz = fd_low_rank_unpack(precond, compression_rank)
  return vecs, inv, const, hz


def low_rank_pack(eigvecs, eigvals, const, compression_rank):
  """DS:548-552."""
  return fd_low_rank_pack(eigvecs, np.zeros_like(eigvals), eigvals, const, 0.0,
                          False, compression_rank)


# --------------------------------------------------------------------------
# eigh-based low-rank root  (DS:1033-1120)
# --------------------------------------------------------------------------
def low_rank_root(matrix, p, compression_rank, ridge_epsilon=1e-6,
                  error_tolerance=1e-6, relative_matrix_epsilon=True,
                  padding_start=None, dtype=np.float32):
  dtype = np.dtype(dtype)
  f = dtype.type
  assert compression_rank != 0
  d = matrix.shape[0]
  assert d > abs(compression_rank) + 2
  orig_dtype = matrix.dtype
  a = matrix.astype(dtype)
  alpha = f(-1.0 / p)
  identity = np.eye(d, dtype=dtype)
  ix = None
  if padding_start is not None:
    ix = (np.arange(d, dtype=np.int32) < padding_start).astype(dtype)
    a = a * ix[np.newaxis, :] * ix[:, np.newaxis]
    identity = identity * ix
  if relative_matrix_epsilon:
    _, max_ev = power_iteration(a, 100, error_tolerance, padding_start)
  else:
    max_ev = f(1.0)
  ridge = f(ridge_epsilon) * np.maximum(max_ev, f(error_tolerance))  # DS:1069
  reg = a + ridge * identity
  e, u = np.linalg.eigh(reg)
  e = e.astype(dtype)
  u = u.astype(dtype)
  if ix is not None:
    e = e * np.flip(ix)
  recovered = u.T @ (reg @ u)
  eig_error = recovered - np.diag(e)
  if ix is not None:
    eig_error = eig_error * np.flip(ix)
  error = np.max(np.abs(eig_error))
  with np.errstate(divide="ignore"):
    inv_e = np.where(e == 0.0, f(0), np.power(np.maximum(e, ridge), alpha))
  real_dim = padding_start if padding_start is not None else d
  if compression_rank < 0:  # DS:1088-1094
    inv_e = np.roll(inv_e, -(d - real_dim))
    u = np.roll(u, -(d - real_dim), axis=1)
  else:  # DS:1095-1098
    inv_e = np.flip(inv_e)
    u = np.flip(u, axis=1)
  k = abs(compression_rank)
  keep_e, to_avg = inv_e[:k], inv_e[k:]
  n_avg = real_dim - k
  const = np.sum(to_avg) / (f(n_avg) if n_avg > 0 else f(1.0))
  val = low_rank_pack(u[:, :k], keep_e, const, compression_rank)
  metrics = RootMetrics(inverse_pth_root_errors=float(error))
  if padding_start is not None and padding_start == 0:
    val = np.zeros_like(val)
    metrics.inverse_pth_root_errors = 0.0
  return val.astype(orig_dtype), metrics


# --------------------------------------------------------------------------
# eigh-based full root  (matrix_inverse_pth_root_eigh, DS:943-1030; `eigh=True`)
# --------------------------------------------------------------------------
def matrix_inverse_pth_root_eigh(matrix, p, ridge_epsilon=1e-6, error_tolerance=1e-6,
                                 relative_matrix_epsilon=True, padding_start=None,
                                 dtype=np.float32):
  dtype = np.dtype(dtype)
  f = dtype.type
  d = matrix.shape[0]
  orig_dtype = matrix.dtype
  a = matrix.astype(dtype)
  alpha = f(-1.0 / p)
  identity = np.eye(d, dtype=dtype)
  ix = None
  if padding_start is not None:  # DS:992-997
    ix = (np.arange(d, dtype=np.int32) < padding_start).astype(dtype)
    a = a * ix[np.newaxis, :] * ix[:, np.newaxis]
    identity = identity * ix
  if relative_matrix_epsilon:  # DS:998-1004
    _, max_ev = power_iteration(a, 100, error_tolerance, padding_start)
  else:
    max_ev = f(1.0)
  ridge = f(ridge_epsilon) * np.maximum(max_ev, f(error_tolerance))  # DS:1008
  reg = a + ridge * identity
  if padding_start is not None and padding_start == 0:  # DS:1026-1030 (reg is NaN here)
    return np.zeros_like(matrix), RootMetrics(inverse_pth_root_errors=0.0)
  e, u = np.linalg.eigh(reg)
  e, u = e.astype(dtype), u.astype(dtype)
  if ix is not None:
    e = e * np.flip(ix)  # DS:1012-1013
  with np.errstate(divide="ignore"):
    inv_e = np.where(e == 0.0, f(0), np.power(np.maximum(e, ridge), alpha))  # DS:1015-1016
  root = u * np.sqrt(inv_e)
  val = root @ root.T  # DS:1018-1019
  eig_error = u.T @ (reg @ u) - np.diag(e)  # DS:1020-1021
  if ix is not None:
    eig_error = eig_error * np.flip(ix)
  error = np.max(np.abs(eig_error))
  metrics = RootMetrics(inverse_pth_root_errors=float(error))
  if padding_start is not None and padding_start == 0:  # DS:1026-1030
    val = np.zeros_like(val)
    metrics.inverse_pth_root_errors = 0.0
  return val.astype(orig_dtype), metrics


# --------------------------------------------------------------------------
# Sketchy / frequent-directions sketch update  (DS:1123-1290)
# --------------------------------------------------------------------------
def fd_update_root(new_grad, p, rank, ridge_epsilon=1e-6, error_tolerance=1e-6,
                   relative_matrix_epsilon=True, decay=1.0, padding_start=None,
                   prev=None):
  """One FD step on the packed [d, rank+2] sketch, DS:1123-1290."""
  assert prev is not None and rank > 0
  d = new_grad.shape[0]
  assert new_grad.shape == (d, d)
  pd = precond_dim(rank, d)
  assert prev.shape == (d, pd) and rank + 2 == pd < d
  f = np.float32
  sketch, fwd_eigs, _, _, tail, _ = fd_low_rank_unpack(prev.astype(f), rank)
  sketch = sketch.copy()
  fwd_eigs = fwd_eigs.copy()
  max_ev = fwd_eigs[0] if relative_matrix_epsilon else f(1.0)  # DS:1155-1158
  ridge = f(ridge_epsilon) * np.maximum(max_ev, f(error_tolerance))  # DS:1159
  act_d = padding_start > np.arange(d)
  act_r = padding_start > np.arange(rank)
  sketch = sketch * act_d[:, None] * act_r  # DS:1167-1168
  fwd_eigs = (fwd_eigs + ridge) * act_r  # DS:1169-1170
  weighted = sketch * np.sqrt(fwd_eigs)  # DS:1171
  g = new_grad.astype(f) * act_d * act_d[:, None]  # DS:1172-1174
  updated = np.concatenate([np.sqrt(f(decay)) * weighted, g], axis=1)  # DS:1180-1192
  u, s, _ = np.linalg.svd(updated, full_matrices=False)  # DS:1193
  u, s = u.astype(f), s.astype(f)
  cutoff = s[rank]
  rho = cutoff**2
  top = s[:rank]
  deflated = (top - cutoff) * (top + cutoff)  # DS:1199
  vecs = u[:, :rank].copy()
  tail = f(tail) * f(decay)  # DS:1201
  new_tail = tail + rho
  alpha = f(-1.0 / p)
  with np.errstate(divide="ignore", invalid="ignore"):
    new_const = f(0) if new_tail <= 0 else new_tail**alpha  # DS:1205
  new_tail = f(0) if new_tail <= 0 else new_tail
  deflated = np.where(deflated <= 0, f(0), deflated)  # DS:1209
  vecs = vecs * (deflated > 0)  # DS:1210
  norms = np.linalg.norm(vecs, axis=0)  # DS:1214
  safe = (0.99 <= norms) & (norms <= 1.01)
  vecs = vecs * safe
  deflated = deflated * safe
  vecs = vecs / np.where(safe, norms, f(1.0))
  pad_ix = np.arange(d) >= padding_start  # DS:1224
  pad_mass = np.linalg.norm(vecs * pad_ix[:, None], axis=0, ord=1)
  has_pad = pad_mass > 0.01
  vecs = vecs * (1 - has_pad)
  deflated = deflated * (1 - has_pad)
  up = (np.square(top) + tail) * (deflated > 0.0)  # DS:1247-1248
  up = np.where(up <= 0, f(0), up)
  with np.errstate(divide="ignore"):
    inverted = np.where(up <= 0, f(0), np.power(np.where(up <= 0, f(1), up), alpha))
  has_zeros = bool(np.any(deflated <= 0) or new_tail <= 0)  # DS:1251
  val = fd_low_rank_pack(vecs.astype(f), deflated.astype(f), inverted.astype(f),
                         new_const, new_tail, has_zeros, rank)
  metrics = RootMetrics(inverse_pth_root_errors=0.0)  # DS:1263-1264
  if padding_start is not None and padding_start == 0:
    val = np.zeros_like(val)
  return val, metrics
