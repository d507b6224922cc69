"""Operator-level host API over the C ABI (torch tensors carry device memory).

Each function mirrors one seam of the reference (DS = distributed_shampoo.py):
``matrix_inverse_pth_root_batched`` <- ``_matrix_inverse_pth_root_vmap``
(DS:2742-2744), ``power_iteration`` (DS:595-652), ``quantize`` / ``dequantize``
(QU:49-113), ``grouped_gemm`` (DS:1468-1470, DS:1707), ``graft_momentum``
(DS:3496-3625).  All of them require a CUDA device; nothing falls back to CPU.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from precondition_b200 import _lib

_QDT = {torch.int16: _lib.PC_QDTYPE_INT16, torch.int8: _lib.PC_QDTYPE_INT8,
        torch.bfloat16: _lib.PC_QDTYPE_BF16, torch.float32: _lib.PC_QDTYPE_F32}

gpu_launches = 0  # count of C-ABI calls that launched kernels (bench bookkeeping)


def _stream() -> int:
  return torch.cuda.current_stream().cuda_stream


def _require_cuda(*tensors):
  for t in tensors:
    if t is None:
      continue
    if not t.is_cuda:
      raise RuntimeError("precondition_b200 needs CUDA tensors: there is no CPU fallback")
    if not t.is_contiguous():
      raise ValueError("tensor must be contiguous")


def _ptr(t: Optional[torch.Tensor]):
  return None if t is None else ctypes.c_void_p(t.data_ptr())


_workspaces = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
  key = (device.index if device.index is not None else torch.cuda.current_device())
  ws = _workspaces.get(key)
  if ws is None or ws.numel() < nbytes:
    ws = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=device)
    _workspaces[key] = ws
  return ws


def _as_i32_device(values, dev):
  """(device tensor, host numpy copy or None) of an exponent / padding vector."""
  if isinstance(values, torch.Tensor):
    host = None if values.is_cuda else values.to(torch.int32).numpy()
    return values.to(device=dev, dtype=torch.int32).contiguous(), host
  host = np.ascontiguousarray(np.asarray(values, dtype=np.int32))
  return torch.from_numpy(host).to(dev), host


def matrix_inverse_pth_root_batched(
    xs: torch.Tensor, ps, padding_starts=None, ridge_epsilon: float = 1e-6,
    error_tolerance: float = 1e-6, num_iters: int = 100,
    relative_matrix_epsilon: bool = True, engine: int = _lib.PC_ENGINE_AUTO,
    out: Optional[torch.Tensor] = None, metrics_out: Optional[torch.Tensor] = None,
    workspace: Optional[torch.Tensor] = None,
    ps_host=None) -> Tuple[torch.Tensor, torch.Tensor]:
  """Batched ``(A + eps I)^(-1/p)``; returns (roots [b,n,n], metrics [b,5]).

  The call only enqueues (CUDA graph with a device-driven Newton loop, see
  ``pc_inverse_pth_root_enqueue``).  ``ps_host`` (or ``ps`` given as a list / CPU tensor) tells
  the host the exponents without a device read-back.  ``workspace``: private scratch for calls
  that run concurrently on different streams (default: one shared buffer per device, which
  serialises through stream order only on ONE stream)."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(xs)
  if xs.dtype != torch.float32 or xs.dim() != 3 or xs.shape[1] != xs.shape[2]:
    raise TypeError("statistics must be a float32 [batch, n, n] tensor")
  b, n = xs.shape[0], xs.shape[1]
  dev = xs.device
  ps_t, host = _as_i32_device(ps, dev)
  if ps_host is None:
    ps_host = host
  pads_t = None
  if padding_starts is not None:
    pads_t, _ = _as_i32_device(padding_starts, dev)
  roots = out if out is not None else torch.empty_like(xs)
  metrics = metrics_out if metrics_out is not None else torch.empty(
      (b, _lib.PC_NUM_METRICS), dtype=torch.float32, device=dev)
  if b == 0:
    return roots, metrics
  hp = None
  if ps_host is not None:
    ps_host = np.ascontiguousarray(np.asarray(ps_host, dtype=np.int32))
    assert ps_host.shape == (b,)
    hp = ctypes.c_void_p(ps_host.ctypes.data)
  opt = _lib.RootOptions()
  lib.pc_root_options_default(ctypes.byref(opt))
  opt.ridge_epsilon, opt.error_tolerance = ridge_epsilon, error_tolerance
  opt.num_iters, opt.relative_matrix_epsilon = num_iters, int(relative_matrix_epsilon)
  opt.engine = engine
  nbytes = lib.pc_inverse_pth_root_workspace_bytes(b, n, engine)
  ws = workspace if workspace is not None else _workspace(nbytes, dev)
  if ws.numel() < nbytes:
    raise ValueError(f"workspace too small: {ws.numel()} < {nbytes}")
  with torch.cuda.device(dev):
    _lib.check(lib.pc_inverse_pth_root_enqueue(
        _ptr(xs), _ptr(ps_t), hp, _ptr(pads_t), b, n, ctypes.byref(opt), _ptr(roots),
        _ptr(metrics), _ptr(ws), ws.numel(), ctypes.c_void_p(_stream())))
  gpu_launches += 1
  return roots, metrics


def root_workspace_bytes(batch: int, n: int, engine: int = _lib.PC_ENGINE_AUTO) -> int:
  return int(_lib.load().pc_inverse_pth_root_workspace_bytes(batch, n, engine))


def select_scatter(src: torch.Tensor, src_off: torch.Tensor, metrics_base: torch.Tensor,
                   met_off: torch.Tensor, dst_idx: torch.Tensor, threshold: float,
                   dst: torch.Tensor, row_bytes: int, metrics_dst: Optional[torch.Tensor]):
  """``pc_select_scatter``: gathered rows -> state rows with the failure fallback of
  DS:2936-2950 (index arrays are device tensors: int64 byte / element offsets, int32 rows)."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(src, src_off, metrics_base, met_off, dst_idx, dst, metrics_dst)
  assert src_off.dtype == torch.int64 and met_off.dtype == torch.int64
  assert dst_idx.dtype == torch.int32 and metrics_base.dtype == torch.float32
  with torch.cuda.device(dst.device):
    _lib.check(lib.pc_select_scatter(
        _ptr(src), _ptr(src_off), _ptr(metrics_base), _ptr(met_off), _ptr(dst_idx),
        float(threshold), _ptr(dst), int(row_bytes), _ptr(metrics_dst), int(dst_idx.numel()),
        ctypes.c_void_p(_stream())))
  gpu_launches += 1


def fd_update_root_batched(
    new_grad: torch.Tensor, prev: torch.Tensor, ps, rank: int, padding_starts=None,
    ridge_epsilon: float = 1e-6, error_tolerance: float = 1e-6,
    relative_matrix_epsilon: bool = True, decay: float = 1.0, input_is_gram: bool = False,
    subspace_iters: Optional[int] = None, oversample: Optional[int] = None,
    full_eigh_max_dim: Optional[int] = None,
    out: Optional[torch.Tensor] = None,
    tearfree_epsilon: Optional[float] = None,
    tearfree_relative_epsilon: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
  """Batched Sketchy / frequent-directions step, ``_fd_update_root`` (DS:1123-1290).
  ``tearfree_epsilon`` (not None) selects tearfree's variant of the update (TF/sketchy.py:380-470,
  see ``pc_fd_options.tearfree``).

  new_grad [b, d, m]: a factor F with F F^T = x x^T (the reference's QR factor has m = d),
  or with ``input_is_gram`` the covariance x x^T [b, d, d].  prev / result: packed sketches
  [b, d, rank + 2].  Returns (packed, metrics [b, 5])."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(new_grad, prev)
  assert new_grad.dtype == torch.float32 and new_grad.dim() == 3
  assert prev.dtype == torch.float32 and prev.dim() == 3 and prev.shape[2] == rank + 2
  b, d, m = new_grad.shape
  assert prev.shape[0] == b and prev.shape[1] == d
  dev = new_grad.device
  ps_t = ps if isinstance(ps, torch.Tensor) and ps.is_cuda else \
      torch.as_tensor(ps, dtype=torch.int32).to(dev).contiguous()
  pads_t = None
  if padding_starts is not None:
    pads_t = padding_starts if isinstance(padding_starts, torch.Tensor) and \
        padding_starts.is_cuda else \
        torch.as_tensor(padding_starts, dtype=torch.int32).to(dev).contiguous()
  res = out if out is not None else torch.empty_like(prev)
  metrics = torch.empty((b, _lib.PC_NUM_METRICS), dtype=torch.float32, device=dev)
  if b == 0:
    return res, metrics
  opt = _lib.FdOptions()
  lib.pc_fd_options_default(ctypes.byref(opt))
  if tearfree_epsilon is not None:
    opt.tearfree, opt.tearfree_epsilon = 1, tearfree_epsilon
    opt.tearfree_relative_epsilon = int(tearfree_relative_epsilon)
  opt.ridge_epsilon, opt.error_tolerance = ridge_epsilon, error_tolerance
  opt.relative_matrix_epsilon, opt.decay = int(relative_matrix_epsilon), decay
  opt.input_is_gram = int(input_is_gram)
  if subspace_iters is not None:
    opt.subspace_iters = subspace_iters
  if oversample is not None:
    opt.oversample = oversample
  if full_eigh_max_dim is not None:
    opt.full_eigh_max_dim = full_eigh_max_dim
  nbytes = lib.pc_fd_update_workspace_bytes(b, d, m, rank, ctypes.byref(opt))
  ws = _workspace(nbytes, dev)
  with torch.cuda.device(dev):
    _lib.check(lib.pc_fd_update_batched(
        _ptr(new_grad), _ptr(prev), _ptr(ps_t), _ptr(pads_t), b, d, m, rank, ctypes.byref(opt),
        _ptr(res), _ptr(metrics), _ptr(ws), ws.numel(), ctypes.c_void_p(_stream())))
  gpu_launches += 1
  return res, metrics


def low_rank_root_batched(xs: torch.Tensor, ps, compression_rank: int, padding_starts=None,
                          ridge_epsilon: float = 1e-6, error_tolerance: float = 1e-6,
                          relative_matrix_epsilon: bool = True):
  """Batched eigh-based low-rank root, ``_low_rank_root`` (DS:1033-1120); d <= 512.
  Returns (packed [b, d, |rank| + 2], metrics [b, 5])."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(xs)
  b, d = xs.shape[0], xs.shape[1]
  dev = xs.device
  k = abs(compression_rank)
  ps_t = torch.as_tensor(ps, dtype=torch.int32).to(dev).contiguous()
  pads_t = None
  if padding_starts is not None:
    pads_t = torch.as_tensor(padding_starts, dtype=torch.int32).to(dev).contiguous()
  out = torch.empty((b, d, k + 2), dtype=torch.float32, device=dev)
  metrics = torch.empty((b, _lib.PC_NUM_METRICS), dtype=torch.float32, device=dev)
  if b == 0:
    return out, metrics
  ws = _workspace(lib.pc_low_rank_root_workspace_bytes(b, d), dev)
  with torch.cuda.device(dev):
    _lib.check(lib.pc_low_rank_root_batched(
        _ptr(xs), _ptr(ps_t), _ptr(pads_t), b, d, compression_rank, ridge_epsilon,
        error_tolerance, int(relative_matrix_epsilon), _ptr(out), _ptr(metrics), _ptr(ws),
        ws.numel(), ctypes.c_void_p(_stream())))
  gpu_launches += 1
  return out, metrics


def matrix_inverse_pth_root_eigh_batched(xs: torch.Tensor, ps, padding_starts=None,
                                         ridge_epsilon: float = 1e-6,
                                         error_tolerance: float = 1e-6,
                                         relative_matrix_epsilon: bool = True,
                                         out: Optional[torch.Tensor] = None, **_unused):
  """Batched ``matrix_inverse_pth_root_eigh`` (DS:943-1030, the ``eigh=True`` root); d <= 512."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(xs)
  b, d = xs.shape[0], xs.shape[1]
  dev = xs.device
  ps_t = torch.as_tensor(ps, dtype=torch.int32).to(dev).contiguous()
  pads_t = None
  if padding_starts is not None:
    pads_t = torch.as_tensor(padding_starts, dtype=torch.int32).to(dev).contiguous()
  roots = out if out is not None else torch.empty_like(xs)
  metrics = torch.empty((b, _lib.PC_NUM_METRICS), dtype=torch.float32, device=dev)
  if b == 0:
    return roots, metrics
  ws = _workspace(lib.pc_low_rank_root_workspace_bytes(b, d), dev)
  with torch.cuda.device(dev):
    _lib.check(lib.pc_inverse_pth_root_eigh_batched(
        _ptr(xs), _ptr(ps_t), _ptr(pads_t), b, d, ridge_epsilon, error_tolerance,
        int(relative_matrix_epsilon), _ptr(roots), _ptr(metrics), _ptr(ws), ws.numel(),
        ctypes.c_void_p(_stream())))
  gpu_launches += 1
  return roots, metrics


def pinv_pth_root_eigh_batched(xs: torch.Tensor, ps, rel_cutoff: float = 1e-6,
                               out: Optional[torch.Tensor] = None,
                               eigvecs: Optional[torch.Tensor] = None,
                               eigvecs_valid: bool = False) -> torch.Tensor:
  """Batched pseudo-inverse p-th root by eigendecomposition, tearfree's ``_pth_inv_root``
  (TF/shampoo.py:440-448): eigenvalues <= rel_cutoff * max are dropped.  xs [b, d, d]; d <= 2048."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(xs, out)
  b, d = xs.shape[0], xs.shape[1]
  dev = xs.device
  ps_t = ps if isinstance(ps, torch.Tensor) else torch.as_tensor(ps, dtype=torch.int32).to(dev)
  roots = out if out is not None else torch.empty_like(xs)
  if b == 0:
    return roots
  ws = _workspace(lib.pc_low_rank_root_workspace_bytes(b, d), dev)
  with torch.cuda.device(dev):
    # eigvecs [b, d, d] (optional, in / out): warm start of the eigen-solve, see the header
    _require_cuda(eigvecs)
    _lib.check(lib.pc_pinv_pth_root_eigh_warm_batched(
        _ptr(xs), _ptr(ps_t), b, d, rel_cutoff, _ptr(roots), _ptr(eigvecs),
        int(bool(eigvecs_valid and eigvecs is not None)), _ptr(ws), ws.numel(),
        ctypes.c_void_p(_stream())))
  gpu_launches += 1
  return roots


class TearfreeTail:
  """Launch list of ``pc_tearfree_transform`` for a fixed set of parameters: grafting, momentum,
  weight decay and the learning rate for all of them in three launches.  The state tensors
  (``acc``, ``velocity``) are fixed at construction; gradients, directions and outputs are
  bound per call (their addresses change from step to step)."""

  def __init__(self, numels: Sequence[int], device):
    lib = _lib.load()
    self.device = device
    self.n = len(numels)
    chunk = int(lib.pc_graft_group_chunk_elems())
    self.segs = (_lib.TearfreeSegment * max(self.n, 1))()
    chunk_seg, first = [], 0
    for i, ne in enumerate(numels):
      nch = max(1, -(-int(ne) // chunk))
      self.segs[i].numel, self.segs[i].first_chunk, self.segs[i].nchunks = int(ne), first, nch
      chunk_seg += [i] * nch
      first += nch
    self.total_chunks = first
    self.chunk_seg = torch.tensor(chunk_seg, dtype=torch.int32).to(device)
    nbytes = lib.pc_tearfree_transform_workspace_bytes(self.n, self.total_chunks)
    self.ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
    self.segs_dev = torch.empty(ctypes.sizeof(self.segs), dtype=torch.uint8, device=device)
    # two pinned staging copies: the upload of step k may still be queued when step k+1 is built
    self.segs_host = [torch.empty(ctypes.sizeof(self.segs), dtype=torch.uint8).pin_memory()
                      for _ in range(2)]
    self.uploaded = [None, None]
    self.turn = 0

  def run(self, grads, params, preconds, accs, velocities, updates, opt: "_lib.TearfreeOptions"):
    """Each argument: list of tensors (fp32, contiguous) or None entries where the stage is off."""
    global gpu_launches
    if not self.n:
      return
    lib = _lib.load()
    for i in range(self.n):
      sg = self.segs[i]
      for name, lst in (("grad", grads), ("param", params), ("precond", preconds),
                        ("acc", accs), ("velocity", velocities), ("update", updates)):
        t = lst[i] if lst is not None else None
        if t is not None:
          _require_cuda(t)
          assert t.numel() == sg.numel, (name, i, t.shape, sg.numel)
          if t.data_ptr() % 16:
            raise ValueError(f"tearfree: {name}[{i}] is not 16-byte aligned")
        setattr(sg, name, None if t is None else t.data_ptr())
    t = self.turn
    self.turn ^= 1
    if self.uploaded[t] is not None:
      self.uploaded[t].synchronize()
    ctypes.memmove(self.segs_host[t].data_ptr(), ctypes.addressof(self.segs),
                   ctypes.sizeof(self.segs))
    self.segs_dev.copy_(self.segs_host[t], non_blocking=True)
    self.uploaded[t] = torch.cuda.Event()
    self.uploaded[t].record()
    with torch.cuda.device(self.device):
      _lib.check(lib.pc_tearfree_transform(_ptr(self.segs_dev), _ptr(self.chunk_seg), self.n,
                                           self.total_chunks, ctypes.byref(opt), _ptr(self.ws),
                                           self.ws.numel(), ctypes.c_void_p(_stream())))
    gpu_launches += 3 if opt.graft_type != _lib.PC_TF_GRAFT_NONE else 1


def low_rank_factors(packed: torch.Tensor, rank: int, w: torch.Tensor, c: torch.Tensor):
  """Factors of the operator of packed low-rank preconditioners [b, d, rank + 2] for the
  application g -> c g + (g V) W^T (DS:1690-1705): w [b, d, rank] = V diag(lambda^- - c),
  c [b] (identity -- w = 0, c = 1 -- when has_zeros)."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(packed, w, c)
  b, d = packed.shape[0], packed.shape[1]
  assert packed.shape[2] == rank + 2 and tuple(w.shape) == (b, d, rank) and c.numel() == b
  with torch.cuda.device(packed.device):
    _lib.check(lib.pc_low_rank_factors(_ptr(packed), b, d, rank, _ptr(w), _ptr(c),
                                       ctypes.c_void_p(_stream())))
  gpu_launches += 1


def low_rank_to_dense(packed: torch.Tensor, rank: int,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
  """Dense operator of packed low-rank preconditioners [b, d, rank+2] -> [b, d, d]
  (``c I + V diag(lambda^- - c) V^T``; identity when has_zeros), DS:1690-1705."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(packed)
  b, d, pd = packed.shape
  assert pd == rank + 2 and packed.dtype == torch.float32
  dense = out if out is not None else torch.empty((b, d, d), dtype=torch.float32,
                                                  device=packed.device)
  if b == 0:
    return dense
  ws = _workspace(lib.pc_low_rank_to_dense_workspace_bytes(b, d, rank), packed.device)
  with torch.cuda.device(packed.device):
    _lib.check(lib.pc_low_rank_to_dense(_ptr(packed), b, d, rank, _ptr(dense), _ptr(ws),
                                        ws.numel(), ctypes.c_void_p(_stream())))
  gpu_launches += 1
  return dense


def debug_tc_gemm(a: torch.Tensor, b: torch.Tensor, passes: int = 6) -> torch.Tensor:
  """Test hook: C = A @ B^T on the tcgen05 split-bf16 engine ([batch, n, n] fp32)."""
  lib = _lib.load()
  _require_cuda(a, b)
  bt, n = a.shape[0], a.shape[1]
  c = torch.empty_like(a)
  nbytes = lib.pc_inverse_pth_root_workspace_bytes(bt, n, _lib.PC_ENGINE_TC_BF16X6) + (1 << 16)
  ws = _workspace(nbytes, a.device)
  with torch.cuda.device(a.device):
    _lib.check(lib.pc_debug_tc_gemm(_ptr(a), _ptr(b), _ptr(c), bt, n, passes, _ptr(ws),
                                    ws.numel(), ctypes.c_void_p(_stream())))
  return c


def power_iteration(xs: torch.Tensor, padding_starts=None, num_iters: int = 100,
                    error_tolerance: float = 1e-6):
  """Batched DS:595-652; returns (lambda [b], iters [b])."""
  lib = _lib.load()
  _require_cuda(xs)
  b, n = xs.shape[0], xs.shape[1]
  pads_t = None
  if padding_starts is not None:
    pads_t = torch.as_tensor(padding_starts, dtype=torch.int32).to(xs.device).contiguous()
  lam = torch.empty(b, dtype=torch.float32, device=xs.device)
  its = torch.empty(b, dtype=torch.int32, device=xs.device)
  with torch.cuda.device(xs.device):
    _lib.check(lib.pc_power_iteration_batched(_ptr(xs), _ptr(pads_t), b, n, num_iters,
                                              error_tolerance, _ptr(lam), _ptr(its),
                                              ctypes.c_void_p(_stream())))
  return lam, its


EIGH_MAX_DIM = 2048  # largest statistic of the Cholesky + Jacobi eigh pipeline (one cluster)


def quantize(x: torch.Tensor, qdtype: torch.dtype, extract_diagonal: bool = False, out=None):
  """QU:49-95 on a [batch, rows, cols] (or [rows, cols]) tensor.

  Returns (quantized, diagonal or None, bucket_size or None); ``out`` = (q, diag, bucket)
  tensors to write into (int16 / int8 only)."""
  lib = _lib.load()
  _require_cuda(x)
  if x.dtype != torch.float32:
    raise TypeError(f"quantize expects float32, got {x.dtype}")
  squeeze = x.dim() == 2
  xb = x.unsqueeze(0) if squeeze else x
  b, rows, cols = xb.shape
  if qdtype == torch.float32:
    return x, None, None
  diag = bucket = None
  if out is not None:
    q, diag, bucket = out
    _require_cuda(q, diag, bucket)
    assert q.dtype == qdtype and q.numel() == xb.numel()
  else:
    q = torch.empty(xb.shape, dtype=qdtype, device=x.device)
    if qdtype != torch.bfloat16:
      bucket = torch.empty((b, cols), dtype=torch.float32, device=x.device)
      if extract_diagonal:
        diag = torch.empty((b, rows), dtype=torch.float32, device=x.device)
  with torch.cuda.device(x.device):
    _lib.check(lib.pc_quantize_batched(_ptr(xb), b, rows, cols, _QDT[qdtype],
                                       int(extract_diagonal), _ptr(q), _ptr(diag),
                                       _ptr(bucket), ctypes.c_void_p(_stream())))
  if squeeze and out is None:
    q = q[0]
    diag = None if diag is None else diag[0]
    bucket = None if bucket is None else bucket[0]
  return q, diag, bucket


def dequantize(q: torch.Tensor, diag, bucket, extract_diagonal: bool = False,
               out: Optional[torch.Tensor] = None):
  """QU:97-113."""
  lib = _lib.load()
  _require_cuda(q)
  if q.dtype == torch.float32:
    return q
  squeeze = q.dim() == 2
  qb = q.unsqueeze(0) if squeeze else q
  b, rows, cols = qb.shape
  x = out if out is not None else torch.empty(qb.shape, dtype=torch.float32, device=q.device)
  with torch.cuda.device(q.device):
    _lib.check(lib.pc_dequantize_batched(_ptr(qb), _ptr(diag), _ptr(bucket), b, rows, cols,
                                         _QDT[q.dtype], int(extract_diagonal), _ptr(x),
                                         ctypes.c_void_p(_stream())))
  return x[0] if (squeeze and out is None) else x


class QuantGroup:
  """Static launch list of ``pc_dequantize_grouped`` / ``pc_quantize_grouped``: the int8 momenta
  of a whole model in one / two launches.  ``items``: (q int8 tensor, bucket fp32 tensor, x fp32
  view) with q.numel() == x.numel(); the per-column layout is [q.shape[0], rest] (QU:86)."""

  def __init__(self, items, device):
    lib = _lib.load()
    self.device = device
    self.n = len(items)
    self.keep = items  # the tensors the table points at
    chunk = int(lib.pc_quant_group_chunk_elems())
    tile_rows = int(lib.pc_quant_group_tile_rows())
    # every segment's column maxima start 16-byte aligned in the shared scratch
    total_cols = sum(-(-int(b.numel()) // 4) * 4 for _, b, _ in items)
    self.colmax = torch.zeros(max(total_cols, 4), dtype=torch.int32, device=device)
    segs = (_lib.QuantSegment * max(self.n, 1))()
    chunk_seg, first, coff = [], 0, 0
    tile_seg, first_tile = [], 0
    for i, (q, bucket, x) in enumerate(items):
      _require_cuda(q, bucket, x)
      assert q.dtype == torch.int8 and bucket.dtype == torch.float32 and x.dtype == torch.float32
      rows = int(q.shape[0]) if q.dim() > 0 else 1
      cols = int(q.numel()) // max(rows, 1)
      assert bucket.numel() == cols and x.numel() == q.numel(), (q.shape, bucket.shape, x.shape)
      nch = max(1, -(-int(q.numel()) // chunk))
      sg = segs[i]
      sg.q, sg.bucket, sg.x = q.data_ptr(), bucket.data_ptr(), x.data_ptr()
      sg.colmax = self.colmax.data_ptr() + 4 * coff
      sg.rows, sg.cols, sg.first_chunk, sg.nchunks = rows, cols, first, nch
      sg.first_tile, sg.col_tiles = first_tile, -(-cols // 128)
      ntiles = -(-rows // tile_rows) * sg.col_tiles
      chunk_seg += [i] * nch
      tile_seg += [i] * ntiles
      first += nch
      first_tile += ntiles
      coff += -(-cols // 4) * 4
    self.total_chunks = first
    self.total_tiles = first_tile
    self.tile_seg = torch.tensor(tile_seg, dtype=torch.int32).to(device)
    self.segs = torch.frombuffer(bytearray(bytes(segs)), dtype=torch.uint8).to(device)
    self.chunk_seg = torch.tensor(chunk_seg, dtype=torch.int32).to(device)

  def dequantize(self):
    global gpu_launches
    if not self.n:
      return
    with torch.cuda.device(self.device):
      _lib.check(_lib.load().pc_dequantize_grouped(_ptr(self.segs), _ptr(self.chunk_seg), self.n,
                                                   self.total_chunks, ctypes.c_void_p(_stream())))
    gpu_launches += 1

  def quantize(self):
    global gpu_launches
    if not self.n:
      return
    with torch.cuda.device(self.device):
      _lib.check(_lib.load().pc_quantize_grouped(
          _ptr(self.segs), _ptr(self.chunk_seg), self.n, self.total_chunks, _ptr(self.tile_seg),
          self.total_tiles, _ptr(self.colmax), self.colmax.numel() * 4,
          ctypes.c_void_p(_stream())))
    gpu_launches += 2


def upload_gemm_descs(descs: Sequence[_lib.GemmDesc], device) -> torch.Tensor:
  """Packs descriptors into a device byte tensor (kept alive by the caller)."""
  arr = (_lib.GemmDesc * len(descs))(*descs)
  host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
  return host.to(device)


def grouped_gemm(descs_dev: torch.Tensor, count: int, max_m: int, max_n: int):
  global gpu_launches
  lib = _lib.load()
  _require_cuda(descs_dev)
  with torch.cuda.device(descs_dev.device):
    _lib.check(lib.pc_grouped_gemm(_ptr(descs_dev), count, max_m, max_n,
                                   ctypes.c_void_p(_stream())))
  gpu_launches += 1


class SimtGemmLists:
  """CUDA-core grouped-GEMM launch lists for one phase.  Descriptors are grouped into size
  classes (tile counts rounded up to powers of two) so that no launch's grid is sized by
  another class's largest block, and descriptors with a small output but a very long
  contraction go through the deterministic split-K kernel."""

  SPLITK_MIN_K = 16384
  SPLITK_MAX_TILES = 4

  def __init__(self, descs: Sequence[_lib.GemmDesc], device):
    self.device = device
    self.groups = []   # (device descriptors, count, max_m, max_n)
    self.splitk = []   # (device descriptors, count, max_m, max_n, splits, workspace)
    self.thin = []     # (device descriptors, count, max_m, max_n, kind): pc_grouped_gemm_thin
    self.group_descs, self.splitk_descs = [], []  # host descriptors of each launch (diagnostics)
    self.thin_descs = []
    classes, long_k, thin = {}, [], {}

    def tiles(x):
      t = (x + 63) // 64
      return 1 << max(t - 1, 0).bit_length()

    for d in descs:
      if d.k >= self.SPLITK_MIN_K and tiles(d.m) * tiles(d.n) <= self.SPLITK_MAX_TILES:
        long_k.append(d)
      elif thin_outer_eligible(d):  # statistic of a rank-1 parameter: streaming outer product
        thin.setdefault(_lib.PC_THIN_OUTER, []).append(d)
      elif d.m <= 4:  # a vector (rank-1 parameter) times a matrix: streaming kernel
        thin.setdefault(_lib.PC_THIN_GEMV, []).append(d)
      elif d.n <= 16 and d.k <= 16 and d.m >= 1024:  # mode product with a tiny preconditioner
        thin.setdefault(_lib.PC_THIN_ROWMAP, []).append(d)
      else:
        classes.setdefault((tiles(d.m), tiles(d.n)), []).append(d)
    for lst in classes.values():
      self.groups.append((upload_gemm_descs(lst, device), len(lst),
                          max(d.m for d in lst), max(d.n for d in lst)))
      self.group_descs.append(lst)
    for kind, lst in thin.items():
      self.thin.append((upload_gemm_descs(lst, device), len(lst), max(d.m for d in lst),
                        max(d.n for d in lst), kind))
      self.thin_descs.append(lst)
    # outputs of at most 16 x 16 (the 9 x 9 statistic of a 3 x 3 kernel) take the thin split-K
    # kernel, which wants many short splits; the others go through the 64 x 64 tile kernel
    for lk in ([d for d in long_k if d.m <= 16 and d.n <= 16],
               [d for d in long_k if not (d.m <= 16 and d.n <= 16)]):
      if lk:
        self._add_splitk(lk, device)

  def _add_splitk(self, long_k, device):
    if long_k:
      lib = _lib.load()
      mm, mn = max(d.m for d in long_k), max(d.n for d in long_k)
      if mm <= 16 and mn <= 16:
        splits = max(1, min(256, min(d.k for d in long_k) // 1024, 65535 // len(long_k)))
      else:
        splits = max(1, min(64, min(d.k for d in long_k) // 2048, 65535 // len(long_k)))
      nbytes = lib.pc_grouped_gemm_splitk_workspace_bytes(len(long_k), mm, mn, splits)
      ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
      self.splitk.append((upload_gemm_descs(long_k, device), len(long_k), mm, mn, splits, ws))
      self.splitk_descs.append(long_k)

  def __bool__(self):
    return bool(self.groups or self.splitk or self.thin)

  def run(self):
    global gpu_launches
    lib = _lib.load()
    for dev, count, mm, mn in self.groups:
      grouped_gemm(dev, count, mm, mn)
    for dev, count, mm, mn, kind in self.thin:
      with torch.cuda.device(self.device):
        _lib.check(lib.pc_grouped_gemm_thin(_ptr(dev), count, mm, mn, kind,
                                            ctypes.c_void_p(_stream())))
      gpu_launches += 1
    for dev, count, mm, mn, splits, ws in self.splitk:
      with torch.cuda.device(self.device):
        _lib.check(lib.pc_grouped_gemm_splitk(_ptr(dev), count, mm, mn, splits, _ptr(ws),
                                              ws.numel(), ctypes.c_void_p(_stream())))
      gpu_launches += 1


# smallest ragged product (multiply-adds) that is sent to the tensor cores: below, zero-padding
# the edge tiles to 128 costs more than the CUDA-core tile
TC_RAGGED_MIN_MACS = 1 << 21


def tc_gemm_eligible(d: _lib.GemmDesc) -> bool:
  """Descriptor can run on the tcgen05 grouped GEMM (include/precond_b200.h).  Output sizes that
  are multiples of 128 always do; ragged blocks (1000 x 1000 statistics, 576-row kernels: edge
  tiles zero-filled by the pack, masked by the epilogue) when they are large enough to pay."""
  if not (d.m > 0 and d.n > 0 and d.k > 0 and (d.c or 0) % 16 == 0 and d.c_sii % 4 == 0 and
          d.c_sio % 4 == 0 and (d.c_in or 0) % 16 == 0):
    return False
  if d.m % 128 == 0 and d.n % 128 == 0:
    return True
  return (d.n % 4 == 0 and d.m >= 64 and d.n >= 64 and
          d.m * d.n * d.k >= TC_RAGGED_MIN_MACS)


def thin_outer_eligible(d: _lib.GemmDesc) -> bool:
  """Contraction of at most 4 into a large output (the statistic of a rank-1 parameter): a
  streaming pass over the output (``PC_THIN_OUTER``) beats any tile kernel."""
  return d.k <= 4 and d.m >= 64 and d.n >= 64


def tc_gemm_fused_quant_eligible(d: _lib.GemmDesc) -> bool:
  """... and its (de)quantisation can be fused into the epilogue (pc_grouped_gemm_tc_quant):
  whole 128 x 128 tiles only (16-byte reads of the quantised rows)."""
  return tc_gemm_eligible(d) and d.m % 128 == 0 and d.n % 128 == 0


class TcGemmList:
  """Host-side descriptor array for ``pc_grouped_gemm_tc`` (kept alive by the caller)."""

  def __init__(self, descs: Sequence[_lib.GemmDesc], device,
               quant: Optional[Sequence[_lib.GemmQuant]] = None):
    self.count = len(descs)
    self.arr = (_lib.GemmDesc * max(self.count, 1))(*descs)
    # optional fused (de)quantisation extensions, one per descriptor (pc_gemm_quant)
    self.quant = (_lib.GemmQuant * max(self.count, 1))(*quant) if quant else None
    self.device = device
    lib = _lib.load()
    self.nbytes = lib.pc_grouped_gemm_tc_workspace_bytes(
        ctypes.cast(self.arr, ctypes.c_void_p), self.count) if self.count else 0
    self.ws = None

  def run(self):
    global gpu_launches
    if not self.count:
      return
    lib = _lib.load()
    reuse = self.ws is not None  # private workspace: the uploaded plan stays valid
    if self.ws is None:
      self.ws = torch.empty(self.nbytes + 4096, dtype=torch.uint8, device=self.device)
    with torch.cuda.device(self.device):
      if self.quant is not None:
        _lib.check(lib.pc_grouped_gemm_tc_quant(
            ctypes.cast(self.arr, ctypes.c_void_p), ctypes.cast(self.quant, ctypes.c_void_p),
            self.count, _ptr(self.ws), self.ws.numel(), int(reuse), ctypes.c_void_p(_stream())))
      else:
        _lib.check(lib.pc_grouped_gemm_tc(ctypes.cast(self.arr, ctypes.c_void_p), self.count,
                                          _ptr(self.ws), self.ws.numel(), int(reuse),
                                          ctypes.c_void_p(_stream())))
    gpu_launches += 1


def quantize_from_colmax(x: torch.Tensor, colmax: torch.Tensor, qdtype, q: torch.Tensor,
                         diag: torch.Tensor, bucket: torch.Tensor):
  """``from_float`` with extracted diagonal (QU:49-95) when the per-column maxima of
  |off-diagonal| are already known (``colmax`` [b, n] int32 bit patterns): one pass over x."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(x, colmax, q, diag, bucket)
  b, n = x.shape[0], x.shape[1]
  with torch.cuda.device(x.device):
    _lib.check(lib.pc_quantize_from_colmax_batched(
        _ptr(x), _ptr(colmax), b, n, _QDT[qdtype], _ptr(q), _ptr(diag), _ptr(bucket),
        ctypes.c_void_p(_stream())))
  gpu_launches += 1


def make_graft_options(**kw) -> _lib.GraftOptions:
  o = _lib.GraftOptions()
  for k, v in kw.items():
    setattr(o, k, v)
  return o


def graft_momentum(grad, param, precond_grad, diagonal_statistics, diagonal_momentum,
                   momentum, update, opt: _lib.GraftOptions):
  """DS:3496-3625 tail for one parameter; state tensors are updated in place."""
  global gpu_launches
  lib = _lib.load()
  _require_cuda(grad, param, precond_grad, diagonal_statistics, diagonal_momentum, momentum,
                update)
  for t in (grad, param, precond_grad, diagonal_statistics, diagonal_momentum, momentum, update):
    if t is not None and t.dtype != torch.float32:
      raise TypeError(f"pc_graft_momentum works on float32 tensors, got {t.dtype}")
  numel = grad.numel()
  nbytes = lib.pc_graft_momentum_workspace_bytes(numel)
  ws = _workspace(nbytes, grad.device)
  with torch.cuda.device(grad.device):
    _lib.check(lib.pc_graft_momentum(
        _ptr(grad), _ptr(param), _ptr(precond_grad), _ptr(diagonal_statistics),
        _ptr(diagonal_momentum), _ptr(momentum), _ptr(update), numel, ctypes.byref(opt),
        _ptr(ws), ws.numel(), ctypes.c_void_p(_stream())))
  gpu_launches += 1


class GraftGroup:
  """Static segment table for ``pc_graft_momentum_grouped``: the tail of ``_transform_grad``
  (DS:3496-3625) for every parameter in a fixed number of launches.  ``segments`` is a list
  of (offset, numel, has_precond) into the optimizer's flat buffers."""

  def __init__(self, segments, device):
    lib = _lib.load()
    chunk = int(lib.pc_graft_group_chunk_elems())
    segs, chunk_seg, first = [], [], 0
    for i, (off, numel, has_precond) in enumerate(segments):
      if numel <= 0:
        continue
      assert off % 4 == 0, "segments must be 16-byte aligned"
      n = (numel + chunk - 1) // chunk
      sg = _lib.GraftSegment()
      sg.offset, sg.numel, sg.first_chunk, sg.nchunks = off, numel, first, n
      sg.has_precond = int(bool(has_precond))
      segs.append(sg)
      chunk_seg.extend([len(segs) - 1] * n)
      first += n
    self.count, self.total_chunks = len(segs), first
    self.device = device
    if self.count:
      arr = (_lib.GraftSegment * self.count)(*segs)
      self.segs = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(device)
      self.chunk_seg = torch.tensor(chunk_seg, dtype=torch.int32, device=device)
      nbytes = lib.pc_graft_momentum_grouped_workspace_bytes(self.count, self.total_chunks)
      self.ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)

  def run(self, grad, param, precond_grad, diagonal_statistics, diagonal_momentum, momentum,
          update, opt: _lib.GraftOptions):
    """All arguments are the flat fp32 buffers (param / precond_grad / diagonal_statistics
    may be None); momenta, diagonal statistics and update are written in place."""
    global gpu_launches
    if not self.count:
      return
    lib = _lib.load()
    _require_cuda(grad, param, precond_grad, diagonal_statistics, diagonal_momentum, momentum,
                  update)
    for t in (grad, param, precond_grad, diagonal_statistics, diagonal_momentum, momentum, update):
      if t is not None and t.dtype != torch.float32:
        raise TypeError("pc_graft_momentum_grouped works on float32 buffers")
    with torch.cuda.device(self.device):
      _lib.check(lib.pc_graft_momentum_grouped(
          _ptr(self.segs), _ptr(self.chunk_seg), self.count, self.total_chunks, _ptr(grad),
          _ptr(param), _ptr(precond_grad), _ptr(diagonal_statistics), _ptr(diagonal_momentum),
          _ptr(momentum), _ptr(update), ctypes.byref(opt), _ptr(self.ws), self.ws.numel(),
          ctypes.c_void_p(_stream())))
    gpu_launches += 1


def batched_matmul(a: torch.Tensor, b: torch.Tensor, transpose_b: bool = False,
                   alpha: float = 1.0, c_in: Optional[torch.Tensor] = None, beta: float = 0.0,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
  """C[z] = alpha * A[z] @ (B[z] or B[z]^T) + beta * C_in[z] through the grouped GEMM (tcgen05
  when the output sizes are multiples of 128, fp32 CUDA cores otherwise).  a [bt, m, k];
  b [bt, k, n] (or [bt, n, k] with transpose_b)."""
  _require_cuda(a, b, c_in, out)
  bt, m, k = a.shape
  n = b.shape[1] if transpose_b else b.shape[2]
  assert (b.shape[2] if transpose_b else b.shape[1]) == k and b.shape[0] == bt
  c = out if out is not None else torch.empty((bt, m, n), dtype=torch.float32, device=a.device)
  descs = []
  for z in range(bt):
    d = _lib.GemmDesc()
    d.a = a[z].data_ptr()
    d.b = b[z].data_ptr()
    d.c = c[z].data_ptr()
    d.c_in = None if c_in is None else c_in[z].data_ptr()
    d.a_iinner, d.a_sio, d.a_si = m, 0, k
    d.a_kinner, d.a_sko, d.a_ski = k, 0, 1
    if transpose_b:   # B(j, kk) = b[j, kk]
      d.b_sj, d.b_kinner, d.b_sko, d.b_ski = k, k, 0, 1
    else:             # B(j, kk) = b[kk, j]
      d.b_sj, d.b_kinner, d.b_sko, d.b_ski = 1, k, 0, n
    d.c_iinner, d.c_sio, d.c_sii = m, 0, n
    d.m, d.n, d.k, d.alpha, d.beta = m, n, k, alpha, beta
    descs.append(d)
  use_tc = bool(_lib.load().pc_device_supports_tcgen05()) and all(tc_gemm_eligible(d) for d in descs)
  (TcGemmList(descs, a.device) if use_tc else SimtGemmLists(descs, a.device)).run()
  return c


def _lobpcg_call(fn, *args):
  global gpu_launches
  _lib.check(fn(*args, ctypes.c_void_p(_stream())))
  gpu_launches += 1


def root_diagnostics(root: torch.Tensor, matrix: torch.Tensor, ps) -> torch.Tensor:
  """InversePthRootDiagnostics (DS:109-142) for a batch: rows {max_diag_error, avg_diag_error,
  max_off_diag_error, avg_off_diag_error, p} of M = root^p @ matrix (mat_power order, DS:655-678)."""
  lib = _lib.load()
  bt, n = root.shape[0], root.shape[1]
  ps_list = [int(x) for x in (ps.tolist() if isinstance(ps, torch.Tensor) else ps)]
  out = torch.zeros((bt, 5), dtype=torch.float32, device=root.device)
  mat_m = torch.empty_like(root)
  for p in sorted(set(ps_list)):
    idx = [z for z, q in enumerate(ps_list) if q == p]
    sel = torch.tensor(idx, device=root.device)
    r, mtx = root[sel].contiguous(), matrix[sel].contiguous()
    power, mat, i = None, r, p
    while i > 0:
      if i % 2 == 1:
        power = mat if power is None else batched_matmul(mat, power)
      i //= 2
      if i > 0:
        mat = batched_matmul(mat, mat)
    mat_m[sel] = batched_matmul(power, mtx)
    out[sel, 4] = float(p)
  out4 = torch.empty((bt, 4), dtype=torch.float32, device=root.device)
  with torch.cuda.device(root.device):
    _lobpcg_call(lib.pc_root_diagnostics, _ptr(mat_m), bt, n, _ptr(out4))
  out[:, :4] = out4
  return out


def matrix_inverse_pth_root_lobpcg_batched(
    xs: torch.Tensor, ps, topk: int, padding_starts=None, ridge_epsilon: float = 1e-6,
    error_tolerance: float = 1e-6, num_iters: int = 100, relative_matrix_epsilon: bool = True,
    engine: int = _lib.PC_ENGINE_AUTO, lobpcg_max_iter: int = 0, diagnostics: bool = True):
  """matrix_inverse_pth_root with ``lobpcg_topk_precondition = topk`` (DS:789-812, DS:889-928):
  the top-k eigenpairs are deflated out of every matrix before the coupled Newton iteration
  (lower condition number -> fewer iterations) and put back into the root afterwards.

  Returns (roots [b,n,n], metrics [b,5], diag) with diag = {"lobpcg": [b,7] LOBPCGDiagnostics
  rows, "inverse_pth_root": [b,5], "conditioned_inverse_pth_root": [b,5]} (None without
  ``diagnostics``).  As in the reference, ``metrics[:, 0]`` is then the entrywise error of the
  UNCONDITIONED problem (DS:913-921) and ``metrics[:, 3]`` the largest deflated eigenvalue."""
  lib = _lib.load()
  _require_cuda(xs)
  b, n = xs.shape[0], xs.shape[1]
  k = int(topk)
  if not 1 <= k < n - 2:
    raise ValueError(f"lobpcg_topk_precondition = {k} needs 1 <= k < n - 2 (n = {n})")
  dev = xs.device
  ps_t, ps_host = _as_i32_device(ps, dev)
  iters = lobpcg_max_iter if lobpcg_max_iter > 0 else max(8, min(k, 16))
  prev = torch.zeros((b, n, k + 2), dtype=torch.float32, device=dev)
  # top-k eigenpairs: block subspace iteration (exact Jacobi eigh for n <= 512) on the matrix
  packed, _ = fd_update_root_batched(xs, prev, ps_t, k, padding_starts, ridge_epsilon=0.0,
                                     relative_matrix_epsilon=False, decay=1.0,
                                     input_is_gram=True, subspace_iters=iters)
  scal = torch.empty((b, 4), dtype=torch.float32, device=dev)
  eig = torch.empty((b, k), dtype=torch.float32, device=dev)
  s1 = torch.empty((b, n, k), dtype=torch.float32, device=dev)
  a_scaled = torch.empty_like(xs)
  with torch.cuda.device(dev):
    _lobpcg_call(lib.pc_lobpcg_deflate_prep, _ptr(packed), _ptr(xs), b, n, k, float(ridge_epsilon),
                 int(relative_matrix_epsilon), _ptr(scal), _ptr(eig), _ptr(s1), _ptr(a_scaled))
  # deflate: A' = A / m - S1 S1^T  (DS:805-812)
  deflated = batched_matmul(s1, s1, transpose_b=True, alpha=-1.0, c_in=a_scaled, beta=1.0)
  eps_abs = ridge_epsilon if relative_matrix_epsilon else None
  roots, metrics = matrix_inverse_pth_root_batched(
      deflated, ps_t, padding_starts, ridge_epsilon=ridge_epsilon, error_tolerance=error_tolerance,
      num_iters=num_iters, relative_matrix_epsilon=False, engine=engine, ps_host=ps_host)
  del eps_abs
  diag = None
  if diagnostics:
    # conditioned problem (scaled units: root^p (A' + eps 10^t I) is scale invariant), DS:903-906
    tries = metrics[:, 4].clamp_min(1.0) - 1.0
    damped = deflated.clone()
    damped.diagonal(dim1=1, dim2=2).add_((ridge_epsilon * 10.0 ** tries)[:, None])
    cond = root_diagnostics(roots, damped, ps_host if ps_host is not None else ps_t)
  s2 = torch.empty_like(s1)
  with torch.cuda.device(dev):
    _lobpcg_call(lib.pc_lobpcg_redeflate_prep, _ptr(packed), _ptr(scal), _ptr(eig), _ptr(ps_t), b, n,
                 k, _ptr(roots), _ptr(s2))
  final = batched_matmul(s2, s2, transpose_b=True, alpha=-1.0, c_in=roots, beta=1.0)  # DS:896-900
  metrics = metrics.clone()
  metrics[:, 3] = scal[:, 0]  # max_eigen_value: from the deflated eigenvalues (DS:815-816)
  if diagnostics:
    undamped = xs.clone()
    undamped.diagonal(dim1=1, dim2=2).add_(scal[:, 2][:, None])  # original + ridge I, DS:907
    uncond = root_diagnostics(final, undamped, ps_host if ps_host is not None else ps_t)
    vecs = packed[:, :, :k].contiguous()
    av = batched_matmul(xs, vecs)
    gram = batched_matmul(vecs.transpose(1, 2).contiguous(), vecs)
    lob = torch.empty((b, 7), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
      _lobpcg_call(lib.pc_lobpcg_diagnostics, _ptr(packed), _ptr(av), _ptr(gram), _ptr(eig), b, n, k,
                   float(iters), _ptr(lob))
    # DS:913-921: report the error of the unconditioned problem
    metrics[:, 0] = torch.maximum(uncond[:, 0], uncond[:, 2])
    diag = {"lobpcg": lob, "inverse_pth_root": uncond, "conditioned_inverse_pth_root": cond}
  return final, metrics, diag
