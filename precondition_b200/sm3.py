"""SM3 optimizer over the B200 library -- host-side mirror of ``precondition/sm3.py``
(reference lines 40-168, "SM3" below): same factory signature, same state layout
(``SM3State(count, stats)``, per parameter ``ParameterStats(diagonal_statistics = one
accumulator vector per axis, diagonal_momentum = int8 QuantizedValue)``).  Parameters and
gradients are pytrees of CUDA tensors; the per-parameter update is one ``pc_sm3_update`` launch
plus the int8 requantisation of the momentum (``pc_quantize_batched``).  No CPU fallback."""
from __future__ import annotations

import ctypes
from typing import Any, NamedTuple

import torch

from precondition_b200 import _lib, ops
from precondition_b200.distributed_shampoo import (GradientTransformation, _tree_flatten,
                                                   _tree_unflatten)
from precondition_b200.quantization_utils import QuantizedValue


class SM3State(NamedTuple):  # SM3:28-30
  count: int
  stats: Any


class ParameterStats(NamedTuple):  # SM3:33-37
  diagonal_statistics: Any   # list of per-axis accumulators
  diagonal_momentum: QuantizedValue


def _flatten_stats(tree):
  out = []

  def rec(t):
    if isinstance(t, ParameterStats):
      out.append(t)
    elif isinstance(t, dict):
      for k in sorted(t.keys()):
        rec(t[k])
    else:
      for v in t:
        rec(v)

  rec(tree)
  return out


def sm3(learning_rate, beta1=0.9, beta2=0.999, diagonal_epsilon=1e-10, weight_decay=0.0,
        normalize_grads=False):
  """SM3 optimizer (Anil, Gupta, Koren, Singer; https://arxiv.org/abs/1901.11150)."""
  lib = _lib.load()

  def init_fn(params):  # SM3:71-80
    leaves, treedef = _tree_flatten(params)
    stats = []
    for p in leaves:
      if not p.is_cuda:
        raise RuntimeError("precondition_b200 needs CUDA tensors: there is no CPU fallback")
      if not 1 <= p.dim() <= 4:
        raise NotImplementedError(f"sm3: tensors of rank 1..4 are supported, got {tuple(p.shape)}")
      acc = [torch.zeros(s, dtype=torch.float32, device=p.device) for s in p.shape]
      mom = QuantizedValue.from_float_value(
          torch.zeros(p.shape, dtype=torch.float32, device=p.device), torch.int8)
      stats.append(ParameterStats(acc, mom))
    return SM3State(0, _tree_unflatten(treedef, stats))

  def update_fn(updates, state, params):
    g_leaves, treedef = _tree_flatten(updates)
    s_leaves = _flatten_stats(state.stats)
    p_leaves = _tree_flatten(params)[0] if params is not None else [None] * len(g_leaves)
    step = int(state.count)
    lr = learning_rate(step) if callable(learning_rate) else learning_rate
    opt = _lib.Sm3Options()
    opt.beta1, opt.beta2 = float(beta1), float(beta2)
    opt.diagonal_epsilon, opt.weight_decay = float(diagonal_epsilon), float(weight_decay)
    opt.learning_rate, opt.normalize_grads = float(lr), int(bool(normalize_grads))
    new_updates, new_stats = [], []
    for g, st, p in zip(g_leaves, s_leaves, p_leaves):
      if weight_decay > 0.0 and p is None:
        raise ValueError("weight_decay needs params")
      dev = g.device
      gf = g.to(torch.float32).contiguous()
      pf = None if p is None else p.to(torch.float32).contiguous()
      rank = g.dim()
      acc_out = [torch.empty_like(a) for a in st.diagonal_statistics]
      mom_f = torch.empty(g.shape, dtype=torch.float32, device=dev)
      upd = torch.empty(g.shape, dtype=torch.float32, device=dev)
      ptrs_in = (ctypes.c_void_p * 4)(*[a.data_ptr() for a in st.diagonal_statistics])
      ptrs_out = (ctypes.c_void_p * 4)(*[a.data_ptr() for a in acc_out])
      dims = (ctypes.c_int32 * 4)(*g.shape)
      mq = st.diagonal_momentum
      ws = ops._workspace(lib.pc_sm3_workspace_bytes(g.numel()), dev)
      with torch.cuda.device(dev):
        _lib.check(lib.pc_sm3_update(
            ctypes.c_void_p(gf.data_ptr()), None if pf is None else ctypes.c_void_p(pf.data_ptr()),
            ctypes.cast(ptrs_in, ctypes.c_void_p), ctypes.cast(ptrs_out, ctypes.c_void_p),
            ctypes.c_void_p(mq.quantized.data_ptr()),
            ctypes.c_void_p(mq.bucket_size.reshape(-1).data_ptr()
                            if mq.bucket_size.dim() else mq.bucket_size.data_ptr()),
            ctypes.c_void_p(mom_f.data_ptr()), ctypes.c_void_p(upd.data_ptr()), rank, dims,
            ctypes.byref(opt), ctypes.c_void_p(ws.data_ptr()), ws.numel(),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
      ops.gpu_launches += 1
      new_stats.append(ParameterStats(acc_out, QuantizedValue.from_float_value(mom_f, torch.int8)))
      new_updates.append(upd if g.dtype == torch.float32 else upd.to(g.dtype))
    return (_tree_unflatten(treedef, new_updates),
            SM3State(step + 1, _tree_unflatten(treedef, new_stats)))

  return GradientTransformation(init_fn, update_fn)
