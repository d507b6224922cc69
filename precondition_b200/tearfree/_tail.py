"""Shared launcher of ``pc_tearfree_transform`` (include/precond_b200.h): grafting, momentum,
weight decay and learning rate for all parameters in three launches.  Every public
transformation of this package that is element-wise runs through here -- alone with the other
stages switched off, or fused when ``optimizer.tearfree`` owns the whole chain."""
from typing import Optional, Sequence

import torch

from precondition_b200 import _lib
from precondition_b200 import ops


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
  if not isinstance(t, torch.Tensor) or not t.is_cuda:
    raise RuntimeError(f"tearfree: {what} must be CUDA tensors (no CPU fallback)")
  if not t.is_floating_point():
    raise TypeError(f"tearfree: {what} must be floating point, got {t.dtype}")
  if t.dtype != torch.float32:  # half-precision leaves are widened for the kernel
    t = t.to(torch.float32)
  return t if t.is_contiguous() else t.contiguous()


class Tail:
  """Caches one launch list per model (keyed by the element counts of the leaves)."""

  def __init__(self):
    self._lists = {}

  def run(self, grads: Sequence[torch.Tensor], params: Optional[Sequence],
          preconds: Optional[Sequence], accs: Optional[Sequence], velocities: Optional[Sequence],
          graft_type: int = _lib.PC_TF_GRAFT_NONE, graft_decay: float = 0.0,
          graft_epsilon: float = 0.0, use_precond: bool = True, ema: bool = False,
          nesterov: bool = False, momentum_decay: float = 0.0, weight_decay: float = 0.0,
          weight_decay_after_momentum: bool = True, scale: float = 1.0):
    """Returns the list of update tensors (new, contiguous); ``accs`` / ``velocities`` are
    updated in place."""
    if not grads:
      return []
    dtypes = [g.dtype for g in grads]
    grads = [_f32c(g, "updates") for g in grads]
    if weight_decay > 0.0:
      if params is None or any(p is None for p in params):
        raise ValueError("tearfree: weight decay needs the parameters")
      params = [_f32c(p, "params") for p in params]
    else:
      params = None
    if preconds is not None:
      preconds = [None if p is None else _f32c(p, "directions") for p in preconds]
    dev = grads[0].device
    key = (dev, tuple(int(g.numel()) for g in grads))
    lst = self._lists.get(key)
    if lst is None:
      lst = self._lists[key] = ops.TearfreeTail(key[1], dev)
    outs = [torch.empty_like(g) for g in grads]
    opt = _lib.TearfreeOptions()
    opt.graft_type = graft_type
    opt.graft_decay = graft_decay
    opt.graft_epsilon = graft_epsilon
    opt.use_precond = int(use_precond)
    opt.ema = int(ema)
    opt.nesterov = int(nesterov)
    opt.momentum_decay = momentum_decay
    opt.weight_decay = weight_decay
    opt.weight_decay_after_momentum = int(weight_decay_after_momentum)
    opt.scale = scale
    lst.run(grads, params, preconds, accs if graft_type == _lib.PC_TF_GRAFT_RMSPROP else None,
            velocities if momentum_decay != 0.0 else None, outs, opt)
    return [o if dt == torch.float32 else o.to(dt) for o, dt in zip(outs, dtypes)]
