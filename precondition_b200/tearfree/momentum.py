"""Mirror of precondition/tearfree/momentum.py (TF/momentum.py:26-139): EMA / trace momentum with
optional Nesterov acceleration, and weight decay before or after it."""
import copy
import dataclasses
from typing import NamedTuple

import torch

from precondition_b200.tearfree import _tail
from precondition_b200.tearfree import _tree
from precondition_b200.tearfree import praxis_shim


def _zeros_f32(p: torch.Tensor) -> torch.Tensor:
  """State is fp32 and contiguous whatever the parameter's dtype / memory format."""
  return torch.zeros(p.shape, dtype=torch.float32, device=p.device)


@dataclasses.dataclass
class Options:
  """Momentum (and weight decay) options, TF/momentum.py:26-77:
  ema: velocity' = decay * velocity + (1 - decay) * update, else trace accumulation
    velocity' = decay * velocity + update (distributed_shampoo's moving_average_for_momentum);
  nesterov: update' = maybe_decay * update + decay * velocity' (maybe_decay = 1 - decay if ema);
  weight_decay: adds weight_decay * param, after the momentum transformation or before it."""
  ema: bool = False
  nesterov: bool = True
  momentum_decay: float = 0.9
  weight_decay: float = 0.0
  weight_decay_after_momentum: bool = True


def _validate(options: Options):  # TF/momentum.py:106-117
  if not (0 <= options.momentum_decay <= 1):
    raise ValueError('momentum_decay ({}) must be in [0, 1]'.format(options.momentum_decay))
  if not (options.weight_decay >= 0):
    raise ValueError('weight_decay ({}) must be >= 0'.format(options.weight_decay))


class _State(NamedTuple):
  """States of the chained parts in the reference's order (a TraceState where momentum is on,
  EmptyStates for the stateless scale / weight-decay parts)."""


def state_layout(options: Options):
  """Kinds of the chained transformations, in order (TF/momentum.py:84-103)."""
  mom = (["scale"] if options.ema else []) + ["trace"] if options.momentum_decay else []
  wd = ["wd"] if options.weight_decay > 0.0 else []
  return mom + wd if options.weight_decay_after_momentum else wd + mom


def init_state(options: Options, params):
  out = []
  for kind in state_layout(options):
    if kind == "trace":
      out.append(praxis_shim.TraceState(trace=_tree.tree_map(_zeros_f32, params)))
    else:
      out.append(praxis_shim.EmptyState())
  return tuple(out)


def trace_of(options: Options, state):
  """The velocity tree inside a momentum state (None when momentum is off)."""
  for kind, st in zip(state_layout(options), state):
    if kind == "trace":
      return st.trace
  return None


def apply(options: Options) -> praxis_shim.ShardedGradientTransformation:
  """Generate the momentum update from options (TF/momentum.py:81-103)."""
  _validate(options)
  tail = _tail.Tail()

  def init_fn(params):
    return init_state(options, params)

  def update_fn(updates, state, params=None):
    leaves = _tree.tree_leaves(updates)
    vel = trace_of(options, state)
    outs = tail.run(leaves, None if params is None else _tree.tree_leaves(params), None, None,
                    None if vel is None else _tree.tree_leaves(vel), ema=options.ema,
                    nesterov=options.nesterov, momentum_decay=options.momentum_decay,
                    weight_decay=options.weight_decay,
                    weight_decay_after_momentum=options.weight_decay_after_momentum)
    it = iter(outs)
    return _tree.tree_map(lambda _: next(it), updates), state

  def init_pspec_fn(mdl_params):  # TF/momentum.py:126-134
    def _spec(var_hparams):
      s = copy.deepcopy(var_hparams)
      return s._replace(init=None) if hasattr(s, "_replace") else s
    out = []
    for kind in state_layout(options):
      if kind == "trace":
        out.append(praxis_shim.TraceState(trace=_tree.tree_map(
            _spec, mdl_params, is_leaf=lambda x: hasattr(x, "shape"))))
      else:
        out.append(praxis_shim.EmptyState())
    return praxis_shim.MaskedState(inner_state=tuple(out))

  return praxis_shim.ShardedGradientTransformation(init_fn, update_fn, init_pspec_fn)
