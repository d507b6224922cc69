"""Mirror of precondition/tearfree/grafting.py (TF/grafting.py:31-342): take the direction of a
second-order update and the norm of a first-order one."""
import copy
import dataclasses
import enum
from typing import Any, NamedTuple

import torch

from precondition_b200 import _lib
from precondition_b200.tearfree import _tail
from precondition_b200.tearfree import _tree
from precondition_b200.tearfree import praxis_shim


@enum.unique
class GraftingType(enum.Enum):
  """Different grafting types (TF/grafting.py:31-38)."""
  NONE = 'none'
  SGD = 'sgd'
  RMSPROP = 'rmsprop'
  ADAFACTOR = 'adafactor'


def _zeros_f32(p: torch.Tensor) -> torch.Tensor:
  """State is fp32 and contiguous whatever the parameter's dtype / memory format."""
  return torch.zeros(p.shape, dtype=torch.float32, device=p.device)


@dataclasses.dataclass
class Options:
  """Grafting configuration (TF/grafting.py:41-86); same fields and defaults."""
  grafting_type: GraftingType = GraftingType.RMSPROP
  second_moment_decay: float = 0.999
  start_preconditioning_step: int = 0
  epsilon: float = 1e-23
  skip_preconditioning_any_dim_gt: int = 4096
  skip_preconditioning_rank1: bool = True
  min_dim_size_to_factor: int = 128
  multiply_by_parameter_scale: float = True
  clipping_threshold: float = 1.0


_KERNEL_TYPE = {GraftingType.NONE: _lib.PC_TF_GRAFT_NONE, GraftingType.SGD: _lib.PC_TF_GRAFT_SGD,
                GraftingType.RMSPROP: _lib.PC_TF_GRAFT_RMSPROP}


def _validate(options: Options):  # TF/grafting.py:131-163
  if options.grafting_type in [GraftingType.RMSPROP, GraftingType.ADAFACTOR]:
    if options.epsilon < 0:
      raise ValueError('epsilon ({}) should be non-negative'.format(options.epsilon))
  if options.grafting_type == GraftingType.RMSPROP:
    if not (0 < options.second_moment_decay <= 1.0):
      raise ValueError('second_moment_decay ({}) not in (0, 1] for graft ({})'.format(
          options.second_moment_decay, options.grafting_type))
  if options.grafting_type == GraftingType.ADAFACTOR:
    if not (0 < options.second_moment_decay < 1.0):
      raise ValueError('second_moment_decay ({}) not in (0, 1) for graft ({})'.format(
          options.second_moment_decay, options.grafting_type))
    if not (0 < options.min_dim_size_to_factor):
      raise ValueError('min_dim_size_to_factor ({}) should be positive for graft ({})'.format(
          options.min_dim_size_to_factor, options.grafting_type))
    if options.clipping_threshold < 1:
      raise ValueError('clipping_threshold ({}) should be >= 1 for graft ({})'.format(
          options.clipping_threshold, options.grafting_type))
    raise NotImplementedError(
        'ADAFACTOR grafting (optax.adafactor, TF/grafting.py:176-194) is not built; use RMSPROP, '
        'SGD or NONE')


class RMSPropAccumulator(NamedTuple):
  """State holding the sum/ema of gradient squares so far (TF/grafting.py:199-202)."""
  acc: Any


class GraftingState(NamedTuple):
  """count, the direction's state and the norm's state (TF/grafting.py:236-241)."""
  count: torch.Tensor
  direction: Any
  norm: Any


class _GraftMask(NamedTuple):
  """Stands in for a parameter that gets no second-order direction (an empty pytree node, like
  the reference's empty struct, TF/grafting.py:317-321)."""


def _mask_skipped(options: Options, tree):  # TF/grafting.py:324-336
  def _maybe_mask(x):
    if options.skip_preconditioning_rank1 and x.ndim <= 1:
      return _GraftMask()
    if any(s > options.skip_preconditioning_any_dim_gt for s in x.shape):
      return _GraftMask()
    return x
  return _tree.tree_map(_maybe_mask, tree)


def _masked(node) -> bool:
  return isinstance(node, _GraftMask)


def norm_init(options: Options, params):
  """State of the first-order optimizer whose norm is grafted."""
  if options.grafting_type == GraftingType.RMSPROP:
    return RMSPropAccumulator(acc=_tree.tree_map(_zeros_f32, params))
  return praxis_shim.EmptyState()


def graft(options: Options,
          direction: praxis_shim.ShardedGradientTransformation
          ) -> praxis_shim.ShardedGradientTransformation:
  """The grafting update from options and a direction update (TF/grafting.py:89-128)."""
  _validate(options)
  if options.grafting_type == GraftingType.NONE:
    return direction
  tail = _tail.Tail()

  def init_fn(params):
    return GraftingState(count=torch.zeros([], dtype=torch.int32),
                         direction=direction.init(_mask_skipped(options, params)),
                         norm=norm_init(options, params))

  def update_fn(updates, state, params=None):
    base_updates, base_state = direction.update(
        _mask_skipped(options, updates), state.direction,
        None if params is None else _mask_skipped(options, params))
    outs = run_tail(tail, options, updates, base_updates, state, None)
    new_state = GraftingState(count=state.count + 1, direction=base_state, norm=state.norm)
    it = iter(outs)
    return _tree.tree_map(lambda _: next(it), updates), new_state

  def init_partition_spec_fn(mdl_params):  # TF/grafting.py:296-309
    count_pspec = praxis_shim.WeightHParams(shape=[], init=None, dtype=torch.int32,
                                            collections=None, tensor_split_dims_mapping=[])
    if options.grafting_type == GraftingType.RMSPROP:
      def _spec(v):
        s = copy.deepcopy(v)
        return s._replace(init=None) if hasattr(s, "_replace") else s
      norm = RMSPropAccumulator(acc=_tree.tree_map(_spec, mdl_params,
                                                   is_leaf=lambda x: hasattr(x, "shape")))
    else:
      norm = praxis_shim.EmptyState()
    return dict(count=count_pspec, direction=direction.init_partition_spec(mdl_params),
                norm=norm)

  return praxis_shim.ShardedGradientTransformation(init_fn, update_fn, init_partition_spec_fn)


def run_tail(tail, options: Options, updates, base_updates, state: GraftingState, params,
             **tail_kwargs):
  """One ``pc_tearfree_transform`` call for the leaves of ``updates``; ``base_updates`` has a
  ``_GraftMask`` where the direction was skipped.  Extra keyword arguments switch on the
  momentum / weight-decay / learning-rate stages (``optimizer.tearfree`` fuses them in)."""
  leaves = _tree.tree_leaves(updates)
  bases = _tree.tree_leaves(base_updates, is_leaf=_masked)
  assert len(bases) == len(leaves)
  accs = None
  if options.grafting_type == GraftingType.RMSPROP:
    accs = _tree.tree_leaves(state.norm.acc)
  use_precond = int(state.count) >= options.start_preconditioning_step  # TF/grafting.py:268
  return tail.run(leaves, None if params is None else _tree.tree_leaves(params),
                  [None if _masked(b) else b for b in bases], accs,
                  tail_kwargs.pop("velocities", None),
                  graft_type=_KERNEL_TYPE[options.grafting_type],
                  graft_decay=options.second_moment_decay, graft_epsilon=options.epsilon,
                  use_precond=use_precond, **tail_kwargs)
