"""Mirror of precondition/tearfree/optimizer.py (TF/optimizer.py:37-99): preconditioned, grafted
updates with momentum, times the negative learning rate.

The state has the reference's nesting -- ``(graft state, momentum states, learning-rate state)``
-- but the three element-wise stages run as ONE fused call of ``pc_tearfree_transform`` per step
(three launches for the whole model) instead of one pass per stage; the stand-alone
transformations of ``grafting`` / ``momentum`` use the same kernel with the other stages off."""
import dataclasses
from typing import Callable, NamedTuple, Union

import torch

from precondition_b200.tearfree import _tail
from precondition_b200.tearfree import _tree
from precondition_b200.tearfree import grafting
from precondition_b200.tearfree import momentum
from precondition_b200.tearfree import praxis_shim
from precondition_b200.tearfree import second_order


@dataclasses.dataclass
class TearfreeOptions:
  """Configuration dataclass for the tearfree optimizer (TF/optimizer.py:37-58)."""
  grafting_options: grafting.Options = dataclasses.field(default_factory=grafting.Options)
  second_order_options: second_order.Options = dataclasses.field(
      default_factory=second_order.Options)
  momentum_options: momentum.Options = dataclasses.field(default_factory=momentum.Options)


class ScaleByScheduleState(NamedTuple):
  """optax.ScaleByScheduleState (learning-rate schedules count their steps)."""
  count: torch.Tensor


def tearfree(learning_rate: Union[float, Callable[[int], float]],
             options: TearfreeOptions) -> praxis_shim.ShardedGradientTransformation:
  """Tearfree optimizer (TF/optimizer.py:61-99): ``chain(graft(second_order), momentum, -lr)``."""
  gopt, mopt = options.grafting_options, options.momentum_options
  second_order_tx = second_order.apply(options.second_order_options, _alias_outputs=True)
  grafting._validate(gopt)
  momentum._validate(mopt)
  none = gopt.grafting_type == grafting.GraftingType.NONE
  tail = _tail.Tail()

  def lr_state():
    if callable(learning_rate):
      return ScaleByScheduleState(count=torch.zeros([], dtype=torch.int32))
    return praxis_shim.EmptyState()

  def init_fn(params):
    if none:
      graft_state = second_order_tx.init(params)
    else:
      graft_state = grafting.GraftingState(
          count=torch.zeros([], dtype=torch.int32),
          direction=second_order_tx.init(grafting._mask_skipped(gopt, params)),
          norm=grafting.norm_init(gopt, params))
    return (graft_state, momentum.init_state(mopt, params), lr_state())

  def update_fn(updates, state, params=None):
    graft_state, mom_state, lr_st = state
    if callable(learning_rate):
      lr = float(learning_rate(int(lr_st.count)))
      lr_st = ScaleByScheduleState(count=lr_st.count + 1)
    else:
      lr = float(learning_rate)
    vel = momentum.trace_of(mopt, mom_state)
    kw = dict(velocities=None if vel is None else _tree.tree_leaves(vel), ema=mopt.ema,
              nesterov=mopt.nesterov, momentum_decay=mopt.momentum_decay,
              weight_decay=mopt.weight_decay,
              weight_decay_after_momentum=mopt.weight_decay_after_momentum, scale=-1.0 * lr)
    if none:
      base, new_graft = second_order_tx.update(updates, graft_state, params)
      outs = tail.run(_tree.tree_leaves(updates),
                      None if params is None else _tree.tree_leaves(params),
                      _tree.tree_leaves(base), None, kw.pop("velocities"), **kw)
    else:
      base, base_state = second_order_tx.update(
          grafting._mask_skipped(gopt, updates), graft_state.direction,
          None if params is None else grafting._mask_skipped(gopt, params))
      outs = grafting.run_tail(tail, gopt, updates, base, graft_state, params, **kw)
      new_graft = grafting.GraftingState(count=graft_state.count + 1, direction=base_state,
                                         norm=graft_state.norm)
    it = iter(outs)
    return _tree.tree_map(lambda _: next(it), updates), (new_graft, mom_state, lr_st)

  def init_partition_spec_fn(mdl_params):
    graft_tx = grafting.graft(gopt, second_order_tx)
    return praxis_shim.MaskedState(inner_state=(
        graft_tx.init_partition_spec(mdl_params),
        momentum.apply(mopt).init_partition_spec(mdl_params),
        praxis_shim.MaskedNode()))

  return praxis_shim.ShardedGradientTransformation(init_fn, update_fn, init_partition_spec_fn)
