"""Mirror of precondition/tearfree/second_order.py (TF/second_order.py:29-102): choose and wrap
the second-order direction."""
import dataclasses
import enum
from typing import Any, Optional

from precondition_b200.tearfree import praxis_shim
from precondition_b200.tearfree import reshaper
from precondition_b200.tearfree import shampoo
from precondition_b200.tearfree import sketchy


@enum.unique
class SecondOrderType(enum.Enum):
  """Different second order covariance tracking methods (TF/second_order.py:29-33)."""
  SHAMPOO = 'shampoo'
  SKETCHY = 'sketchy'


@dataclasses.dataclass
class Options:
  """Which second order statistics to track (TF/second_order.py:36-52)."""
  merge_dims: int = 1024
  second_order_type: SecondOrderType = SecondOrderType.SHAMPOO
  shampoo_options: Optional[shampoo.Options] = dataclasses.field(
      default_factory=shampoo.Options)
  sketchy_options: Optional[sketchy.Options] = None


def _reshaper_options(options: Options) -> reshaper.Options:  # TF/second_order.py:78-89
  if options.second_order_type == SecondOrderType.SHAMPOO:
    assert options.shampoo_options
    return reshaper.Options(options.merge_dims, options.shampoo_options.block_size)
  if options.second_order_type == SecondOrderType.SKETCHY:
    return reshaper.Options(options.merge_dims, 0)
  raise ValueError('unknown second order type {}'.format(options.second_order_type))


def _update_stats_and_precondition(options: Options, _alias_outputs=False):  # TF/second_order.py:92-102
  if options.second_order_type == SecondOrderType.SHAMPOO:
    assert options.shampoo_options
    return shampoo.apply(options.shampoo_options, _alias_outputs)
  if options.second_order_type == SecondOrderType.SKETCHY:
    assert options.sketchy_options
    return sketchy.apply(options.sketchy_options, _alias_outputs)
  raise ValueError('unknown second order type {}'.format(options.second_order_type))


def apply(options: Options, _alias_outputs: bool = False
          ) -> praxis_shim.ShardedGradientTransformation:
  """The second order update: merge -> statistics / preconditioning -> unmerge
  (TF/second_order.py:55-75)."""
  reshaper_options = _reshaper_options(options)
  merge_tx = reshaper.merge(reshaper_options)
  precond_tx = _update_stats_and_precondition(options, _alias_outputs)

  def wrap_init(params):
    reshaped_params, _ = merge_tx.update(params, merge_tx.init(params), params)
    return precond_tx.init(reshaped_params)

  wrapped_precond_tx = praxis_shim.ShardedGradientTransformation(
      wrap_init, precond_tx.update, precond_tx.init_partition_spec)
  unmerge_tx = reshaper.unmerge(reshaper_options)
  as_sharded = lambda tx: praxis_shim.ShardedGradientTransformation(
      tx.init, tx.update, lambda params: praxis_shim.MaskedNode())
  return praxis_shim.sharded_chain(as_sharded(merge_tx), wrapped_precond_tx,
                                   as_sharded(unmerge_tx))
