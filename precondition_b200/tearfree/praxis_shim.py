"""Mirror of precondition/tearfree/praxis_shim.py: the sharded transformation record and
``sharded_chain`` (TF/praxis_shim.py:25-90)."""
import dataclasses
from typing import Any, Callable, NamedTuple


@dataclasses.dataclass(frozen=True)
class ShardedGradientTransformation:
  """GradientTransformation that also describes how its state is sharded (TF/praxis_shim.py:25-31)."""
  init: Callable
  update: Callable
  init_partition_spec: Any


class GradientTransformation(NamedTuple):
  """optax.GradientTransformation."""
  init: Callable
  update: Callable


class MaskedNode(NamedTuple):
  """optax.MaskedNode: the empty state."""


class EmptyState(NamedTuple):
  """optax.EmptyState."""


class MaskedState(NamedTuple):
  """optax.MaskedState."""
  inner_state: Any


class TraceState(NamedTuple):
  """optax.TraceState."""
  trace: Any


NestedHParams = Any


class WeightHParams(NamedTuple):  # TF/praxis_shim.py:37-42
  shape: list
  init: Any
  dtype: Any
  collections: Any
  tensor_split_dims_mapping: list


def sharded_chain(*args) -> ShardedGradientTransformation:
  """Chain as in praxis.optimizers.sharded_chain (TF/praxis_shim.py:45-90)."""

  def init_fn(params):
    return tuple(fn.init(params) for fn in args)

  def update_fn(updates, state, params=None):
    if len(args) != len(state):
      raise ValueError('The number of updates and states has to be the same in '
                       f'sharded chain. got {len(args)=}, {len(state)=}')
    new_state = []
    for s, fn in zip(state, args):
      updates, new_s = fn.update(updates, s, params)
      new_state.append(MaskedNode() if new_s is None else new_s)
    return updates, tuple(new_state)

  def init_partition_spec_fn(mdl_vars):
    partition_specs = []
    for fn in args:
      init_partition_spec = getattr(fn, 'init_partition_spec', None)
      if callable(init_partition_spec):
        partition_specs.append(init_partition_spec(mdl_vars))
      else:
        raise ValueError('Attempting to use an optimizer in sharded_chain that '
                         'does not have an init_partition_spec.')
    return MaskedState(inner_state=tuple(partition_specs))

  return ShardedGradientTransformation(init=init_fn, update=update_fn,
                                       init_partition_spec=init_partition_spec_fn)
