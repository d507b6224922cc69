"""Records and combinators of the tearfree front-end, named after
precondition/tearfree/praxis_shim.py (TF/praxis_shim.py:25-90) and optax: a gradient
transformation that also knows how its state is partitioned, the weight description used for
that, the empty optax states, and ``sharded_chain``."""
import dataclasses
from typing import Any, Callable, NamedTuple


class GradientTransformation(NamedTuple):
  """optax.GradientTransformation: ``init(params)``, ``update(updates, state, params)``."""
  init: Callable
  update: Callable


@dataclasses.dataclass(frozen=True)
class ShardedGradientTransformation:
  """A GradientTransformation plus ``init_partition_spec(params)`` (TF/praxis_shim.py:25-31)."""
  init: Callable
  update: Callable
  init_partition_spec: Any


class MaskedNode(NamedTuple):
  """optax.MaskedNode -- stands where a state has nothing to hold."""


class EmptyState(NamedTuple):
  """optax.EmptyState."""


class MaskedState(NamedTuple):
  """optax.MaskedState."""
  inner_state: Any


class TraceState(NamedTuple):
  """optax.TraceState: the momentum buffers."""
  trace: Any


class WeightHParams(NamedTuple):
  """What praxis says about one weight (TF/praxis_shim.py:37-42)."""
  shape: list
  init: Any
  dtype: Any
  collections: Any
  tensor_split_dims_mapping: list


NestedHParams = Any


class _Chain:
  """Transformations applied one after the other; the state is the tuple of their states."""

  def __init__(self, parts):
    self.parts = tuple(parts)

  def init(self, params):
    return tuple(part.init(params) for part in self.parts)

  def update(self, updates, state, params=None):
    if len(self.parts) != len(state):
      raise ValueError('The number of updates and states has to be the same in '
                       f'sharded chain. got {len(self.parts)=}, {len(state)=}')
    out = []
    for part, part_state in zip(self.parts, state):
      updates, part_state = part.update(updates, part_state, params)
      out.append(MaskedNode() if part_state is None else part_state)
    return updates, tuple(out)

  def init_partition_spec(self, mdl_vars):
    specs = []
    for part in self.parts:
      fn = getattr(part, 'init_partition_spec', None)
      if not callable(fn):
        raise ValueError('Attempting to use an optimizer in sharded_chain that '
                         'does not have an init_partition_spec.')
      specs.append(fn(mdl_vars))
    return MaskedState(inner_state=tuple(specs))


def sharded_chain(*args) -> ShardedGradientTransformation:
  """praxis.optimizers.sharded_chain (TF/praxis_shim.py:45-90)."""
  chain = _Chain(args)
  return ShardedGradientTransformation(init=chain.init, update=chain.update,
                                       init_partition_spec=chain.init_partition_spec)
