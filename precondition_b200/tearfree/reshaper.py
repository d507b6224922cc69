"""Layout stage of the tearfree front-end (the role of precondition/tearfree/reshaper.py,
TF/reshaper.py:27-133): before the second-order transformation every gradient is viewed with its
small dimensions merged and its large dimensions zero-padded up to a multiple of the block size;
afterwards the padding is cut off and the original shape restored.  Pure layout work on the
device: merging is a view, padding one copy."""
import dataclasses
from typing import List, Tuple

import torch

from precondition_b200 import distributed_shampoo
from precondition_b200.tearfree import _tree
from precondition_b200.tearfree import praxis_shim


@dataclasses.dataclass
class Options:
  """merge_dims: dimensions are collapsed left to right while their product stays <= merge_dims
  ([3, 1, 5, 2, 2] -> [3, 5, 4] at 4).  block_size: dimensions >= block_size are padded to a
  multiple of it; 0 switches padding off.  (TF/reshaper.py:27-40)"""
  merge_dims: int = 1024
  block_size: int = 1024


class _Layout:
  """The three shapes of one leaf: as given, merged, merged and padded."""

  def __init__(self, options: Options, shape):
    self.given: List[int] = [int(d) for d in shape]
    merged = distributed_shampoo.merge_small_dims(self.given, options.merge_dims)
    # a tensor that merges down to a single element is carried as a scalar (TF/reshaper.py:55-60)
    self.merged: List[int] = [] if merged == [1] else list(merged)
    bs = options.block_size
    self.padded: List[int] = [(-(-d // bs)) * bs if bs and d >= bs else d for d in self.merged]

  @property
  def grows(self) -> bool:
    return self.padded != self.merged

  def to_padded(self, x: torch.Tensor) -> torch.Tensor:
    if list(x.shape) != self.given:
      raise ValueError(f"update of shape {tuple(x.shape)} where {tuple(self.given)} was expected")
    y = x.reshape(self.merged)
    if not self.grows:
      return y
    # torch pads from the last dimension backwards: (left, right) pairs
    spec: Tuple[int, ...] = ()
    for have, want in zip(reversed(self.merged), reversed(self.padded)):
      spec += (0, want - have)
    return torch.nn.functional.pad(y, spec)

  def from_padded(self, y: torch.Tensor) -> torch.Tensor:
    if list(y.shape) != self.padded:
      raise ValueError(f"update of shape {tuple(y.shape)} where {tuple(self.padded)} was expected")
    if self.grows:
      y = y[tuple(slice(0, d) for d in self.merged)]
    return y.reshape(self.given)


def _check(options: Options) -> None:
  if options.merge_dims < 2:
    raise ValueError('merge_dims ({}) must be at least 2'.format(options.merge_dims))
  if options.block_size < 2 and options.block_size != 0:
    raise ValueError('block_size ({}) must be at least 2 (or 0 to disable)'.format(
        options.block_size))


def _stage(options: Options, forward: bool) -> praxis_shim.GradientTransformation:
  def update(updates, state, params):
    # the layout follows the PARAMETER shapes; the state is empty
    def one(u, p):
      layout = _Layout(options, p.shape)
      return layout.to_padded(u) if forward else layout.from_padded(u)
    return _tree.tree_map(one, updates, params), state

  return praxis_shim.GradientTransformation(lambda _: praxis_shim.MaskedNode(), update)


def merge(options: Options) -> praxis_shim.GradientTransformation:
  """Gradients to their merged (and padded) form; parameters are left alone."""
  _check(options)
  return _stage(options, True)


def unmerge(options: Options) -> praxis_shim.GradientTransformation:
  """Inverse of ``merge``."""
  return _stage(options, False)
