"""Mirror of precondition/tearfree/reshaper.py: merge small dimensions and pad large ones to a
multiple of the block size before the second-order transformation, undo it after
(TF/reshaper.py:27-133).  Pure layout work: reshapes are views, padding is one copy."""
import dataclasses
import functools

import torch

from precondition_b200 import distributed_shampoo
from precondition_b200.tearfree import _tree
from precondition_b200.tearfree import praxis_shim


@dataclasses.dataclass
class Options:
  """Parameter reshaping options (TF/reshaper.py:27-40)."""
  merge_dims: int = 1024
  block_size: int = 1024


@dataclasses.dataclass
class _Shapes:
  original_shape: list
  merged_shape: list
  padded_shape: list


def _derive_shapes(options: Options, param) -> _Shapes:  # TF/reshaper.py:52-77
  merged = distributed_shampoo.merge_small_dims(list(param.shape), options.merge_dims)
  if merged == [1]:
    return _Shapes(original_shape=list(param.shape), merged_shape=[], padded_shape=[])
  if options.block_size == 0:
    padded = merged
  else:
    padded = []
    for s in merged:
      if s >= options.block_size:
        s = (s + options.block_size - 1) // options.block_size
        s *= options.block_size
      padded.append(s)
  return _Shapes(original_shape=list(param.shape), merged_shape=merged, padded_shape=padded)


def merge(options: Options) -> praxis_shim.GradientTransformation:
  """Merge and maybe pad gradients, leaving params alone (TF/reshaper.py:80-112)."""
  if options.merge_dims < 2:
    raise ValueError('merge_dims ({}) must be at least 2'.format(options.merge_dims))
  if options.block_size < 2 and options.block_size != 0:
    raise ValueError('block_size ({}) must be at least 2 (or 0 to disable)'.format(
        options.block_size))

  def _merge(update: torch.Tensor, shapes: _Shapes) -> torch.Tensor:
    assert list(update.shape) == shapes.original_shape, (update.shape, shapes)
    merged = update.reshape(shapes.merged_shape)
    if options.block_size > 0 and shapes.padded_shape != shapes.merged_shape:
      pad = []
      for p, m in reversed(list(zip(shapes.padded_shape, shapes.merged_shape))):
        pad += [0, p - m]
      return torch.nn.functional.pad(merged, pad)
    return merged

  def update(updates, state, params):
    shapes = _tree.tree_map(functools.partial(_derive_shapes, options), params)
    return _tree.tree_map(_merge, updates, shapes), state

  return praxis_shim.GradientTransformation(lambda _: praxis_shim.MaskedNode(), update)


def unmerge(options: Options) -> praxis_shim.GradientTransformation:
  """Unmerge and unpad gradients, leaving params alone (TF/reshaper.py:115-133)."""

  def _unmerge(update: torch.Tensor, shapes: _Shapes) -> torch.Tensor:
    assert list(update.shape) == shapes.padded_shape, (update.shape, shapes)
    if options.block_size == 0:
      merged = update
    else:
      merged = update[tuple(slice(0, m) for m in shapes.merged_shape)]
    return merged.reshape(shapes.original_shape)

  def update(updates, state, params):
    shapes = _tree.tree_map(functools.partial(_derive_shapes, options), params)
    return _tree.tree_map(_unmerge, updates, shapes), state

  return praxis_shim.GradientTransformation(lambda _: praxis_shim.MaskedNode(), update)
