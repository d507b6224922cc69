"""Minimal pytrees for the tearfree front-end: dict / list / tuple / NamedTuple containers, every
other object (tensors, masks, metadata) is a leaf."""
from typing import Any, Callable


def _is_container(x) -> bool:
  return isinstance(x, (dict, list, tuple))


def tree_map(f: Callable, tree: Any, *rest: Any, is_leaf: Callable = None) -> Any:
  """``jax.tree.map`` over the first tree's structure; the other trees only need to have that
  structure as a prefix (what sits at a leaf position is passed whole)."""
  if (is_leaf is not None and is_leaf(tree)) or not _is_container(tree):
    return f(tree, *rest)
  if isinstance(tree, dict):
    return {k: tree_map(f, tree[k], *[r[k] for r in rest], is_leaf=is_leaf) for k in tree}
  items = [tree_map(f, v, *[r[i] for r in rest], is_leaf=is_leaf) for i, v in enumerate(tree)]
  if hasattr(tree, "_fields"):
    return type(tree)(*items)
  return type(tree)(items)


def tree_map_with_path(f: Callable, tree: Any, *rest: Any, is_leaf: Callable = None,
                       _path=()) -> Any:
  if (is_leaf is not None and is_leaf(tree)) or not _is_container(tree):
    return f(_path, tree, *rest)
  if isinstance(tree, dict):
    return {k: tree_map_with_path(f, tree[k], *[r[k] for r in rest], is_leaf=is_leaf,
                                  _path=_path + (k,)) for k in tree}
  items = [tree_map_with_path(f, v, *[r[i] for r in rest], is_leaf=is_leaf, _path=_path + (i,))
           for i, v in enumerate(tree)]
  if hasattr(tree, "_fields"):
    return type(tree)(*items)
  return type(tree)(items)


def tree_leaves(tree: Any, is_leaf: Callable = None) -> list:
  out = []
  tree_map(lambda x: out.append(x), tree, is_leaf=is_leaf)
  return out
