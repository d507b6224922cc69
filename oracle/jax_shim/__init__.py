"""numpy-backed stand-in for the sliver of jax / flax / chex / optax that
``/root/reference/precondition/{distributed_shampoo,quantization_utils}.py``
import -- TEST INFRASTRUCTURE, build-container only.

Purpose: JAX cannot be installed here (no wheel, no network), so the reference
cannot run on XLA.  ``install()`` registers fake ``jax``, ``jax.numpy``,
``jax.lax``, ``flax.struct``, ``chex`` and ``optax`` modules so that the
UNMODIFIED reference sources can be imported and executed with numpy as the
array backend.  ``oracle/gen_golden.py`` uses that to record golden vectors under
``tests/golden/``.  Semantics kept faithful to JAX where they affect numerics:

* ``jax_enable_x64`` off (default): float64/int64 are canonicalised to
  float32/int32 on every array-producing call and every ufunc result, and
  ``jnp.float64`` *is* ``float32`` -- so the root routine runs in fp32 exactly as
  the reference does in practice (DS:35-38).  ``install(x64=True)`` gives the
  float64 twin.
* Python scalars are weakly typed (numpy>=2 NEP 50 matches JAX here).
* ``lax.while_loop`` / ``lax.cond`` run as Python control flow; ``jax.vmap`` maps
  a Python loop over the leading axis, i.e. each matrix is processed as if
  un-batched -- which is what batched ``while_loop`` semantics guarantee.
* ``jax.pmap`` runs one thread per "device"; ``lax.all_gather`` is a barrier
  exchange, ``lax.psum(1, axis)`` the device count, ``lax.axis_index`` the rank.

GEMMs are numpy (BLAS sgemm) instead of XLA's Eigen contraction: same fp32
arithmetic, different summation order.  Nothing here is imported by the product.
"""
from __future__ import annotations

import collections
import dataclasses
import enum
import sys
import threading
import types

import numpy as np

_X64 = False


# ----------------------------------------------------------------------------
# array type
# ----------------------------------------------------------------------------
def _canon_dtype(dt):
  dt = np.dtype(dt)
  if not _X64:
    if dt == np.float64:
      return np.dtype(np.float32)
    if dt == np.int64:
      return np.dtype(np.int32)
    if dt == np.uint64:
      return np.dtype(np.uint32)
    if dt == np.complex128:
      return np.dtype(np.complex64)
  return dt


class _At:

  def __init__(self, arr):
    self._arr = arr

  def __getitem__(self, idx):
    return _AtIdx(self._arr, idx)


class _AtIdx:

  def __init__(self, arr, idx):
    self._arr, self._idx = arr, idx

  def set(self, value):
    out = np.array(self._arr, copy=True)
    out[self._idx] = value
    return _wrap_out(out)

  def add(self, value):
    out = np.array(self._arr, copy=True)
    out[self._idx] += value
    return _wrap_out(out)


class Arr(np.ndarray):
  """ndarray with ``.at[...]`` and JAX-style dtype canonicalisation of results."""

  @property
  def at(self):
    return _At(self)

  def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
    ins = tuple(np.asarray(x) if isinstance(x, Arr) else x for x in inputs)
    if out is not None:
      kwargs["out"] = tuple(np.asarray(o) if isinstance(o, Arr) else o for o in out)
    res = getattr(ufunc, method)(*ins, **kwargs)
    if out is not None:
      return out[0] if len(out) == 1 else out
    return _wrap_out(res)

  def astype(self, dtype, *a, **k):
    return _wrap_out(np.asarray(self).astype(_canon_dtype(_resolve_dtype(dtype)), *a, **k))

  def __getitem__(self, idx):
    return _wrap_out(np.asarray(self)[idx])

  def __iter__(self):
    base = np.asarray(self)
    for i in range(base.shape[0]):
      yield _wrap_out(base[i])

  def __hash__(self):
    return id(self)


def _resolve_dtype(dt):
  return dt


def _wrap_out(x):
  if isinstance(x, tuple) and hasattr(x, "_fields"):  # numpy linalg result tuples
    return tuple(_wrap_out(v) for v in x)
  if isinstance(x, tuple):
    return tuple(_wrap_out(v) for v in x)
  if isinstance(x, list):
    return [_wrap_out(v) for v in x]
  if isinstance(x, (np.ndarray, np.generic)):
    a = np.asarray(x)
    cd = _canon_dtype(a.dtype) if a.dtype.kind in "fiuc" else a.dtype
    if cd != a.dtype:
      a = a.astype(cd)
    return a.view(Arr)
  return x


def _np_fn(fn, drop=("precision",)):

  def wrapped(*args, **kwargs):
    for k in drop:
      kwargs.pop(k, None)
    if "dtype" in kwargs and kwargs["dtype"] is not None:
      kwargs["dtype"] = _canon_dtype(kwargs["dtype"])
    if isinstance(kwargs.get("axis"), list):  # jnp accepts a list of axes, numpy wants a tuple
      kwargs["axis"] = tuple(kwargs["axis"])
    args = tuple(np.asarray(a) if isinstance(a, Arr) else a for a in args)
    return _wrap_out(fn(*args, **kwargs))

  wrapped.__name__ = getattr(fn, "__name__", "fn")
  return wrapped


# ----------------------------------------------------------------------------
# pytrees
# ----------------------------------------------------------------------------
_LEAF = object()


class TreeDef:

  def __init__(self, kind, meta, children):
    self.kind, self.meta, self.children = kind, meta, children

  @property
  def num_leaves(self):
    if self.kind is _LEAF:
      return 1
    return sum(c.num_leaves for c in self.children)

  def flatten_up_to(self, tree):
    out = []
    _flatten_up_to(self, tree, out)
    return out

  def unflatten(self, leaves):
    it = iter(leaves)
    return _unflatten(self, it)


def _is_namedtuple(x):
  return isinstance(x, tuple) and hasattr(x, "_fields")


def _node_children(x):
  """Returns (kind, meta, children) or None if x is a leaf."""
  if x is None:
    return ("none", None, [])
  if _is_namedtuple(x):
    return ("namedtuple", type(x), list(x))
  if isinstance(x, tuple):
    return ("tuple", None, list(x))
  if isinstance(x, list):
    return ("list", None, list(x))
  if isinstance(x, dict):
    keys = sorted(x.keys())
    return ("dict", keys, [x[k] for k in keys])
  if dataclasses.is_dataclass(x) and not isinstance(x, type) and getattr(
      type(x), "_shim_struct", False):
    node_fields = [f.name for f in dataclasses.fields(x)
                   if f.metadata.get("pytree_node", True)]
    static = {f.name: getattr(x, f.name) for f in dataclasses.fields(x)
              if not f.metadata.get("pytree_node", True)}
    return ("struct", (type(x), node_fields, static),
            [getattr(x, n) for n in node_fields])
  return None


def _flatten(x, leaves, is_leaf=None):
  if is_leaf is not None and is_leaf(x):
    leaves.append(x)
    return TreeDef(_LEAF, None, [])
  node = _node_children(x)
  if node is None:
    leaves.append(x)
    return TreeDef(_LEAF, None, [])
  kind, meta, children = node
  return TreeDef(kind, meta, [_flatten(c, leaves, is_leaf) for c in children])


def _flatten_up_to(td, x, out):
  if td.kind is _LEAF:
    out.append(x)
    return
  node = _node_children(x)
  assert node is not None and node[0] == td.kind, (td.kind, type(x))
  assert len(node[2]) == len(td.children)
  for c_td, c in zip(td.children, node[2]):
    _flatten_up_to(c_td, c, out)


def _unflatten(td, it):
  if td.kind is _LEAF:
    return next(it)
  ch = [_unflatten(c, it) for c in td.children]
  if td.kind == "none":
    return None
  if td.kind == "namedtuple":
    return td.meta(*ch)
  if td.kind == "tuple":
    return tuple(ch)
  if td.kind == "list":
    return ch
  if td.kind == "dict":
    return dict(zip(td.meta, ch))
  if td.kind == "struct":
    cls, names, static = td.meta
    return cls(**dict(zip(names, ch)), **static)
  raise TypeError(td.kind)


def tree_flatten(tree, is_leaf=None):
  leaves = []
  td = _flatten(tree, leaves, is_leaf)
  return leaves, td


def tree_unflatten(td, leaves):
  return td.unflatten(leaves)


def tree_map(f, tree, *rest, is_leaf=None):
  leaves, td = tree_flatten(tree, is_leaf)
  others = [td.flatten_up_to(r) for r in rest]
  return td.unflatten([f(*xs) for xs in zip(leaves, *others)])


def tree_leaves(tree, is_leaf=None):
  return tree_flatten(tree, is_leaf)[0]


class _Key:
  """Path entry of tree_map_with_path: `.key` for dict / attribute nodes, `.idx` for sequences."""

  def __init__(self, key=None, idx=None):
    if key is not None:
      self.key = key
    if idx is not None:
      self.idx = idx

  def __repr__(self):
    return f"[{self.key!r}]" if hasattr(self, "key") else f"[{self.idx}]"


def _paths(x, prefix, out, is_leaf=None):
  if is_leaf is not None and is_leaf(x):
    out.append(prefix)
    return
  node = _node_children(x)
  if node is None:
    out.append(prefix)
    return
  kind, meta, children = node
  if kind == "dict":
    keys = [_Key(key=k) for k in meta]
  elif kind == "namedtuple":
    keys = [_Key(key=k) for k in meta._fields]
  elif kind == "struct":
    keys = [_Key(key=k) for k in meta[1]]
  else:
    keys = [_Key(idx=i) for i in range(len(children))]
  for k, c in zip(keys, children):
    _paths(c, prefix + (k,), out, is_leaf)


def tree_map_with_path(f, tree, *rest, is_leaf=None):
  leaves, td = tree_flatten(tree, is_leaf)
  paths = []
  _paths(tree, (), paths, is_leaf)
  others = [td.flatten_up_to(r) for r in rest]
  return td.unflatten([f(p, *xs) for p, xs in zip(paths, zip(leaves, *others))])


def tree_all(tree):
  return all(bool(x) for x in tree_leaves(tree))


# ----------------------------------------------------------------------------
# lax
# ----------------------------------------------------------------------------
class Precision(enum.Enum):
  DEFAULT = 0
  HIGH = 1
  HIGHEST = 2


def while_loop(cond_fun, body_fun, init_val):
  # JAX turns Python-float carries into (weak) f32 arrays; ints/bools are left
  # as Python values so that mixed int/float arithmetic stays weakly typed.
  val = tree_map(lambda x: _wrap_out(np.asarray(x, dtype=np.float64))
                 if isinstance(x, float) else x, init_val)
  while bool(cond_fun(val)):
    val = body_fun(val)
  return val


_NOTHING = object()


def cond(pred, true_fun, false_fun, *operands, operand=_NOTHING):
  if operand is not _NOTHING:
    operands = (operand,)
  return true_fun(*operands) if bool(pred) else false_fun(*operands)


class _AxisEnv(threading.local):
  name = None
  index = 0
  size = 1
  shared = None


_axis = _AxisEnv()


def psum(x, axis_name):
  assert _axis.name == axis_name, "psum outside pmap"
  if isinstance(x, (int, float)):
    return x * _axis.size
  return _wrap_out(np.stack(_exchange(np.asarray(x))).sum(0))


def axis_index(axis_name):
  assert _axis.name == axis_name
  return _axis.index


def _exchange(value):
  sh = _axis.shared
  if _axis.size == 1:
    return [value]
  sh["slots"][_axis.index] = value
  sh["barrier"].wait()
  vals = list(sh["slots"])
  sh["barrier"].wait()
  return vals


def all_gather(x, axis_name):
  assert _axis.name == axis_name
  leaves, td = tree_flatten(x)
  out = []
  for leaf in leaves:
    out.append(_wrap_out(np.stack([np.asarray(v) for v in _exchange(np.asarray(leaf))])))
  return td.unflatten(out)


def with_sharding_constraint(x, *a, **k):
  return x


def pmap(fn, axis_name=None, **unused):

  def run(*args):
    size = tree_leaves(args)[0].shape[0]
    shared = {"slots": [None] * size, "barrier": threading.Barrier(size)}
    results, errors = [None] * size, []

    def worker(r):
      _axis.name, _axis.index, _axis.size, _axis.shared = axis_name, r, size, shared
      try:
        results[r] = fn(*tree_map(lambda a: a[r], args))
      except BaseException as e:  # pylint: disable=broad-except
        errors.append(e)
        shared["barrier"].abort()
      finally:
        _axis.name = None

    if size == 1:
      worker(0)
    else:
      ts = [threading.Thread(target=worker, args=(r,)) for r in range(size)]
      [t.start() for t in ts]
      [t.join() for t in ts]
    if errors:
      raise errors[0]
    return tree_map(lambda *xs: _wrap_out(np.stack([np.asarray(x) for x in xs])),
                    results[0], *results[1:])

  return run


def vmap(fn, in_axes=0, out_axes=0):

  def _ix(a, i):
    if a is None:
      return None
    v = np.asarray(a)[i]
    if v.ndim == 0 and v.dtype.kind in "iub":
      return int(v) if v.dtype.kind != "b" else bool(v)
    return _wrap_out(v)

  def run(*args, **kwargs):
    if isinstance(in_axes, int) and in_axes != 0:  # one mapped axis for every positional array
      args = tuple(_wrap_out(np.moveaxis(np.asarray(a), in_axes, 0)) for a in args)
    sized = [a for a in list(args) + list(kwargs.values()) if a is not None]
    b = len(sized[0])
    outs = []
    for i in range(b):
      outs.append(fn(*[tree_map(lambda a: _ix(a, i), x) if x is not None else None
                       for x in args],
                     **{k: (tree_map(lambda a: _ix(a, i), v) if v is not None else None)
                        for k, v in kwargs.items()}))
    return tree_map(lambda *xs: _wrap_out(np.stack([np.asarray(x) for x in xs])),
                    outs[0], *outs[1:])

  return run


# ----------------------------------------------------------------------------
# flax.struct
# ----------------------------------------------------------------------------
def struct_field(pytree_node=True, **kwargs):
  md = dict(kwargs.pop("metadata", {}) or {})
  md["pytree_node"] = pytree_node
  return dataclasses.field(metadata=md, **kwargs)


def struct_dataclass(cls=None, **kw):

  def wrap(c):
    dc = dataclasses.dataclass(frozen=True)(c)
    dc._shim_struct = True

    def replace(self, **updates):
      return dataclasses.replace(self, **updates)

    dc.replace = replace
    return dc

  return wrap(cls) if cls is not None else wrap


# ----------------------------------------------------------------------------
# install
# ----------------------------------------------------------------------------
def _make_jnp():
  jnp = types.ModuleType("jax.numpy")
  names = """where array eye stack max abs diag zeros sqrt matmul maximum asarray
  zeros_like sum arange concatenate square flip tensordot reshape logical_or
  split min logical_and isnan squeeze power ones_like transpose roll pad mean
  einsum any trace sign round repeat moveaxis log expm1 log1p expand_dims greater
  ones minimum dot exp cumsum argsort sort outer tril triu all full
  count_nonzero clip isfinite floor ceil prod linspace identity allclose
  array_equal take diagonal argmax argmin nan_to_num""".split()
  for n in names:
    setattr(jnp, n, _np_fn(getattr(np, n)))
  jnp.linalg = types.ModuleType("jax.numpy.linalg")
  for n in "norm eigh svd qr eigvalsh inv cholesky pinv det cond matrix_power".split():
    setattr(jnp.linalg, n, _np_fn(getattr(np.linalg, n)))
  jnp.ndarray = np.ndarray
  jnp.newaxis = None
  jnp.float32 = np.float32
  jnp.float64 = np.float64 if _X64 else np.float32
  jnp.float16 = np.float16
  jnp.int8, jnp.int16, jnp.int32 = np.int8, np.int16, np.int32
  jnp.int64 = np.int64 if _X64 else np.int32
  jnp.uint8, jnp.bool_ = np.uint8, np.bool_
  jnp.inf, jnp.pi, jnp.nan = np.inf, np.pi, np.nan
  jnp.dtype = np.dtype
  try:
    import ml_dtypes
    jnp.bfloat16 = ml_dtypes.bfloat16
  except ImportError:  # pragma: no cover
    jnp.bfloat16 = None
  return jnp


def _install_optax(optax, masked_node):
  """The handful of optax transformations precondition/tearfree composes (optax 0.2 semantics:
  `trace` = g + decay * t, Nesterov = g + decay * new_trace; `add_decayed_weights` = g + wd * p;
  `scale_by_schedule` multiplies by schedule(count) and counts up)."""
  GT = optax.GradientTransformation
  EmptyState = collections.namedtuple("EmptyState", [])
  TraceState = collections.namedtuple("TraceState", ["trace"])
  MaskedState = collections.namedtuple("MaskedState", ["inner_state"])
  ScaleByScheduleState = collections.namedtuple("ScaleByScheduleState", ["count"])
  optax.EmptyState, optax.TraceState, optax.MaskedState = EmptyState, TraceState, MaskedState
  optax.ScaleByScheduleState = ScaleByScheduleState
  for alias in ("Updates", "Params", "OptState", "TransformInitFn", "TransformUpdateFn",
                "Schedule"):
    setattr(optax, alias, object)

  def identity():
    return GT(lambda params: EmptyState(), lambda u, s, p=None: (u, s))

  def scale(step_size):
    return GT(lambda params: EmptyState(),
              lambda u, s, p=None: (tree_map(lambda g: step_size * g, u), s))

  def scale_by_schedule(step_size_fn):
    def update(u, s, p=None):
      step = step_size_fn(s.count)
      return tree_map(lambda g: _wrap_out(np.asarray(step, np.float32) * g), u), (
          ScaleByScheduleState(count=s.count + 1))
    return GT(lambda params: ScaleByScheduleState(count=_wrap_out(np.zeros([], np.int32))),
              update)

  def add_decayed_weights(weight_decay=0.0, mask=None):
    assert mask is None
    return GT(lambda params: EmptyState(),
              lambda u, s, p=None: (tree_map(lambda g, w: g + weight_decay * w, u, p), s))

  def trace(decay, nesterov=False, accumulator_dtype=None):
    def init(params):
      return TraceState(trace=tree_map(lambda x: _wrap_out(np.zeros_like(np.asarray(x))), params))

    def update(u, s, p=None):
      f = lambda g, t: g + decay * t
      new_trace = tree_map(f, u, s.trace)
      out = tree_map(f, u, new_trace) if nesterov else new_trace
      return out, TraceState(trace=new_trace)
    return GT(init, update)

  def chain(*txs):
    def init(params):
      return tuple(t.init(params) for t in txs)

    def update(u, state, p=None):
      new = []
      for t, st in zip(txs, state):
        u, st = t.update(u, st, p)
        new.append(st)
      return u, tuple(new)
    return GT(init, update)

  def adafactor(*a, **k):
    raise NotImplementedError("optax.adafactor is not part of the shim")

  optax.identity, optax.scale, optax.scale_by_schedule = identity, scale, scale_by_schedule
  optax.add_decayed_weights, optax.trace, optax.chain = add_decayed_weights, trace, chain
  optax.adafactor = adafactor


def install(x64: bool = False):
  """Registers the fake modules.  Call before importing the reference."""
  global _X64
  _X64 = bool(x64)
  for k in list(sys.modules):
    if k.split(".")[0] in ("jax", "flax", "chex", "optax", "precondition"):
      del sys.modules[k]

  jax = types.ModuleType("jax")
  jnp = _make_jnp()
  lax = types.ModuleType("jax.lax")
  lax.Precision = Precision
  lax.while_loop, lax.cond = while_loop, cond
  lax.psum, lax.axis_index, lax.all_gather = psum, axis_index, all_gather
  lax.with_sharding_constraint = with_sharding_constraint
  lax.rsqrt = lambda x: _wrap_out(1.0 / np.sqrt(np.asarray(x)))
  jax.numpy, jax.lax = jnp, lax
  jax.vmap, jax.pmap = vmap, pmap
  import contextlib
  jax.named_scope = lambda name: contextlib.nullcontext()
  jax.jit = lambda f, **k: f
  jax.Array = np.ndarray
  tree = types.ModuleType("jax.tree")
  tree.map, tree.flatten, tree.unflatten, tree.leaves = (
      tree_map, tree_flatten, tree_unflatten, tree_leaves)
  jax.tree = tree
  tu = types.ModuleType("jax.tree_util")
  tu.tree_all, tu.tree_map, tu.tree_flatten = tree_all, tree_map, tree_flatten
  tu.tree_unflatten, tu.tree_leaves = tree_unflatten, tree_leaves
  tu.tree_map_with_path = tree_map_with_path
  jax.tree_util = tu
  sharding = types.ModuleType("jax.sharding")
  sharding.PartitionSpec = lambda *a: tuple(a)
  jax.sharding = sharding
  experimental = types.ModuleType("jax.experimental")
  sparse = types.ModuleType("jax.experimental.sparse")
  splinalg = types.ModuleType("jax.experimental.sparse.linalg")

  def _no_lobpcg(*a, **k):
    raise NotImplementedError("lobpcg is outside the hot path")

  splinalg.lobpcg_standard = _no_lobpcg
  sparse.linalg = splinalg
  experimental.sparse = sparse
  jax.experimental = experimental
  config = types.SimpleNamespace(update=lambda *a, **k: None)
  jax.config = config

  flax = types.ModuleType("flax")
  struct = types.ModuleType("flax.struct")
  struct.dataclass, struct.field = struct_dataclass, struct_field
  flax.struct = struct

  chex = types.ModuleType("chex")
  chex.Array = np.ndarray
  chex.Numeric = float
  chex.ArrayTree = object

  optax = types.ModuleType("optax")
  optax.GradientTransformation = collections.namedtuple(
      "GradientTransformation", ["init", "update"])

  class MaskedNode(tuple):
    """Empty pytree node (optax.MaskedNode is an empty NamedTuple)."""
    _fields = ()
    __slots__ = ()

    def __new__(cls):
      return tuple.__new__(cls)

  optax.MaskedNode = MaskedNode
  _install_optax(optax, MaskedNode)

  mods = {
      "jax": jax, "jax.numpy": jnp, "jax.lax": lax, "jax.tree": tree,
      "jax.tree_util": tu, "jax.sharding": sharding,
      "jax.experimental": experimental, "jax.experimental.sparse": sparse,
      "jax.experimental.sparse.linalg": splinalg, "flax": flax,
      "flax.struct": struct, "chex": chex, "optax": optax,
  }
  sys.modules.update(mods)
  return jax


def import_reference(root="/root/reference", x64=False, extra=()):
  """Imports the unmodified reference ``distributed_shampoo`` over the shim."""
  install(x64=x64)
  # ``precondition/__init__.py`` is empty apart from __version__; import the two
  # hot-path modules directly so nothing else (sm3, tearfree, ...) is pulled in.
  import importlib.util
  pkg = types.ModuleType("precondition")
  pkg.__path__ = [root + "/precondition"]
  sys.modules["precondition"] = pkg
  out = {}
  for name in ("quantization_utils", "distributed_shampoo") + tuple(extra):
    spec = importlib.util.spec_from_file_location(
        f"precondition.{name}", f"{root}/precondition/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"precondition.{name}"] = mod
    spec.loader.exec_module(mod)
    out[name] = mod
  if extra:
    return (out["distributed_shampoo"], out["quantization_utils"]) + tuple(out[e] for e in extra)
  return out["distributed_shampoo"], out["quantization_utils"]
