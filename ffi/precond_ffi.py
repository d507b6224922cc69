"""Registers the XLA FFI handlers of ffi/precond_ffi.cc with JAX and shows the call sites a
maintainer of precondition/distributed_shampoo.py (DS) would change.

Needs `jax` (not installed in this image: importing this module there raises ImportError, and
nothing in precondition_b200/ depends on it).  Build the adapter with
`python -c "import __graft_entry__ as g; g.build()"` (it compiles ffi/libprecond_b200_ffi.so
when `jax.ffi.include_dir()` exists)."""
import ctypes
import os

import jax
import jax.numpy as jnp
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_core = ctypes.CDLL(os.path.join(_HERE, "..", "precondition_b200", "libprecond_b200.so"),
                    mode=ctypes.RTLD_GLOBAL)
_ffi = ctypes.CDLL(os.path.join(_HERE, "libprecond_b200_ffi.so"))
_core.pc_inverse_pth_root_workspace_bytes.restype = ctypes.c_size_t
_core.pc_inverse_pth_root_workspace_bytes.argtypes = [ctypes.c_int] * 3

for _name, _sym in (("pc_inverse_root", "PcInverseRoot"), ("pc_inverse_root_eigh", "PcInverseRootEigh"),
                    ("pc_low_rank_root", "PcLowRankRoot"), ("pc_power_iteration", "PcPowerIteration"),
                    ("pc_fd_update", "PcFdUpdate"), ("pc_low_rank_to_dense", "PcLowRankToDense"),
                    ("pc_low_rank_factors", "PcLowRankFactors"),
                    ("pc_pinv_pth_root_eigh", "PcPinvRootEigh"),
                    ("pc_select_preconditioners", "PcSelectPreconditioners"),
                    ("pc_quantize_int16", "PcQuantizeInt16"), ("pc_quantize_int8", "PcQuantizeInt8"),
                    ("pc_dequantize_int16", "PcDequantizeInt16"),
                    ("pc_dequantize_int8", "PcDequantizeInt8"), ("pc_gram_update", "PcGramUpdate"),
                    ("pc_graft_momentum", "PcGraftMomentum")):
  jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_ffi, _sym)), platform="CUDA")


def matrix_inverse_pth_root_vmap(xs, ps, padding_starts, ridge_epsilon=1e-6, error_tolerance=1e-6,
                                 num_iters=100, relative_matrix_epsilon=True, engine=0):
  """Drop-in body of `_matrix_inverse_pth_root_vmap` (DS:2742-2744): one custom call for the
  whole batch; returns (roots [b,n,n], metrics [b,5]) -- metrics columns are the
  TrainingMetrics scalars of DS:902-907 in declaration order."""
  b, n, _ = xs.shape
  ws = int(_core.pc_inverse_pth_root_workspace_bytes(b, n, engine))
  roots, metrics, _ = jax.ffi.ffi_call(
      "pc_inverse_root",
      (jax.ShapeDtypeStruct(xs.shape, jnp.float32), jax.ShapeDtypeStruct((b, 5), jnp.float32),
       jax.ShapeDtypeStruct((ws,), jnp.uint8)))(
           xs.astype(jnp.float32), ps.astype(jnp.int32), padding_starts.astype(jnp.int32),
           ridge_epsilon=np.float32(ridge_epsilon), error_tolerance=np.float32(error_tolerance),
           num_iters=np.int32(num_iters),
           relative_matrix_epsilon=np.int32(relative_matrix_epsilon), engine=np.int32(engine))
  return roots, metrics
