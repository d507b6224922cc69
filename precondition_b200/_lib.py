"""ctypes binding of libprecond_b200.so (the C ABI in include/precond_b200.h).

There is NO fallback: if the shared library is missing or no CUDA device is
present, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprecond_b200.so")
CSRC = os.path.join(_HERE, "csrc")

PC_ENGINE_AUTO, PC_ENGINE_SIMT_FP32, PC_ENGINE_TC_BF16X6, PC_ENGINE_TC_BF16X3 = 0, 1, 2, 3
PC_ENGINE_TC_FP16X3 = 4
PC_ENGINE_TC_SMALL = 5
PC_TF_GRAFT_NONE, PC_TF_GRAFT_SGD, PC_TF_GRAFT_RMSPROP = 0, 1, 2
PC_QDTYPE_F32, PC_QDTYPE_INT16, PC_QDTYPE_INT8, PC_QDTYPE_BF16 = 0, 1, 2, 3
PC_NUM_METRICS = 5
PC_THIN_GEMV, PC_THIN_ROWMAP, PC_THIN_OUTER = 0, 1, 2
PC_MAX_PEERS = 16
PC_PEER_FLAG_WORDS = 2 * PC_MAX_PEERS + 16

EXPORTED_SYMBOLS = (
    "pc_version", "pc_last_error", "pc_device_supports_tcgen05", "pc_stats_reset",
    "pc_stats_get",
    "pc_root_options_default", "pc_resolve_engine",
    "pc_inverse_pth_root_workspace_bytes",
    "pc_inverse_pth_root_batched", "pc_debug_tc_gemm", "pc_power_iteration_batched", "pc_grouped_gemm", "pc_select_preconditioners",
    "pc_quantize_batched", "pc_dequantize_batched",
    "pc_graft_momentum_workspace_bytes", "pc_graft_momentum",
    "pc_fd_options_default", "pc_fd_update_workspace_bytes", "pc_fd_update_batched",
    "pc_low_rank_to_dense_workspace_bytes", "pc_low_rank_to_dense", "pc_low_rank_factors",
    "pc_grouped_gemm_tc_workspace_bytes", "pc_grouped_gemm_tc",
    "pc_low_rank_root_workspace_bytes", "pc_low_rank_root_batched",
    "pc_inverse_pth_root_eigh_batched",
    "pc_grouped_gemm_tc_quant", "pc_quantize_from_colmax_batched",
    "pc_grouped_gemm_splitk_workspace_bytes", "pc_grouped_gemm_splitk", "pc_grouped_gemm_thin",
    "pc_graft_group_chunk_elems", "pc_graft_momentum_grouped_workspace_bytes",
    "pc_graft_momentum_grouped", "pc_inverse_pth_root_enqueue", "pc_root_mode",
    "pc_select_scatter", "pc_ipc_export", "pc_ipc_open", "pc_peer_all_gather", "pc_peer_release",
    "pc_sm3_workspace_bytes", "pc_sm3_update",
    "pc_lobpcg_deflate_prep", "pc_lobpcg_redeflate_prep", "pc_root_diagnostics",
    "pc_lobpcg_diagnostics",
    "pc_pinv_pth_root_eigh_batched", "pc_pinv_pth_root_eigh_warm_batched",
    "pc_tearfree_transform_workspace_bytes",
    "pc_tearfree_transform",
    "pc_quant_group_chunk_elems", "pc_quant_group_tile_rows", "pc_dequantize_grouped",
    "pc_quantize_grouped",
)


class RootOptions(ctypes.Structure):
  _fields_ = [("ridge_epsilon", ctypes.c_float), ("error_tolerance", ctypes.c_float),
              ("num_iters", ctypes.c_int), ("relative_matrix_epsilon", ctypes.c_int),
              ("engine", ctypes.c_int), ("reserved", ctypes.c_int)]


class FdOptions(ctypes.Structure):
  _fields_ = [("ridge_epsilon", ctypes.c_float), ("error_tolerance", ctypes.c_float),
              ("relative_matrix_epsilon", ctypes.c_int), ("decay", ctypes.c_float),
              ("input_is_gram", ctypes.c_int), ("subspace_iters", ctypes.c_int),
              ("oversample", ctypes.c_int), ("full_eigh_max_dim", ctypes.c_int),
              ("tearfree", ctypes.c_int), ("tearfree_epsilon", ctypes.c_float),
              ("tearfree_relative_epsilon", ctypes.c_int)]


class Stats(ctypes.Structure):
  _fields_ = [("kernel_launches", ctypes.c_int64), ("gemm_launches", ctypes.c_int64),
              ("gemm_ms", ctypes.c_double), ("gemm_flops", ctypes.c_double)]


class GemmDesc(ctypes.Structure):
  _fields_ = [("a", ctypes.c_void_p), ("b", ctypes.c_void_p), ("c_in", ctypes.c_void_p),
              ("c", ctypes.c_void_p),
              ("a_sio", ctypes.c_int64), ("a_si", ctypes.c_int64), ("a_sko", ctypes.c_int64),
              ("a_ski", ctypes.c_int64),
              ("b_sj", ctypes.c_int64), ("b_sko", ctypes.c_int64), ("b_ski", ctypes.c_int64),
              ("c_sio", ctypes.c_int64), ("c_sii", ctypes.c_int64),
              ("a_iinner", ctypes.c_int32), ("a_kinner", ctypes.c_int32),
              ("b_kinner", ctypes.c_int32), ("c_iinner", ctypes.c_int32),
              ("m", ctypes.c_int32), ("n", ctypes.c_int32), ("k", ctypes.c_int32),
              ("alpha", ctypes.c_float), ("beta", ctypes.c_float),
              ("reserved", ctypes.c_int32), ("beta_dev", ctypes.c_void_p)]


class GemmQuant(ctypes.Structure):
  _fields_ = [("q_in", ctypes.c_void_p), ("diag_in", ctypes.c_void_p),
              ("bucket_in", ctypes.c_void_p), ("colmax_out", ctypes.c_void_p),
              ("qdtype", ctypes.c_int32), ("reserved", ctypes.c_int32),
              ("b_q", ctypes.c_void_p), ("b_diag", ctypes.c_void_p), ("b_bucket", ctypes.c_void_p),
              ("b_ld", ctypes.c_int32), ("b_qdtype", ctypes.c_int32)]


class GraftOptions(ctypes.Structure):
  _fields_ = [("beta1", ctypes.c_double), ("beta2", ctypes.c_double),
              ("graft_type", ctypes.c_int), ("diagonal_epsilon", ctypes.c_float),
              ("weight_decay", ctypes.c_float), ("learning_rate", ctypes.c_float),
              ("nesterov", ctypes.c_int), ("moving_average_for_momentum", ctypes.c_int),
              ("decoupled_learning_rate", ctypes.c_int),
              ("decoupled_weight_decay", ctypes.c_int), ("run_shampoo", ctypes.c_int),
              ("clip_by_scaled_gradient_norm", ctypes.c_float)]


class GraftSegment(ctypes.Structure):
  _fields_ = [("offset", ctypes.c_int64), ("numel", ctypes.c_int64),
              ("first_chunk", ctypes.c_int32), ("nchunks", ctypes.c_int32),
              ("has_precond", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class QuantSegment(ctypes.Structure):
  _fields_ = [("q", ctypes.c_void_p), ("bucket", ctypes.c_void_p), ("x", ctypes.c_void_p),
              ("colmax", ctypes.c_void_p), ("rows", ctypes.c_int32), ("cols", ctypes.c_int32),
              ("first_chunk", ctypes.c_int32), ("nchunks", ctypes.c_int32),
              ("first_tile", ctypes.c_int32), ("col_tiles", ctypes.c_int32)]


class TearfreeSegment(ctypes.Structure):
  _fields_ = [("grad", ctypes.c_void_p), ("param", ctypes.c_void_p), ("precond", ctypes.c_void_p),
              ("acc", ctypes.c_void_p), ("velocity", ctypes.c_void_p), ("update", ctypes.c_void_p),
              ("numel", ctypes.c_int64), ("first_chunk", ctypes.c_int32),
              ("nchunks", ctypes.c_int32)]


class TearfreeOptions(ctypes.Structure):
  _fields_ = [("graft_type", ctypes.c_int), ("graft_decay", ctypes.c_float),
              ("graft_epsilon", ctypes.c_float), ("use_precond", ctypes.c_int),
              ("ema", ctypes.c_int), ("nesterov", ctypes.c_int),
              ("momentum_decay", ctypes.c_float), ("weight_decay", ctypes.c_float),
              ("weight_decay_after_momentum", ctypes.c_int), ("scale", ctypes.c_float)]


class Sm3Options(ctypes.Structure):
  _fields_ = [("beta1", ctypes.c_double), ("beta2", ctypes.c_double),
              ("diagonal_epsilon", ctypes.c_float), ("weight_decay", ctypes.c_float),
              ("learning_rate", ctypes.c_float), ("normalize_grads", ctypes.c_int)]


class IpcHandle(ctypes.Structure):
  _fields_ = [("handle", ctypes.c_ubyte * 64), ("offset", ctypes.c_int64),
              ("size", ctypes.c_int64), ("device", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class PeerGroup(ctypes.Structure):
  _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("slot_bytes", ctypes.c_int64),
              ("recv", ctypes.c_void_p * PC_MAX_PEERS), ("flags", ctypes.c_void_p * PC_MAX_PEERS)]


_lib = None


def build(verbose: bool = False) -> str:
  """Compiles the CUDA library in-tree for sm_100a (nvcc cross-compiles on CPU)."""
  out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
  if verbose or out.returncode != 0:
    print(out.stdout[-4000:])
    print(out.stderr[-4000:])
  if out.returncode != 0:
    raise RuntimeError("building libprecond_b200.so failed")
  return LIB_PATH


def load() -> ctypes.CDLL:
  """Loads the library (no CUDA call is made here)."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(
        f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
        "(precondition_b200 has no CPU or PyTorch fallback)")
  lib = ctypes.CDLL(LIB_PATH)
  vp, i32, i64, f32, sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float,
                           ctypes.c_size_t)
  lib.pc_version.restype = i32
  lib.pc_last_error.restype = ctypes.c_char_p
  lib.pc_device_supports_tcgen05.restype = i32
  lib.pc_stats_reset.argtypes = [i32]
  lib.pc_stats_reset.restype = None
  lib.pc_stats_get.argtypes = [ctypes.POINTER(Stats)]
  lib.pc_stats_get.restype = None
  lib.pc_root_options_default.argtypes = [ctypes.POINTER(RootOptions)]
  lib.pc_root_options_default.restype = None
  lib.pc_resolve_engine.argtypes = [i32, i32]
  lib.pc_resolve_engine.restype = i32
  lib.pc_inverse_pth_root_workspace_bytes.argtypes = [i32, i32, i32]
  lib.pc_inverse_pth_root_workspace_bytes.restype = sz
  lib.pc_inverse_pth_root_batched.argtypes = [vp, vp, vp, i32, i32,
                                              ctypes.POINTER(RootOptions), vp, vp, vp, sz, vp]
  lib.pc_inverse_pth_root_batched.restype = i32
  lib.pc_inverse_pth_root_enqueue.argtypes = [vp, vp, vp, vp, i32, i32,
                                              ctypes.POINTER(RootOptions), vp, vp, vp, sz, vp]
  lib.pc_inverse_pth_root_enqueue.restype = i32
  lib.pc_root_mode.restype = i32
  lib.pc_debug_tc_gemm.argtypes = [vp, vp, vp, i32, i32, i32, vp, sz, vp]
  lib.pc_debug_tc_gemm.restype = i32
  lib.pc_power_iteration_batched.argtypes = [vp, vp, i32, i32, i32, f32, vp, vp, vp]
  lib.pc_power_iteration_batched.restype = i32
  lib.pc_grouped_gemm.argtypes = [vp, i32, i32, i32, vp]
  lib.pc_grouped_gemm.restype = i32
  lib.pc_select_preconditioners.argtypes = [vp, vp, f32, vp, i32, i32, i32, i32, i32, vp]
  lib.pc_select_preconditioners.restype = i32
  lib.pc_select_scatter.argtypes = [vp, vp, vp, vp, vp, f32, vp, i64, vp, i32, vp]
  lib.pc_select_scatter.restype = i32
  lib.pc_lobpcg_deflate_prep.argtypes = [vp, vp, i32, i32, i32, f32, i32, vp, vp, vp, vp, vp]
  lib.pc_lobpcg_deflate_prep.restype = i32
  lib.pc_lobpcg_redeflate_prep.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp]
  lib.pc_lobpcg_redeflate_prep.restype = i32
  lib.pc_root_diagnostics.argtypes = [vp, i32, i32, vp, vp]
  lib.pc_root_diagnostics.restype = i32
  lib.pc_lobpcg_diagnostics.argtypes = [vp, vp, vp, vp, i32, i32, i32, f32, vp, vp]
  lib.pc_lobpcg_diagnostics.restype = i32
  lib.pc_sm3_workspace_bytes.argtypes = [i64]
  lib.pc_sm3_workspace_bytes.restype = sz
  lib.pc_sm3_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, vp,
                                ctypes.POINTER(Sm3Options), vp, sz, vp]
  lib.pc_sm3_update.restype = i32
  lib.pc_ipc_export.argtypes = [vp, ctypes.POINTER(IpcHandle)]
  lib.pc_ipc_export.restype = i32
  lib.pc_ipc_open.argtypes = [ctypes.POINTER(IpcHandle), ctypes.POINTER(ctypes.c_void_p)]
  lib.pc_ipc_open.restype = i32
  lib.pc_peer_all_gather.argtypes = [ctypes.POINTER(PeerGroup), vp, sz, ctypes.c_uint32, vp]
  lib.pc_peer_all_gather.restype = i32
  lib.pc_peer_release.argtypes = [ctypes.POINTER(PeerGroup), ctypes.c_uint32, vp]
  lib.pc_peer_release.restype = i32
  lib.pc_quantize_batched.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp]
  lib.pc_quantize_batched.restype = i32
  lib.pc_dequantize_batched.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
  lib.pc_dequantize_batched.restype = i32
  lib.pc_graft_momentum_workspace_bytes.argtypes = [i64]
  lib.pc_graft_momentum_workspace_bytes.restype = sz
  lib.pc_graft_momentum.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64,
                                    ctypes.POINTER(GraftOptions), vp, sz, vp]
  lib.pc_graft_momentum.restype = i32
  lib.pc_fd_options_default.argtypes = [ctypes.POINTER(FdOptions)]
  lib.pc_fd_options_default.restype = None
  lib.pc_fd_update_workspace_bytes.argtypes = [i32, i32, i32, i32, ctypes.POINTER(FdOptions)]
  lib.pc_fd_update_workspace_bytes.restype = sz
  lib.pc_fd_update_batched.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32,
                                       ctypes.POINTER(FdOptions), vp, vp, vp, sz, vp]
  lib.pc_fd_update_batched.restype = i32
  lib.pc_low_rank_to_dense_workspace_bytes.argtypes = [i32, i32, i32]
  lib.pc_low_rank_to_dense_workspace_bytes.restype = sz
  lib.pc_low_rank_to_dense.argtypes = [vp, i32, i32, i32, vp, vp, sz, vp]
  lib.pc_low_rank_to_dense.restype = i32
  lib.pc_low_rank_factors.argtypes = [vp, i32, i32, i32, vp, vp, vp]
  lib.pc_low_rank_factors.restype = i32
  lib.pc_grouped_gemm_tc_workspace_bytes.argtypes = [vp, i32]
  lib.pc_grouped_gemm_tc_workspace_bytes.restype = sz
  lib.pc_grouped_gemm_tc.argtypes = [vp, i32, vp, sz, i32, vp]
  lib.pc_grouped_gemm_tc.restype = i32
  lib.pc_low_rank_root_workspace_bytes.argtypes = [i32, i32]
  lib.pc_low_rank_root_workspace_bytes.restype = sz
  lib.pc_low_rank_root_batched.argtypes = [vp, vp, vp, i32, i32, i32, f32, f32, i32, vp, vp, vp,
                                           sz, vp]
  lib.pc_low_rank_root_batched.restype = i32
  lib.pc_inverse_pth_root_eigh_batched.argtypes = [vp, vp, vp, i32, i32, f32, f32, i32, vp, vp, vp,
                                                   sz, vp]
  lib.pc_inverse_pth_root_eigh_batched.restype = i32
  lib.pc_grouped_gemm_tc_quant.argtypes = [vp, vp, i32, vp, sz, i32, vp]
  lib.pc_grouped_gemm_tc_quant.restype = i32
  lib.pc_quantize_from_colmax_batched.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
  lib.pc_quantize_from_colmax_batched.restype = i32
  lib.pc_grouped_gemm_splitk_workspace_bytes.argtypes = [i32, i32, i32, i32]
  lib.pc_grouped_gemm_splitk_workspace_bytes.restype = sz
  lib.pc_grouped_gemm_splitk.argtypes = [vp, i32, i32, i32, i32, vp, sz, vp]
  lib.pc_grouped_gemm_splitk.restype = i32
  lib.pc_grouped_gemm_thin.argtypes = [vp, i32, i32, i32, i32, vp]
  lib.pc_grouped_gemm_thin.restype = i32
  lib.pc_graft_group_chunk_elems.argtypes = []
  lib.pc_graft_group_chunk_elems.restype = i64
  lib.pc_graft_momentum_grouped_workspace_bytes.argtypes = [i32, i64]
  lib.pc_graft_momentum_grouped_workspace_bytes.restype = sz
  lib.pc_graft_momentum_grouped.argtypes = [vp, vp, i32, i64, vp, vp, vp, vp, vp, vp, vp,
                                            ctypes.POINTER(GraftOptions), vp, sz, vp]
  lib.pc_graft_momentum_grouped.restype = i32
  lib.pc_quant_group_chunk_elems.argtypes = []
  lib.pc_quant_group_chunk_elems.restype = i64
  lib.pc_dequantize_grouped.argtypes = [vp, vp, i32, i64, vp]
  lib.pc_dequantize_grouped.restype = i32
  lib.pc_quant_group_tile_rows.argtypes = []
  lib.pc_quant_group_tile_rows.restype = i32
  lib.pc_quantize_grouped.argtypes = [vp, vp, i32, i64, vp, i64, vp, sz, vp]
  lib.pc_quantize_grouped.restype = i32
  lib.pc_pinv_pth_root_eigh_batched.argtypes = [vp, vp, i32, i32, f32, vp, vp, sz, vp]
  lib.pc_pinv_pth_root_eigh_batched.restype = i32
  lib.pc_pinv_pth_root_eigh_warm_batched.argtypes = [vp, vp, i32, i32, f32, vp, vp, i32, vp, sz, vp]
  lib.pc_pinv_pth_root_eigh_warm_batched.restype = i32
  lib.pc_tearfree_transform_workspace_bytes.argtypes = [i32, i64]
  lib.pc_tearfree_transform_workspace_bytes.restype = sz
  lib.pc_tearfree_transform.argtypes = [vp, vp, i32, i64, ctypes.POINTER(TearfreeOptions), vp, sz,
                                        vp]
  lib.pc_tearfree_transform.restype = i32
  _lib = lib
  return lib


def check(rc: int):
  if rc != 0:
    msg = load().pc_last_error().decode("utf-8", "replace")
    raise RuntimeError(f"precond_b200 error {rc}: {msg}")
