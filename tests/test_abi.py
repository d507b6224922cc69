"""CPU-only checks of the C-ABI library: it builds, loads and exports every
symbol declared in include/precond_b200.h (no compute calls without a GPU)."""
import os
import re

from precondition_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(ROOT, "include", "precond_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(pc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
  if not os.path.exists(_lib.LIB_PATH):
    _lib.build()
  lib = _lib.load()
  declared = _declared_symbols()
  assert declared, "no symbols parsed from the header"
  for sym in declared:
    assert hasattr(lib, sym), f"{sym} declared in include/precond_b200.h but not exported"
  assert set(declared) == set(_lib.EXPORTED_SYMBOLS)
  assert lib.pc_version() >= 100


def test_workspace_queries_are_host_only():
  lib = _lib.load()
  assert lib.pc_inverse_pth_root_workspace_bytes(0, 128, 0) == 0
  small = lib.pc_inverse_pth_root_workspace_bytes(4, 128, 1)
  big = lib.pc_inverse_pth_root_workspace_bytes(8, 128, 1)
  assert 0 < small < big
  assert lib.pc_graft_momentum_workspace_bytes(1000) > 0
  import ctypes
  opt = _lib.FdOptions()
  lib.pc_fd_options_default(ctypes.byref(opt))
  assert opt.subspace_iters > 0 and opt.full_eigh_max_dim == 512
  exact = lib.pc_fd_update_workspace_bytes(2, 256, 256, 16, ctypes.byref(opt))
  sub = lib.pc_fd_update_workspace_bytes(2, 1024, 1024, 16, ctypes.byref(opt))
  assert 0 < exact < sub
  assert lib.pc_low_rank_to_dense_workspace_bytes(2, 256, 16) > 0


def test_struct_layouts_match_header(tmp_path):
  """ctypes mirrors must have the sizes / offsets the C compiler gives the header."""
  import ctypes
  import subprocess
  src = tmp_path / "sz.c"
  src.write_text(
      '#include <stdio.h>\n#include <stddef.h>\n#include "precond_b200.h"\n'
      'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pc_root_options),'
      'sizeof(pc_gemm_desc), sizeof(pc_graft_options), offsetof(pc_gemm_desc, m),'
      'offsetof(pc_gemm_desc, alpha), offsetof(pc_graft_options, run_shampoo),'
      'sizeof(pc_fd_options), offsetof(pc_fd_options, full_eigh_max_dim),'
      'sizeof(pc_graft_segment), offsetof(pc_graft_segment, has_precond),'
      'sizeof(pc_ipc_handle), sizeof(pc_peer_group), offsetof(pc_peer_group, flags),'
      'sizeof(pc_tearfree_segment), offsetof(pc_tearfree_segment, first_chunk),'
      'sizeof(pc_tearfree_options), offsetof(pc_tearfree_options, scale),'
      'offsetof(pc_gemm_desc, beta_dev), sizeof(pc_quant_segment),'
      'offsetof(pc_quant_segment, first_tile));return 0;}\n')
  exe = tmp_path / "sz"
  subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                 check=True)
  got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True,
                                        check=True).stdout.split()]
  want = [ctypes.sizeof(_lib.RootOptions), ctypes.sizeof(_lib.GemmDesc),
          ctypes.sizeof(_lib.GraftOptions), _lib.GemmDesc.m.offset,
          _lib.GemmDesc.alpha.offset, _lib.GraftOptions.run_shampoo.offset,
          ctypes.sizeof(_lib.FdOptions), _lib.FdOptions.full_eigh_max_dim.offset,
          ctypes.sizeof(_lib.GraftSegment), _lib.GraftSegment.has_precond.offset,
          ctypes.sizeof(_lib.IpcHandle), ctypes.sizeof(_lib.PeerGroup), _lib.PeerGroup.flags.offset,
          ctypes.sizeof(_lib.TearfreeSegment), _lib.TearfreeSegment.first_chunk.offset,
          ctypes.sizeof(_lib.TearfreeOptions), _lib.TearfreeOptions.scale.offset,
          _lib.GemmDesc.beta_dev.offset, ctypes.sizeof(_lib.QuantSegment),
          _lib.QuantSegment.first_tile.offset]
  assert got == want


def test_ffi_adapter_type_checks():
  """ffi/precond_ffi.cc (the jax.ffi handlers over this C ABI) compiles against the header --
  with the real XLA FFI headers when JAX is present, else against the API stub."""
  import subprocess
  ffi_dir = os.path.join(ROOT, "ffi")
  cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-I",
         "/usr/local/cuda/include"]
  try:
    import jax.ffi
    cmd += ["-I", jax.ffi.include_dir()]
  except Exception:  # pylint: disable=broad-except
    cmd += ["-DPC_FFI_SYNTAX_CHECK", "-I", ffi_dir]
  subprocess.run(cmd + [os.path.join(ffi_dir, "precond_ffi.cc")], check=True)
