"""GPU tests of the tcgen05 split-bf16 GEMM engine through its C-ABI test hook."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tc_ok():
  from precondition_b200 import _lib
  return torch.cuda.is_available() and bool(_lib.load().pc_device_supports_tcgen05())


@pytest.mark.parametrize("n,batch,passes,tol", [(128, 1, 6, 2e-6), (256, 3, 6, 2e-6),
                                                (512, 2, 6, 2e-6), (256, 2, 3, 2e-4),
                                                (256, 3, -3, 4e-6), (1024, 1, -3, 4e-6)])
def test_tc_gemm_matches_float64(n, batch, passes, tol):
  """lower(C) = lower(A B^T) with fp32 operands, mirrored into the upper triangle
  (the engine only ever multiplies commuting symmetric matrices and keeps its
  outputs bitwise symmetric): BF16x6 must be fp32-accurate (<= 2e-6 of the
  largest entry), BF16x3 ~2^-16, scaled FP16x3 (passes = -3) ~2^-21."""
  if not _tc_ok():
    pytest.skip("needs sm_100")
  from precondition_b200 import ops
  rng = np.random.default_rng(n)
  a = rng.standard_normal((batch, n, n)).astype(np.float32)
  b = rng.standard_normal((batch, n, n)).astype(np.float32)
  c = ops.debug_tc_gemm(torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda(), passes)
  torch.cuda.synchronize()
  want = np.einsum("bik,bjk->bij", a.astype(np.float64), b.astype(np.float64))
  want = np.tril(want) + np.transpose(np.tril(want, -1), (0, 2, 1))
  got = c.cpu().numpy()
  np.testing.assert_array_equal(got, np.transpose(got, (0, 2, 1)))
  err = np.abs(got - want).max() / np.abs(want).max()
  assert err <= tol, err


def test_tc_gemm_exact_on_bf16_representable_integers():
  """Small integers are exact in bf16 and in fp32 accumulation: bit-exact result."""
  if not _tc_ok():
    pytest.skip("needs sm_100")
  from precondition_b200 import ops
  rng = np.random.default_rng(1)
  a = rng.integers(-4, 5, (2, 256, 256)).astype(np.float32)
  b = rng.integers(-4, 5, (2, 256, 256)).astype(np.float32)
  c = ops.debug_tc_gemm(torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda(), 6)
  torch.cuda.synchronize()
  want = np.einsum("bik,bjk->bij", a, b)
  want = np.tril(want) + np.transpose(np.tril(want, -1), (0, 2, 1))
  np.testing.assert_array_equal(c.cpu().numpy(), want)


@pytest.mark.parametrize("env", [{"PC_TC_PAIR256": "1"}, {"PC_TC_WS2": "1"}, {"PC_TC_CHUNK": "2"}])
def test_opt_in_kernel_variants_stay_correct(env, monkeypatch):
  """The opt-in tcgen05 kernels (CTA-pair 256 x 256, CTA-pair 256 x 128 with output stages,
  128-column K-chunks) must give the same products and roots as the default 1-CTA kernel."""
  if not _tc_ok():
    pytest.skip("needs sm_100")
  from precondition_b200 import ops
  from oracle import numerics as N
  from oracle.gen_golden import ema_statistics, gen_symmetric_matrix
  for k, v in env.items():
    monkeypatch.setenv(k, v)  # read by the engine at every call
  rng = np.random.default_rng(2)
  a = rng.standard_normal((2, 512, 512)).astype(np.float32)
  b = rng.standard_normal((2, 512, 512)).astype(np.float32)
  c = ops.debug_tc_gemm(torch.as_tensor(a).cuda(), torch.as_tensor(b).cuda(), -3)
  torch.cuda.synchronize()
  want = np.einsum("bik,bjk->bij", a.astype(np.float64), b.astype(np.float64))
  want = np.tril(want) + np.transpose(np.tril(want, -1), (0, 2, 1))
  got = c.cpu().numpy()
  np.testing.assert_array_equal(got, np.transpose(got, (0, 2, 1)))
  assert np.abs(got - want).max() / np.abs(want).max() <= 4e-6
  xs = np.stack([gen_symmetric_matrix(rng, 256, 1e3), ema_statistics(rng, 256, 512)])
  xs = xs.astype(np.float32)
  roots, metrics = ops.matrix_inverse_pth_root_batched(torch.as_tensor(xs).cuda(), [4, 2],
                                                       engine=4)
  torch.cuda.synchronize()
  for i, p in enumerate([4, 2]):
    w, wm = N.matrix_inverse_pth_root(xs[i], p)
    rel = np.linalg.norm(roots[i].cpu().numpy() - w) / np.linalg.norm(w)
    assert rel <= 1e-4, (env, i, rel)
    assert abs(float(metrics[i, 1]) - wm.inverse_pth_root_iters) <= 1
