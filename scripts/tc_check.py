"""GPU diagnostic (not a test): exercises the tcgen05 engine bottom-up and prints
errors instead of asserting, so one gpurun call localises a fault."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from precondition_b200 import ops  # noqa: E402
from oracle import numerics as N  # noqa: E402
from oracle.gen_golden import ema_statistics, gen_symmetric_matrix  # noqa: E402


def gemm_case(n, batch, passes, kind):
  rng = np.random.default_rng(n + batch)
  if kind == "identity":
    a = np.stack([np.eye(n)] * batch)
    b = rng.standard_normal((batch, n, n))
  elif kind == "rowid":
    a = rng.standard_normal((batch, n, n))
    b = np.stack([np.eye(n)] * batch)
  else:
    a = rng.standard_normal((batch, n, n))
    b = rng.standard_normal((batch, n, n))
  a32, b32 = a.astype(np.float32), b.astype(np.float32)
  c = ops.debug_tc_gemm(torch.as_tensor(a32).cuda(), torch.as_tensor(b32).cuda(), passes)
  torch.cuda.synchronize()
  c = c.cpu().numpy()
  want = np.einsum("bik,bjk->bij", a32.astype(np.float64), b32.astype(np.float64))
  want = np.tril(want) + np.transpose(np.tril(want, -1), (0, 2, 1))  # engine mirrors lower tiles
  err = np.abs(c - want).max() / np.abs(want).max()
  bad = np.argwhere(np.abs(c - want) > 1e-3 * np.abs(want).max())
  print(f"gemm n={n} batch={batch} passes={passes} {kind}: max rel err {err:.3e}; "
        f"#bad={len(bad)} first_bad={bad[:3].tolist()}", flush=True)
  if len(bad):
    b0, i0, j0 = bad[0]
    print("   got", c[b0, i0, j0:j0 + 4], "want", want[b0, i0, j0:j0 + 4], flush=True)


def root_case(n, batch, engine):
  rng = np.random.default_rng(5)
  xs = np.stack([ema_statistics(rng, n, n // 4) if i % 2 else gen_symmetric_matrix(rng, n, 1e3)
                 for i in range(batch)]).astype(np.float32)
  t0 = time.time()
  r, m = ops.matrix_inverse_pth_root_batched(torch.as_tensor(xs).cuda(), [4] * batch,
                                             engine=engine)
  torch.cuda.synchronize()
  dt = time.time() - t0
  r, m = r.cpu().numpy(), m.cpu().numpy()
  for b in range(min(batch, 2)):
    want, wm = N.matrix_inverse_pth_root(xs[b], 4)
    eps = 1e-6 * wm.max_eigen_value
    rel = np.linalg.norm(r[b] - want) / np.linalg.norm(want)
    print(f"root n={n} engine={engine} b={b}: rel={rel:.3e} iters={m[b,1]} (ref {wm.inverse_pth_root_iters}) "
          f"err={m[b,0]:.3e} (ref {wm.inverse_pth_root_errors:.3e}) resid={N.root_residual(r[b], xs[b], 4, eps):.3e} "
          f"(ref {N.root_residual(want, xs[b], 4, eps):.3e}) t={dt:.3f}s", flush=True)


if __name__ == "__main__":
  for kind in ("identity", "rowid", "random"):
    gemm_case(128, 1, 6, kind)
  gemm_case(256, 3, 6, "random")
  gemm_case(256, 2, 3, "random")
  gemm_case(1024, 2, 6, "random")
  for kind in ("identity", "rowid", "random"):
    gemm_case(128, 1, -3, kind)
  gemm_case(256, 3, -3, "random")
  gemm_case(1024, 2, -3, "random")
  for n in (128, 256, 1024):
    root_case(n, 4, 2)
    root_case(n, 4, 4)
  root_case(256, 4, 1)
