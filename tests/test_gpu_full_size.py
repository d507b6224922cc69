"""GPU checks at BASELINE.json's FULL sizes (1024 x 1024 statistics, p = 4; Sketchy 4096 x 4096,
rank 256).  A few matrices are compared with the oracle directly; the whole batch goes through
size-independent properties (residual in float64, vmap semantics = batch-permutation invariance,
scaling law, padding invariance, symmetry).  torch float64 is used as a checker only."""
import numpy as np
import pytest
import torch

from oracle import numerics as N

pytestmark = pytest.mark.gpu


def _stats(batch, n, seed):
  import bench
  return bench.make_statistics_torch(batch, n, seed=seed, device=torch.device("cuda", 0))


def _residual64(root, a, p, eps):
  x = root.double()
  d = a.double() + eps * torch.eye(a.shape[0], dtype=torch.float64, device=a.device)
  xp = torch.linalg.matrix_power(x, p)
  return float((xp @ d - torch.eye(a.shape[0], dtype=torch.float64, device=a.device)).abs().max())


def test_roots_1024_batch_properties_and_oracle_subset():
  from precondition_b200 import ops
  b, n, p = 12, 1024, 4
  xs = _stats(b, n, 77)
  roots, m = ops.matrix_inverse_pth_root_batched(xs, [p] * b)
  torch.cuda.synchronize()
  assert torch.isfinite(roots).all() and torch.equal(roots, roots.transpose(1, 2))
  assert float(m[:, 0].max()) <= 1e-6 and float(m[:, 4].max()) == 1.0  # converged, one try
  # direct parity on 8 of the 12 matrices (the oracle needs ~0.5 s per 1024^2 matrix)
  for i in (0, 1, 2, 3, 5, 7, 9, 11):
    a = xs[i].cpu().numpy()
    want, wm = N.matrix_inverse_pth_root(a, p)
    rel = np.linalg.norm(roots[i].cpu().numpy() - want) / np.linalg.norm(want)
    eps = 1e-6 * wm.max_eigen_value
    ours = _residual64(roots[i], xs[i], p, eps)
    ref = _residual64(torch.as_tensor(want).cuda(), xs[i], p, eps)
    assert abs(float(m[i, 1]) - wm.inverse_pth_root_iters) <= 1, (m[i], wm)
    assert abs(float(m[i, 3]) - wm.max_eigen_value) <= 1e-5 * wm.max_eigen_value
    assert ours <= 2 * ref + 1e-6, (i, ours, ref)  # residual no worse than the reference's
    assert rel <= 4 * max(ours, ref) + 1e-3, (i, rel)
  # every matrix: residual bounded like the compared ones
  lam = m[:, 3]
  res = [_residual64(roots[i], xs[i], p, 1e-6 * float(lam[i])) for i in range(b)]
  assert max(res) <= 4 * max(res[0], res[5]) + 1e-6, res
  # vmap semantics: each matrix behaves as if it ran alone -> batch permutation invariance
  perm = torch.randperm(b, generator=torch.Generator().manual_seed(0)).cuda()
  r2, m2 = ops.matrix_inverse_pth_root_batched(xs[perm].contiguous(), [p] * b)
  torch.cuda.synchronize()
  assert torch.equal(r2, roots[perm]) and torch.equal(m2, m[perm])
  # scaling law: (cA)^(-1/p) = c^(-1/p) A^(-1/p); a power-of-two c only shifts exponents
  r3, m3 = ops.matrix_inverse_pth_root_batched((xs[:4] * 16.0).contiguous(), [p] * 4)
  torch.cuda.synchronize()
  assert torch.equal(m3[:, 1], m[:4, 1])  # same iteration counts
  err = float(((r3 * 2.0 - roots[:4]).abs().amax((1, 2)) / roots[:4].abs().amax((1, 2))).max())
  assert err <= 1e-4, err
  # padding invariance (DST:367-398 at full size): embed a 700 x 700 statistic
  sub = xs[0, :700, :700].contiguous()
  padded = torch.eye(n, device=xs.device).repeat(2, 1, 1)
  padded[:, :700, :700] = sub
  rp, mp = ops.matrix_inverse_pth_root_batched(padded.contiguous(), [p, p], [700, 700])
  torch.cuda.synchronize()
  assert float(rp[:, 700:].abs().sum()) == 0 and float(rp[:, :, 700:].abs().sum()) == 0
  want, wm = N.matrix_inverse_pth_root(sub.cpu().numpy(), p)
  eps = 1e-6 * wm.max_eigen_value
  ours = _residual64(rp[0, :700, :700].contiguous(), sub, p, eps)
  ref = _residual64(torch.as_tensor(want).cuda(), sub, p, eps)
  rel = np.linalg.norm(rp[0, :700, :700].cpu().numpy() - want) / np.linalg.norm(want)
  # ill-conditioned (kappa ~ 1e6): two fp32 solvers differ by ~ their residuals
  assert ours <= 2 * ref + 1e-6, (ours, ref)
  assert rel <= 4 * max(ours, ref) + 1e-3, (rel, ours, ref)
  assert abs(float(mp[0, 1]) - wm.inverse_pth_root_iters) <= 1
  assert torch.equal(rp[0], rp[1])


def test_sketchy_4096_rank256_against_float64_eigh():
  """BASELINE config 5's per-GPU share: two chained sketch updates of 4096 x 4096 blocks,
  rank 256, checked against torch.linalg.eigvalsh (float64) of the same covariance."""
  from precondition_b200 import ops
  d, rank, batch = 4096, 256, 2
  dev = torch.device("cuda", 0)
  g = torch.Generator(device=dev).manual_seed(5)
  u = torch.linalg.qr(torch.randn(d, d, generator=g, device=dev))[0]
  spec = torch.cat([torch.logspace(0, -1.5, rank + 64, device=dev),
                    torch.full((d - rank - 64,), 0.01, device=dev)])
  xs = torch.stack([(u * spec) @ torch.randn(d, d, generator=g, device=dev) / d**0.5
                    for _ in range(batch)]).contiguous()
  prev = torch.zeros((batch, d, rank + 2), device=dev)
  for _ in range(2):
    prev, _ = ops.fd_update_root_batched(xs, prev, [4] * batch, rank, decay=0.999)
  out, _ = ops.fd_update_root_batched(xs, prev, [4] * batch, rank, decay=0.999)
  torch.cuda.synchronize()
  for b in range(batch):
    pk = prev[b].double()
    vecs, lam, tail = pk[:, :rank], pk[-rank:, -1], pk[1, -1]
    ridge = 1e-6 * max(float(lam[0]), 1e-6)
    half = vecs * torch.sqrt(0.999 * (lam + ridge))
    c = half @ half.T + xs[b].double() @ xs[b].double().T
    s = torch.linalg.eigvalsh(c).flip(0)
    got = out[b].double()
    gv, ge, gt = got[:, :rank], got[-rank:, -1], got[1, -1]
    assert float(((ge + s[rank]) - s[:rank]).abs().max() / s[0]) <= 2e-5       # eigenvalues
    assert float(abs(gt - (0.999 * tail + s[rank])) / (0.999 * tail + s[rank])) <= 1e-2  # tail
    eye = torch.eye(rank, device=dev, dtype=torch.float64)
    assert float((gv.T @ gv - eye).abs().max()) <= 2e-5                          # orthonormal
    assert float((c @ gv - gv * (ge + s[rank])).norm(dim=0).max() / s[0]) <= 5e-5  # eigenpairs
    assert float(got[-1, -2]) == 0.0
