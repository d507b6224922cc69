"""GPU parity tests of the Sketchy / frequent-directions sketch update (C ABI
``pc_fd_update_batched``) against the CPU oracle (``oracle.numerics.fd_update_root``,
DS:1123-1290) and the golden vectors recorded from the reference.

Tolerances follow the reference's own tests (DST:770-885): eigenvector alignment and
eigenvalues to 1e-2; the exact small path (d <= 512, cyclic Jacobi on the whole
covariance) is held to 2e-4.  Singular vectors are only defined up to sign (and up to a
rotation inside a degenerate cluster), so sketches are compared through sign-invariant
quantities: |V^T V_ref|, the eigenvalue slots, and the low-rank preconditioner the
packed sketch represents (DS:1690-1705).
"""
import numpy as np
import pytest
import torch

from oracle import numerics as N

pytestmark = pytest.mark.gpu


def _run(fac, prev, p, rank, pad=None, **kw):
  from precondition_b200 import ops
  f = torch.as_tensor(np.ascontiguousarray(fac, dtype=np.float32)).cuda()
  pv = torch.as_tensor(np.ascontiguousarray(prev, dtype=np.float32)).cuda()
  b = f.shape[0]
  out, metrics = ops.fd_update_root_batched(
      f, pv, [p] * b if np.isscalar(p) else p, rank,
      None if pad is None else ([pad] * b if np.isscalar(pad) else pad), **kw)
  torch.cuda.synchronize()
  return out.cpu().numpy(), metrics.cpu().numpy()


def _operator(packed, rank):
  """Dense matrix of the low-rank preconditioner a packed sketch encodes, DS:1690-1705:
  g -> c (g - g V V^T) + (g V lambda^-) V^T."""
  vecs, _, inv, const, _, _ = N.fd_low_rank_unpack(packed.astype(np.float64), rank)
  d = packed.shape[0]
  return const * (np.eye(d) - vecs @ vecs.T) + (vecs * inv) @ vecs.T


def _compare(got, want, rank, tol, tag):
  gv, ge, gi, gc, gt, gz = N.fd_low_rank_unpack(got, rank)
  wv, we, wi, wc, wt, wz = N.fd_low_rank_unpack(want, rank)
  scale = max(float(np.abs(we).max()), float(wt), 1e-30)
  np.testing.assert_allclose(ge, we, rtol=tol, atol=tol * scale, err_msg=f"{tag}: deflated eigs")
  np.testing.assert_allclose(gt, wt, rtol=tol, atol=tol * scale, err_msg=f"{tag}: tail")
  np.testing.assert_allclose(gc, wc, rtol=tol, atol=1e-30, err_msg=f"{tag}: const")
  np.testing.assert_allclose(gi, wi, rtol=tol, atol=tol * float(np.abs(wi).max() + 1e-30),
                             err_msg=f"{tag}: inverted eigs")
  assert gz == wz, f"{tag}: has_zeros"
  # kept directions: same count, unit norm, aligned with the reference's up to sign
  kept_g, kept_w = np.linalg.norm(gv, axis=0) > 0, np.linalg.norm(wv, axis=0) > 0
  np.testing.assert_array_equal(kept_g, kept_w, err_msg=f"{tag}: kept directions")
  np.testing.assert_allclose(np.linalg.norm(gv[:, kept_g], axis=0), 1.0, atol=1e-4)
  og, ow = _operator(got, rank), _operator(want, rank)
  err = np.abs(og - ow).max() / max(np.abs(ow).max(), 1e-30)
  assert err <= tol, f"{tag}: preconditioner operator differs by {err}"


def test_fd_golden_sequences(golden_fd):
  """The reference's own outputs (tests/golden/fd.npz, generated from the unmodified
  reference by oracle/gen_golden.py): 4 chained steps, then 3 steps with padding."""
  g = golden_fd
  for step in range(4):
    out, m = _run(g[f"fd/{step}/factor"][None], g[f"fd/{step}/prev"][None], 4, 4, 24,
                  decay=0.9)
    _compare(out[0], g[f"fd/{step}/new"], 4, 2e-4, f"fd/{step}")
    assert m[0, 0] == 0.0
  for step in range(3):
    out, _ = _run(g[f"fdpad/{step}/factor"][None], g[f"fdpad/{step}/prev"][None], 2, 4, 17,
                  decay=1.0)
    _compare(out[0], g[f"fdpad/{step}/new"], 4, 2e-4, f"fdpad/{step}")
    assert np.abs(out[0][17:, :4]).sum() == 0.0  # no mass on padding rows


def test_fd_gradient_block_equals_qr_factor(golden_fd):
  """Only F F^T enters the update: the raw gradient unfolding [d, m] (m != d) and the
  covariance x x^T give the same sketch as the reference's QR factor."""
  g = golden_fd
  for step in (0, 1):
    x, prev, want = g[f"fd/{step}/g"], g[f"fd/{step}/prev"], g[f"fd/{step}/new"]
    out, _ = _run(x[None], prev[None], 4, 4, 24, decay=0.9)
    _compare(out[0], want, 4, 2e-4, f"block/{step}")
    gram = (x.astype(np.float64) @ x.astype(np.float64).T).astype(np.float32)
    out, _ = _run(gram[None], prev[None], 4, 4, 24, decay=0.9, input_is_gram=True)
    _compare(out[0], want, 4, 2e-4, f"gram/{step}")


@pytest.mark.parametrize("p", [2, 4, 6, 8])
def test_fd_dynamic_exponent(p):
  """DST:706-721."""
  rank, size = 1, 4
  prev = N.fd_low_rank_pack(np.eye(size, rank, dtype=np.float32), np.zeros(1, np.float32),
                            np.zeros(1, np.float32), 0.0, 0.0, False, rank)
  grad = np.zeros((size, size), np.float32)
  grad[0, 0] = 2**(p / 2)
  out, _ = _run(grad[None], prev[None], p, rank)
  assert abs(N.fd_low_rank_unpack(out[0], rank)[2][0] - 0.5) <= 1e-6
  prev = N.fd_low_rank_pack(np.eye(size, rank, dtype=np.float32),
                            np.array([2.0**p], np.float32), np.zeros(1, np.float32), 0.0, 0.0,
                            False, rank)
  out, _ = _run(np.zeros((1, size, size), np.float32), prev[None], p, rank)
  assert abs(N.fd_low_rank_unpack(out[0], rank)[2][0] - 0.5) <= 1e-6


@pytest.mark.parametrize("relative", [True, False])
def test_fd_nonzero_epsilon(relative):
  """DST:723-745."""
  rank, size, p = 1, 4, 2
  eig0 = 2.0**p
  prev = N.fd_low_rank_pack(np.eye(size, rank, dtype=np.float32), np.array([eig0], np.float32),
                            np.zeros(1, np.float32), 0.0, 0.0, False, rank)
  out, _ = _run(np.zeros((1, size, size), np.float32), prev[None], p, rank, ridge_epsilon=0.1,
                relative_matrix_epsilon=relative)
  vecs, eigs, _, _, tail, _ = N.fd_low_rank_unpack(out[0], rank)
  eps = 0.1 * (eig0 if relative else 1.0)
  assert abs(eigs[0] - (eig0 + eps)) <= 1e-5
  assert np.abs(np.abs(vecs[:, 0]) - np.array([1, 0, 0, 0])).max() <= 10 * np.finfo(np.float32).eps
  assert abs(tail) <= 1e-6


@pytest.mark.parametrize("size,padding,rank", [(5, 0, 2), (5, 3, 2), (40, 0, 8), (96, 32, 16),
                                               (300, 0, 32)])
def test_fd_matches_oracle_and_float64_eigh(size, padding, rank):
  """DST:770-827 (test_basic) at several sizes, batched, against the oracle and float64 eigh."""
  rng = np.random.default_rng(size + rank)
  d = size + padding
  batch = 3
  facs, prevs, wants = [], [], []
  for b in range(batch):
    grad = rng.standard_normal((size, size)) * np.logspace(0, -2, size)[None, :]
    cov = grad @ grad.T
    top = np.linalg.eigvalsh(cov).max()
    eigs = np.sort(rng.uniform(top / 8, top * 4, rank))[::-1].astype(np.float32)
    q, _ = np.linalg.qr(rng.standard_normal((size, rank)))
    vec = np.zeros((d, rank), np.float32)
    vec[:size] = q
    prev = N.fd_low_rank_pack(vec, eigs, np.zeros(rank, np.float32), 0.0, 0.25 * b, False, rank)
    fac = np.zeros((d, d), np.float32)
    fac[:size, :size] = grad
    want, _ = N.fd_update_root(fac, 4, rank, decay=0.95, padding_start=size, prev=prev)
    facs.append(fac); prevs.append(prev); wants.append(want)
  out, _ = _run(np.stack(facs), np.stack(prevs), 4, rank, size, decay=0.95)
  for b in range(batch):
    _compare(out[b], wants[b], rank, 2e-3, f"size{size}/b{b}")
    assert np.abs(out[b][size:, :rank]).sum() == 0.0
    # float64 ground truth of the covariance spectrum (DST:806-827)
    vecs, eigs, _, _, tail, _ = N.fd_low_rank_unpack(out[b], rank)
    pv, pe, _, _, ptail, _ = N.fd_low_rank_unpack(prevs[b].astype(np.float64), rank)
    ridge = 1e-6 * max(pe[0], 1e-6)
    half = pv[:size] * np.sqrt(0.95 * (pe + ridge))
    g64 = facs[b][:size, :size].astype(np.float64)
    s = np.linalg.eigvalsh(half @ half.T + g64 @ g64.T)[::-1]
    np.testing.assert_allclose(tail, 0.95 * ptail + s[rank], rtol=1e-2)
    np.testing.assert_allclose(eigs + s[rank], s[:rank], rtol=1e-2)


def test_fd_subspace_path_matches_oracle():
  """d above full_eigh_max_dim: block subspace iteration + Rayleigh-Ritz.  Forced at a
  size the oracle's SVD finishes quickly (d = 384 with full_eigh_max_dim = 128)."""
  rng = np.random.default_rng(11)
  d, rank, batch = 384, 24, 2
  facs, prevs, wants = [], [], []
  for b in range(batch):
    spectrum = np.concatenate([np.logspace(0, -1, rank + 8), np.full(d - rank - 8, 0.02)])
    u, _ = np.linalg.qr(rng.standard_normal((d, d)))
    fac = ((u * spectrum) @ rng.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
    q, _ = np.linalg.qr(rng.standard_normal((d, rank)))
    eigs = np.sort(rng.uniform(0.2, 2.0, rank))[::-1].astype(np.float32)
    prev = N.fd_low_rank_pack(q.astype(np.float32), eigs, np.zeros(rank, np.float32), 0.0, 0.1,
                              False, rank)
    want, _ = N.fd_update_root(fac, 4, rank, decay=0.999, padding_start=d, prev=prev)
    facs.append(fac); prevs.append(prev); wants.append(want)
  out, _ = _run(np.stack(facs), np.stack(prevs), 4, rank, d, decay=0.999, full_eigh_max_dim=128,
                subspace_iters=12)
  for b in range(batch):
    _compare(out[b], wants[b], rank, 1e-2, f"subspace/b{b}")


def test_fd_all_padding_and_bad_arguments():
  from precondition_b200 import ops
  prev = np.zeros((1, 8, 4), np.float32)
  out, m = _run(np.ones((1, 8, 8), np.float32), prev, 4, 2, 0)
  assert np.abs(out).sum() == 0.0 and m[0, 0] == 0.0
  with pytest.raises(RuntimeError):  # rank + 2 must be < d (DS:535-537)
    ops.fd_update_root_batched(torch.zeros((1, 4, 4)).cuda(), torch.zeros((1, 4, 4)).cuda(), [4], 2)
  with pytest.raises(RuntimeError):  # CPU tensors: no fallback
    ops.fd_update_root_batched(torch.zeros((1, 8, 8)), torch.zeros((1, 8, 4)), [4], 2)


def _lr_operator(packed, rank):
  vecs, inv, const, skip = N.low_rank_unpack(packed.astype(np.float64), abs(rank))
  d = packed.shape[0]
  return const * (np.eye(d) - vecs @ vecs.T) + (vecs * inv) @ vecs.T


def test_low_rank_root_matches_reference_golden(golden_fd):
  """eigh-based _low_rank_root (DS:1033-1120) on the reference's own outputs: positive rank
  keeps the largest eigenvalues, negative the smallest; with and without padding.
  Eigenvectors are defined up to sign, so the packed operator and the scalar slots are
  compared."""
  from precondition_b200 import ops
  g = golden_fd
  a = torch.as_tensor(g["lowrank/a"]).cuda()[None].contiguous()
  for cr in (3, -3):
    for pad in (20, 15):
      out, m = ops.low_rank_root_batched(a, [4], cr, [pad])
      torch.cuda.synchronize()
      got, want = out[0].cpu().numpy(), g[f"lowrank/{cr}/{pad}/root"]
      k = abs(cr)
      np.testing.assert_allclose(got[:, k:], want[:, k:], rtol=2e-4, atol=1e-6,
                                 err_msg=f"slots rank {cr} pad {pad}")
      og, ow = _lr_operator(got, cr), _lr_operator(want, cr)
      assert np.abs(og - ow).max() <= 5e-4 * np.abs(ow).max(), (cr, pad)
      assert np.abs(got[pad:, :k]).sum() == 0.0
      err, werr = float(m[0, 0]), float(g[f"lowrank/{cr}/{pad}/err"])
      assert err <= max(10 * werr, 1e-5), (err, werr)


@pytest.mark.parametrize("d,rank", [(64, 8), (200, -16), (512, 32)])
def test_low_rank_root_matches_oracle_batched(d, rank):
  rng = np.random.default_rng(d)
  from oracle.gen_golden import gen_symmetric_matrix
  from precondition_b200 import ops
  mats = np.stack([gen_symmetric_matrix(rng, d, 10.0**(2 + b)) * (b + 1) for b in range(3)])
  mats = mats.astype(np.float32)
  ps = [4, 2, 4]
  out, m = ops.low_rank_root_batched(torch.as_tensor(mats).cuda(), ps, rank)
  torch.cuda.synchronize()
  for b in range(3):
    want, wm = N.low_rank_root(mats[b], ps[b], rank)
    got = out[b].cpu().numpy()
    k = abs(rank)
    np.testing.assert_allclose(got[:k, k], want[:k, k], rtol=2e-3)       # inverted eigenvalues
    np.testing.assert_allclose(got[0, k + 1], want[0, k + 1], rtol=2e-3)  # const
    og, ow = _lr_operator(got, rank), _lr_operator(want, rank)
    assert np.abs(og - ow).max() <= 2e-3 * np.abs(ow).max(), (b, d, rank)
    assert float(m[b, 0]) < 1e-3


def test_fd_1024_against_the_oracle_svd():
  """d = 1024, rank 64 (subspace path, tcgen05 products): the sketch eigenvalues, tail and
  vectors against the ORACLE's SVD-based _fd_update_root (numpy gesdd of [1024, 1088]) over three
  chained updates -- eigenvalues to 1e-4 of the largest, tail to 1e-3, subspace to 1e-3."""
  from precondition_b200 import ops
  rng = np.random.default_rng(21)
  d, rank = 1024, 64
  u, _ = np.linalg.qr(rng.standard_normal((d, d)))
  spec = np.concatenate([np.logspace(0, -1.2, rank + 32), np.full(d - rank - 32, 0.03)])
  prev_o = np.zeros((d, rank + 2), np.float32)
  prev_g = torch.zeros((1, d, rank + 2), device="cuda")
  for step in range(3):
    fac = ((u * spec) @ rng.standard_normal((d, d)) / np.sqrt(d)).astype(np.float32)
    want, _ = N.fd_update_root(fac, 4, rank, decay=0.999, padding_start=d, prev=prev_o)
    got, _ = ops.fd_update_root_batched(torch.as_tensor(fac[None]).cuda(), prev_g, [4], rank,
                                        decay=0.999)
    torch.cuda.synchronize()
    g = got[0].cpu().numpy().astype(np.float64)
    w = want.astype(np.float64)
    top = w[-rank:, -1].max() + w[1, -1]
    assert np.abs(g[-rank:, -1] - w[-rank:, -1]).max() <= 1e-4 * top, step      # deflated eigenvalues
    assert abs(g[1, -1] - w[1, -1]) <= 1e-3 * w[1, -1], step                     # tail
    assert abs(g[0, -1] - w[0, -1]) <= 1e-3 * abs(w[0, -1]), step                # const = tail^(-1/p)
    assert g[-1, -2] == w[-1, -2] == 0.0                                         # has_zeros
    # same subspace: projector difference of the leading 48 vectors (well separated eigenvalues)
    pg, pw = g[:, :48] @ g[:, :48].T, w[:, :48] @ w[:, :48].T
    assert np.abs(pg - pw).max() <= 1e-3, step
    prev_o, prev_g = want, got
