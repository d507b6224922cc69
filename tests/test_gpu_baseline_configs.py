"""GPU parity at the shapes BASELINE.json names (its configs are parity-test cases, not
bench lines), each against the CPU oracle on the same seeded inputs:

  C1  matrix_inverse_pth_root p=4 on a batch of 64 random SPD 128x128 statistics
  C3  ResNet-50-shaped parameters, block_size=1024, preconditioning_compute_steps=1
  C4  transformer-shaped parameter, block_size=2048, int16-quantised statistics
      (best_effort_memory_usage_reduction with a batch axis, DS:2051-2064)

C2 (MLP 512->2048->512, block 128) is covered in miniature by
test_gpu_optimizer.test_blocked_mlp_matches_oracle and at full size by bench.py; C5 (Sketchy)
by test_gpu_fd.py / test_sketchy_full_rank_blocks_match_oracle at sizes the oracle's SVD
finishes in seconds.
"""
import numpy as np
import pytest
import torch

from oracle import numerics as N
from oracle import optimizer as O
from oracle.gen_golden import ema_statistics, gen_symmetric_matrix

pytestmark = pytest.mark.gpu


def test_c1_batch_of_64_spd_128():
  from precondition_b200 import ops
  rng = np.random.default_rng(0)
  xs = np.stack([gen_symmetric_matrix(rng, 128, 1e4) if i % 2 == 0 else
                 ema_statistics(rng, 128, 512) for i in range(64)]).astype(np.float32)
  roots, metrics = ops.matrix_inverse_pth_root_batched(torch.as_tensor(xs).cuda(), [4] * 64)
  torch.cuda.synchronize()
  roots, metrics = roots.cpu().numpy(), metrics.cpu().numpy()
  for b in range(64):
    trace = []
    want, wm = N.matrix_inverse_pth_root(xs[b], 4, trace=trace)
    rel = np.linalg.norm(roots[b] - want) / np.linalg.norm(want)
    assert rel <= 1e-3, (b, rel)  # north_star tolerance
    knife = trace[-1][2] >= 2.5e-7 or (len(trace) > 1 and trace[-2][2] <= 4e-6)
    assert abs(metrics[b, 1] - wm.inverse_pth_root_iters) <= (1 if knife else 0), (b, metrics[b])
    assert metrics[b, 4] == wm.total_retries
    eps = 1e-6 * wm.max_eigen_value
    assert N.root_residual(roots[b], xs[b], 4, eps) <= 2 * N.root_residual(want, xs[b], 4, eps) + 1e-6


def _run_both(shapes, block, steps, gscale, seed, tol, **kw):
  from precondition_b200 import distributed_shampoo as DS
  rng = np.random.default_rng(seed)
  params = [rng.standard_normal(s).astype(np.float32) * 0.05 for s in shapes]
  oracle = O.distributed_shampoo(0.1, block, start_preconditioning_step=1, **kw)
  ostate = oracle.init(params)
  opt = DS.distributed_shampoo(0.1, block, start_preconditioning_step=1, **kw)
  tparams = [torch.as_tensor(p).cuda() for p in params]
  state = opt.init(tparams)
  for t in range(steps):
    grads = [(rng.standard_normal(s) * gscale).astype(np.float32) for s in shapes]
    want, ostate = oracle.update(grads, ostate, params)
    got, state = opt.update([torch.as_tensor(g).cuda() for g in grads], state, tparams)
    torch.cuda.synchronize()
    for i, (u, w) in enumerate(zip(got, want)):
      err = np.abs(u.cpu().numpy() - w).max() / max(np.abs(w).max(), 1e-12)
      assert err <= tol, (t, i, err)
  return state, ostate


def test_c3_resnet50_shapes_block_1024():
  """1x1 conv 1024->1024 (one 1024 x 1024 block, tcgen05 engine), a 3x3 conv that merges to
  rank 3 (p = 6) and a BN vector: 2 steps, second one preconditioned."""
  shapes = [(1, 1, 1024, 1024), (3, 3, 64, 64), (1024,)]
  state, ostate = _run_both(shapes, 1024, 2, 1e-2, 1, 1e-3, preconditioning_compute_steps=1)
  tm = state.stats[0].training_metrics.cpu().numpy()
  otm = ostate.stats[0].training_metrics
  assert tm.shape == (2, 5) and np.all(np.abs(tm[:, 1] - otm[:, 1]) <= 1), (tm, otm)
  assert np.all(tm[:, 0] < 1e-5)
  for a, b in zip(state.stats[0].preconditioners, ostate.stats[0].preconditioners):
    rel = np.linalg.norm(a.cpu().numpy() - b) / np.linalg.norm(b)
    assert rel <= 1e-3, rel


def test_c4_transformer_shapes_block_2048_int16():
  """FFN-shaped 1024 x 2048 parameter, block_size 2048: statistics of 1024 and 2048, stored as
  int16 + diagonal + bucket sizes; int8 momenta.  One int8 quantum (1/127) on the updates."""
  state, ostate = _run_both([(1024, 2048)], 2048, 2, 1e-2, 2, 1.2e-2,
                            best_effort_memory_usage_reduction=True, batch_axis_name="batch")
  for a, b in zip(state.stats[0].preconditioners, ostate.stats[0].preconditioners):
    fa, fb = a.to_float().cpu().numpy(), b.to_float()
    assert fa.shape == fb.shape
    rel = np.linalg.norm(fa - fb) / np.linalg.norm(fb)
    assert rel <= 5e-3, rel
