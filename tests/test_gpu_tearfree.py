"""GPU tests of the tearfree front-end (precondition_b200/tearfree) against the golden trajectories
recorded from the unmodified reference (tests/golden/tearfree.npz) and against the numpy oracle
(oracle/tearfree.py): the fused optimizer, its stand-alone parts, the pseudo-inverse root and the
grouped tail kernel."""
import os

import numpy as np
import pytest
import torch

from oracle import gen_golden as G
from oracle import tearfree as T

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tearfree.npz")


def _options(kw):
  from precondition_b200.tearfree import grafting, momentum, optimizer, second_order, shampoo
  from precondition_b200.tearfree import sketchy
  g = lambda k, d: kw.get(k, d)
  gtype = {"none": grafting.GraftingType.NONE, "sgd": grafting.GraftingType.SGD,
           "rmsprop": grafting.GraftingType.RMSPROP}
  return optimizer.TearfreeOptions(
      grafting_options=grafting.Options(
          grafting_type=gtype[g("graft", "rmsprop")], second_moment_decay=g("graft_decay", 0.999),
          epsilon=g("graft_epsilon", 1e-23),
          start_preconditioning_step=g("start_preconditioning_step", 0),
          skip_preconditioning_any_dim_gt=g("skip_preconditioning_any_dim_gt", 4096),
          skip_preconditioning_rank1=g("skip_preconditioning_rank1", True)),
      second_order_options=second_order.Options(
          merge_dims=g("merge_dims", 1024),
          second_order_type=(second_order.SecondOrderType.SKETCHY
                             if g("second_order", "") == "sketchy"
                             else second_order.SecondOrderType.SHAMPOO),
          sketchy_options=sketchy.Options(
              epsilon=g("sketchy_epsilon", 1e-7), rank=g("sketchy_rank", 128),
              relative_epsilon=g("sketchy_relative_epsilon", True),
              second_moment_decay=g("sketchy_decay", 0.999),
              update_freq=g("sketchy_update_freq", 1)),
          shampoo_options=shampoo.Options(
              block_size=g("block_size", 1024),
              update_preconditioners_freq=g("update_preconditioners_freq", 1),
              update_statistics_freq=g("update_statistics_freq", 1),
              second_moment_decay=g("second_moment_decay", 0.999))),
      momentum_options=momentum.Options(
          ema=g("ema", False), nesterov=g("nesterov", True),
          momentum_decay=g("momentum_decay", 0.9), weight_decay=g("weight_decay", 0.0),
          weight_decay_after_momentum=g("weight_decay_after_momentum", True)))


def _rel(a, b):
  return float(np.abs(np.asarray(a, np.float64) - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("tag", list(G.TEARFREE_CASES))
def test_tearfree_matches_reference_golden(tag):
  """optimizer.tearfree on CUDA follows the unmodified reference step for step.  Tolerance 2e-3 of
  the largest update entry: the roots come from a Jacobi eigensolver instead of LAPACK, and a
  2*rank-th root of an fp32 eigenvalue near the 1e-6 cutoff moves by that much in either."""
  from precondition_b200 import ops
  from precondition_b200.tearfree import optimizer
  g = np.load(GOLDEN)
  params, grads, kw = G.tearfree_inputs(tag)
  lr = G.tearfree_schedule if kw["learning_rate"] == "schedule" else kw["learning_rate"]
  tx = optimizer.tearfree(lr, _options(kw))
  dparams = [torch.as_tensor(p).cuda() for p in params]
  state = tx.init(dparams)
  before = ops.gpu_launches
  for t, gr in enumerate(grads):
    u, state = tx.update([torch.as_tensor(x).cuda() for x in gr], state, dparams)
    for i, ui in enumerate(u):
      assert ui.shape == dparams[i].shape
      r = _rel(ui.cpu().numpy(), g[f"{tag}/update{t}_{i}"])
      assert r <= 2e-3, (tag, t, i, r)
  assert ops.gpu_launches > before
  graft_state = state[0]
  direction = graft_state if kw.get("graft") == "none" else graft_state.direction
  if kw.get("second_order") == "sketchy":
    for i, t in enumerate(direction[1].sketches):
      for a, ax in enumerate(getattr(t, "axes", [])):
        for name in ("eigvals", "inv_eigvals", "tail", "inv_tail"):
          want = g[f"{tag}/{name}{i}_{a}"]
          got = getattr(ax, name).cpu().numpy()
          assert got.shape == want.shape, (name, got.shape, want.shape)
          assert np.abs(got - want).max() <= 1e-3 * (np.abs(want).max() + 1e-30), (tag, i, a, name)
        v = ax.eigvecs.cpu().numpy()
        assert np.abs(v @ v.T - g[f"{tag}/projector{i}_{a}"]).max() <= 2e-3, (tag, i, a)
    return
  blocks = direction[1].blocks
  for i, b in enumerate(blocks):
    if not hasattr(b, "stats"):
      continue
    for a, (st, rt) in enumerate(zip(b.stats, b.roots)):
      assert _rel(st.cpu().numpy(), g[f"{tag}/stats{i}_{a}"]) <= 1e-5, (tag, i, a)
      assert _rel(rt.cpu().numpy(), g[f"{tag}/roots{i}_{a}"]) <= 2e-3, (tag, i, a)


def test_tearfree_rank3_and_two_blocked_axes_against_the_oracle():
  """A rank-3 tensor (mode products along a middle axis), two blocked axes with padding and a
  skipped vector, 6 steps with statistics every step and roots every second step."""
  from precondition_b200.tearfree import optimizer
  rng = np.random.default_rng(5)
  shapes = [(12, 5, 6), (20, 9), (40,), (3, 3, 8, 5)]
  kw = dict(learning_rate=0.02, graft="rmsprop", merge_dims=9, block_size=8,
            update_preconditioners_freq=2, second_moment_decay=0.95, momentum_decay=0.8)
  params = [rng.standard_normal(s).astype(np.float32) for s in shapes]
  ref = T.Tearfree(params, **kw)
  tx = optimizer.tearfree(kw["learning_rate"], _options(kw))
  dparams = [torch.as_tensor(p).cuda() for p in params]
  state = tx.init(dparams)
  for t in range(6):
    gr = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    want = ref.update(gr, params)
    got, state = tx.update([torch.as_tensor(x).cuda() for x in gr], state, dparams)
    for i in range(len(shapes)):
      assert _rel(got[i].cpu().numpy(), want[i]) <= 2e-3, (t, i)


@pytest.mark.parametrize("d,rank", [(16, 16), (128, 40), (256, 256), (1024, 300)])
def test_pinv_root_against_the_oracle(d, rank):
  """pc_pinv_pth_root_eigh_batched == oracle pth_inv_root (TF/shampoo.py:440-448) on full-rank
  and rank-deficient Gram matrices (the null space must come out as exact zeros of the root, not
  as a huge inverse), p = 4 and 6."""
  from precondition_b200 import ops
  rng = np.random.default_rng(d + rank)
  xs = []
  for _ in range(2):
    f = rng.standard_normal((d, rank)).astype(np.float32) * np.logspace(0, -1.5, rank).astype(
        np.float32)
    xs.append(f @ f.T)
  xs = np.stack(xs).astype(np.float32)
  ps = [4, 6]
  got = ops.pinv_pth_root_eigh_batched(torch.as_tensor(xs).cuda(), ps).cpu().numpy()
  for b in range(2):
    want = T.pth_inv_root(ps[b], xs[b])
    assert _rel(got[b], want) <= 5e-4, (b, _rel(got[b], want))
    assert np.abs(got[b] - got[b].T).max() <= 1e-5 * np.abs(want).max()
  zero = ops.pinv_pth_root_eigh_batched(torch.zeros(1, d, d, device="cuda"), [4]).cpu().numpy()
  assert np.all(zero == 0.0)  # w <= eps * max(w) masks everything (TF/shampoo.py:444)


def test_stand_alone_parts_equal_the_fused_optimizer():
  """grafting.graft(second_order) -> momentum.apply -> -lr, one transformation at a time (as the
  reference chains them, TF/optimizer.py:85-99), gives the fused optimizer's updates bit for bit
  except for the order of the last scaling (tolerance 1 ulp)."""
  from precondition_b200.tearfree import grafting, momentum, optimizer, second_order
  kw = dict(graft="rmsprop", merge_dims=8, block_size=8, second_moment_decay=0.9, ema=True,
            momentum_decay=0.7, weight_decay=0.05, start_preconditioning_step=2)
  opts = _options(kw)
  rng = np.random.default_rng(11)
  shapes = [(16, 8), (5,), (6, 4)]
  params = [torch.as_tensor(rng.standard_normal(s).astype(np.float32)).cuda() for s in shapes]
  fused = optimizer.tearfree(0.3, opts)
  graft_tx = grafting.graft(opts.grafting_options, second_order.apply(opts.second_order_options))
  mom_tx = momentum.apply(opts.momentum_options)
  s_fused, s_graft, s_mom = fused.init(params), graft_tx.init(params), mom_tx.init(params)
  for t in range(4):
    gr = [torch.as_tensor(rng.standard_normal(s).astype(np.float32)).cuda() for s in shapes]
    a, s_fused = fused.update(gr, s_fused, params)
    u, s_graft = graft_tx.update(gr, s_graft, params)
    u, s_mom = mom_tx.update(u, s_mom, params)
    for x, y in zip(a, u):
      torch.testing.assert_close(x, -0.3 * y, rtol=2e-7, atol=0)


def test_momentum_and_grafting_options_against_formulas():
  """momentum.apply alone over every (ema, nesterov, weight-decay position) against the optax
  formulas restated in numpy; grafting.graft with an SGD norm and a fixed direction."""
  from precondition_b200.tearfree import grafting, momentum, praxis_shim
  rng = np.random.default_rng(3)
  p = rng.standard_normal((37, 5)).astype(np.float32)
  dp = [torch.as_tensor(p).cuda()]
  for ema in (False, True):
    for nesterov in (False, True):
      for after in (False, True):
        o = momentum.Options(ema=ema, nesterov=nesterov, momentum_decay=0.6, weight_decay=0.02,
                             weight_decay_after_momentum=after)
        tx = momentum.apply(o)
        st = tx.init(dp)
        v = np.zeros_like(p)
        for _ in range(3):
          g = rng.standard_normal(p.shape).astype(np.float32)
          x = g + np.float32(0.02) * p if not after else g
          if ema:
            x = x * np.float32(1 - 0.6)
          v = x + np.float32(0.6) * v
          x = x + np.float32(0.6) * v if nesterov else v
          if after:
            x = x + np.float32(0.02) * p
          got, st = tx.update([torch.as_tensor(g).cuda()], st, dp)
          np.testing.assert_allclose(got[0].cpu().numpy(), x, rtol=1e-6, atol=1e-7)
  # grafting: the direction is 3 * sign pattern, the norm comes from the raw update (SGD)
  fixed = praxis_shim.ShardedGradientTransformation(
      lambda params: praxis_shim.EmptyState(),
      lambda u, s, params=None: ([3.0 * torch.sign(x) if isinstance(x, torch.Tensor) else x
                                  for x in u], s),
      lambda params: praxis_shim.EmptyState())
  tx = grafting.graft(grafting.Options(grafting_type=grafting.GraftingType.SGD,
                                       second_moment_decay=0.0, start_preconditioning_step=1), fixed)
  vec = torch.as_tensor(rng.standard_normal(9).astype(np.float32)).cuda()
  st = tx.init(dp + [vec])
  g = [torch.as_tensor(rng.standard_normal(p.shape).astype(np.float32)).cuda(), vec.clone()]
  out0, st = tx.update(g, st, dp + [vec])       # step 0 < start: the graft update itself
  torch.testing.assert_close(out0[0], g[0])
  out1, st = tx.update(g, st, dp + [vec])       # step 1: direction with the graft's norm
  want = 3.0 * torch.sign(g[0])
  want = want * (g[0].norm() / want.norm())
  torch.testing.assert_close(out1[0], want, rtol=1e-5, atol=1e-6)
  torch.testing.assert_close(out1[1], vec)      # rank-1: skipped, graft update only
  assert int(st.count) == 2


def test_tearfree_validation_errors():
  """Same ValueErrors as the reference's validators (TF/shampoo.py:174-229, TF/grafting.py:131-163,
  TF/momentum.py:106-117, TF/reshaper.py:83-93)."""
  from precondition_b200.tearfree import grafting, momentum, reshaper, second_order, shampoo
  with pytest.raises(ValueError, match="block_size"):
    shampoo.apply(shampoo.Options(block_size=1))
  with pytest.raises(ValueError, match="update_statistics_freq"):
    shampoo.apply(shampoo.Options(update_statistics_freq=0))
  with pytest.raises(ValueError, match="second_moment_decay"):
    shampoo.apply(shampoo.Options(second_moment_decay=1.5))
  with pytest.raises(ValueError, match="momentum_decay"):
    momentum.apply(momentum.Options(momentum_decay=1.5))
  with pytest.raises(ValueError, match="weight_decay"):
    momentum.apply(momentum.Options(weight_decay=-1.0))
  with pytest.raises(ValueError, match="merge_dims"):
    reshaper.merge(reshaper.Options(merge_dims=1))
  with pytest.raises(ValueError, match="second_moment_decay"):
    grafting.graft(grafting.Options(second_moment_decay=0.0), None)
  with pytest.raises(NotImplementedError):
    grafting.graft(grafting.Options(grafting_type=grafting.GraftingType.ADAFACTOR,
                                    second_moment_decay=0.5), None)
  tx = shampoo.apply(shampoo.Options(block_size=4))
  with pytest.raises(ValueError, match="unit dimensions"):
    tx.init([torch.zeros(3, 1, device="cuda")])
  with pytest.raises(ValueError, match="indivisible"):
    tx.init([torch.zeros(6, 3, device="cuda")])
  with pytest.raises(ValueError, match=">2 large dims"):
    tx.init([torch.zeros(4, 4, 4, device="cuda")])
  with pytest.raises(RuntimeError, match="CUDA"):
    tx.init([torch.zeros(4, 3)])
  from precondition_b200.tearfree import sketchy
  with pytest.raises(ValueError, match="rank"):
    sketchy.apply(sketchy.Options(rank=0))
  with pytest.raises(NotImplementedError):
    sketchy.apply(sketchy.Options(ekfac_svd=True))


def test_tearfree_sketchy_large_axis_against_the_oracle():
  """Sketchy on a 1024 x 768 matrix with rank 32 (the subspace-iteration path of the
  eigen-solver, tcgen05 Gram and mode products): 4 steps against the oracle's QR + SVD."""
  from precondition_b200.tearfree import optimizer
  rng = np.random.default_rng(9)
  shapes = [(1024, 768)]
  kw = dict(learning_rate=0.05, graft="rmsprop", merge_dims=1024, second_order="sketchy",
            sketchy_rank=32, sketchy_decay=0.95, momentum_decay=0.5)
  params = [rng.standard_normal(s).astype(np.float32) for s in shapes]
  ref = T.Tearfree(params, **kw)
  tx = optimizer.tearfree(kw["learning_rate"], _options(kw))
  dparams = [torch.as_tensor(p).cuda() for p in params]
  state = tx.init(dparams)
  low = rng.standard_normal((1024, 40)).astype(np.float32)
  for t in range(4):
    # gradients with a decaying spectrum so that the sketch has something to find
    gr = [(low * np.logspace(0, -2, 40, dtype=np.float32)) @
          rng.standard_normal((40, 768)).astype(np.float32) +
          0.01 * rng.standard_normal(shapes[0]).astype(np.float32)]
    want = ref.update(gr, params)
    got, state = tx.update([torch.as_tensor(x).cuda() for x in gr], state, dparams)
    assert _rel(got[0].cpu().numpy(), want[0]) <= 5e-3, (t, _rel(got[0].cpu().numpy(), want[0]))


def test_tearfree_half_precision_and_strided_leaves():
  """bf16 parameters / gradients and a transposed (non-contiguous) gradient: the state is fp32 and
  contiguous whatever the leaves look like, updates come back in the gradient's dtype and equal
  the fp32 run on the same (bf16-rounded) values."""
  from precondition_b200.tearfree import optimizer
  kw = dict(graft="rmsprop", merge_dims=16, block_size=8, second_moment_decay=0.9,
            momentum_decay=0.5, weight_decay=0.01)
  rng = np.random.default_rng(2)
  shapes = [(16, 8), (24,), (8, 16)]
  p32 = [torch.as_tensor(rng.standard_normal(s).astype(np.float32)).cuda().bfloat16().float()
         for s in shapes]
  tx_a, tx_b = optimizer.tearfree(0.1, _options(kw)), optimizer.tearfree(0.1, _options(kw))
  p16 = [p.bfloat16() for p in p32]
  sa, sb = tx_a.init(p32), tx_b.init(p16)
  for t in range(3):
    g32 = [torch.as_tensor(rng.standard_normal(s).astype(np.float32)).cuda().bfloat16().float()
           for s in shapes]
    g16 = [g.bfloat16() for g in g32]
    g16[2] = g16[2].t().contiguous().t()  # same values, column-major storage
    assert not g16[2].is_contiguous()
    ua, sa = tx_a.update(g32, sa, p32)
    ub, sb = tx_b.update(g16, sb, p16)
    for a, b in zip(ua, ub):
      assert b.dtype == torch.bfloat16
      torch.testing.assert_close(a.bfloat16(), b, rtol=2e-2, atol=1e-3)
  assert sb[1][0].trace[0].dtype == torch.float32 and sb[1][0].trace[2].is_contiguous()
