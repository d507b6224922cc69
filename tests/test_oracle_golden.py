"""Pins the CPU oracle to golden vectors recorded from the UNMODIFIED reference
(run over ``oracle/jax_shim`` by ``oracle/gen_golden.py``).  CPU-only."""
import json
import os

import numpy as np
import pytest

from oracle import numerics as N
from oracle import optimizer as O
from oracle.gen_golden import OPT_CONFIGS, OPT_SHAPES, OPT_STEPS


def _root_names(g):
  return [str(n) for n in g["names"]]


def test_roots_match_reference_bitwise(golden_roots):
  g = golden_roots
  for k in _root_names(g):
    pad = int(g[f"{k}/pad"])
    root, m = N.matrix_inverse_pth_root(
        g[f"{k}/a"], int(g[f"{k}/p"]), ridge_epsilon=float(g[f"{k}/ridge"]),
        relative_matrix_epsilon=bool(g[f"{k}/relative"]),
        padding_start=None if pad < 0 else pad)
    np.testing.assert_array_equal(root, g[f"{k}/root"], err_msg=k)
    np.testing.assert_allclose(m.as_row(), g[f"{k}/metrics"].astype(np.float32),
                               rtol=0, atol=0, err_msg=k)


def test_roots_f64_twin(golden_roots_f64):
  g = golden_roots_f64
  for k in _root_names(g):
    if g[f"{k}/a"].shape[0] > 128:
      continue
    pad = int(g[f"{k}/pad"])
    root, m = N.matrix_inverse_pth_root(
        g[f"{k}/a"], int(g[f"{k}/p"]), ridge_epsilon=float(g[f"{k}/ridge"]),
        relative_matrix_epsilon=bool(g[f"{k}/relative"]),
        padding_start=None if pad < 0 else pad, dtype=np.float64)
    # x64 reference returns the root cast back to the f32 input dtype
    np.testing.assert_allclose(root, g[f"{k}/root"], rtol=1e-6, atol=1e-7, err_msg=k)
    assert m.inverse_pth_root_iters == g[f"{k}/metrics"][1], k


def test_power_iteration_and_mat_power(golden_roots):
  g = golden_roots
  v, s = N.power_iteration(g["pi/a"])
  np.testing.assert_array_equal(v, g["pi/v"])
  assert s == g["pi/s"]
  v, s = N.power_iteration(g["pi/a"], padding_start=25)
  np.testing.assert_array_equal(v, g["pi_pad/v"])
  assert s == g["pi_pad/s"]
  for p in range(1, 9):
    np.testing.assert_array_equal(N.mat_power(g["mp/m"], p), g[f"mp/p{p}"])
  np.testing.assert_array_equal(
      N.pad_square_matrix(np.ones((3, 3), np.float32), 5), g["pad_square/3to5"])


def test_quantization(golden_quant):
  g = golden_quant
  for name in ("sym24", "rect", "zero_col", "halves", "rank3"):
    x = g[f"{name}/x"]
    for dt, tag in ((np.int8, "i8"), (np.int16, "i16")):
      for ext in (False, True):
        key = f"{name}/{tag}{'_diag' if ext else ''}"
        if f"{key}/q" not in g:
          continue
        qv = N.QuantizedValue.from_float_value(x, dt, ext)
        np.testing.assert_array_equal(qv.quantized, g[f"{key}/q"], err_msg=key)
        np.testing.assert_array_equal(qv.bucket_size, g[f"{key}/bucket"])
        if ext:
          np.testing.assert_array_equal(qv.diagonal, g[f"{key}/diag"])
        np.testing.assert_array_equal(qv.to_float(), g[f"{key}/float"])
    bf = N.from_bfloat16_bits(N.to_bfloat16_bits(x))
    np.testing.assert_array_equal(bf, g[f"{name}/bf16/float"])


def test_fd_and_low_rank(golden_fd):
  g = golden_fd
  for step in range(4):
    fac = N.frequent_directions_update(None, g[f"fd/{step}/g"], 0, 1.0, 1.0)
    np.testing.assert_allclose(fac, g[f"fd/{step}/factor"], rtol=0, atol=0)
    new, _ = N.fd_update_root(fac, 4, rank=4, ridge_epsilon=1e-6, decay=0.9,
                              padding_start=24, prev=g[f"fd/{step}/prev"])
    np.testing.assert_allclose(new, g[f"fd/{step}/new"], rtol=2e-5, atol=1e-6)
  for step in range(3):
    new, _ = N.fd_update_root(g[f"fdpad/{step}/factor"], 2, rank=4,
                              ridge_epsilon=1e-6, decay=1.0, padding_start=17,
                              prev=g[f"fdpad/{step}/prev"])
    np.testing.assert_allclose(new, g[f"fdpad/{step}/new"], rtol=2e-5, atol=1e-6)
  for cr in (3, -3):
    for pad in (20, 15):
      root, m = N.low_rank_root(g["lowrank/a"], 4, compression_rank=cr,
                                ridge_epsilon=1e-6, padding_start=pad)
      np.testing.assert_allclose(root, g[f"lowrank/{cr}/{pad}/root"], rtol=1e-6,
                                 atol=1e-7)
      np.testing.assert_allclose(m.inverse_pth_root_errors,
                                 g[f"lowrank/{cr}/{pad}/err"], rtol=1e-6)


def test_shape_logic(golden_shapes):
  g = golden_shapes
  for key, want in json.loads(str(g["merge_small_dims"])).items():
    shape, md = json.loads(key)
    assert [int(x) for x in O.merge_small_dims(shape, md)] == want, key
  for key, want in json.loads(str(g["preconditioner_meta"])).items():
    shape, bs, mbs, typ = json.loads(key)
    pre = O.Preconditioner(shape, bs, mbs, True, O.PreconditionerType(typ), 0)
    assert [[int(a) for a in s] for s in pre.shapes_for_preconditioners()] == want["shapes"]
    assert pre.exponent_for_preconditioner() == want["exponent"]
    assert pre.should_precondition_dims() == want["dims"]
  x = g["partition/x"]
  bp = O.BlockPartitioner(x.shape, 4)
  parts = bp.partition(x)
  assert len(parts) == int(g["partition/n"])
  for i, p in enumerate(parts):
    np.testing.assert_array_equal(p, g[f"partition/{i}"])
  np.testing.assert_array_equal(bp.merge_partitions(parts), x)


def _oracle_cfg(cfg):
  kw = {k: v for k, v in cfg.items() if not k.startswith("_")}
  if "graft_type" in kw:
    kw["graft_type"] = O.GraftingType(kw["graft_type"])
  if "precondtioner_type" in kw:
    kw["precondtioner_type"] = O.PreconditionerType(kw["precondtioner_type"])
  return kw, cfg.get("_D", 1)


@pytest.mark.parametrize("name", sorted(OPT_CONFIGS))
def test_optimizer_trajectory(golden_optimizer, name):
  g = golden_optimizer
  kw, D = _oracle_cfg(OPT_CONFIGS[name])
  params = [g[f"param/{i}"] for i in range(len(OPT_SHAPES))]
  opt = O.distributed_shampoo(0.1, 8, batch_axis_name="batch", num_devices=D, **kw)
  state = opt.init(params)
  with np.errstate(all="ignore"):
    for t in range(OPT_STEPS):
      grads = [g[f"grad/{t}/{i}"] for i in range(len(OPT_SHAPES))]
      updates, state = opt.update(grads, state, params)
      for i, u in enumerate(updates):
        np.testing.assert_allclose(u, g[f"{name}/update/{t}/{i}"], rtol=1e-6,
                                   atol=1e-7, err_msg=f"{name} step {t} param {i}")
  for i, st in enumerate(state.stats):
    for k, pc in enumerate(st.preconditioners):
      pc = pc.to_float() if hasattr(pc, "to_float") else pc
      np.testing.assert_allclose(pc, g[f"{name}/final_precond/{i}/{k}"], rtol=1e-5,
                                 atol=1e-6)
    if f"{name}/final_metrics/{i}" in g and st.training_metrics is not None:
      np.testing.assert_allclose(st.training_metrics, g[f"{name}/final_metrics/{i}"],
                                 rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tag,expected", [("dst_small", -0.57), ("dst_small_q", -0.57),
                                          ("dst_larger", -0.17019942),
                                          ("dst_larger_q", -0.17019942)])
def test_reference_end_to_end_goldens(golden_optimizer, tag, expected):
  """DST:116-261: step-0 golden scalars and 6 finite steps."""
  g = golden_optimizer
  base = tag.replace("_q", "")
  params = [g[f"{base}/param/{i}"] for i in range(2)]
  grads = [g[f"{base}/grad/{i}"] for i in range(2)]
  opt = O.distributed_shampoo(0.1, 32, batch_axis_name="batch",
                              preconditioning_compute_steps=2,
                              best_effort_memory_usage_reduction=tag.endswith("_q"))
  state = opt.init(params)
  for t in range(6):
    updates, state = opt.update(grads, state, params)
    if t == 0:
      assert abs(updates[1].reshape(-1)[-1] - expected) < 1e-4
      assert abs(g[f"{tag}/update/0/1"].reshape(-1)[-1] - expected) < 1e-4
    for i, u in enumerate(updates):
      assert np.all(np.isfinite(u))
      np.testing.assert_allclose(u, g[f"{tag}/update/{t}/{i}"], rtol=1e-6, atol=1e-7)


def test_eigh_root_matches_reference_bitwise():
  """matrix_inverse_pth_root_eigh (DS:943-1030): the oracle reproduces the unmodified
  reference (tests/golden/roots_eigh.npz, oracle/gen_golden.py run_eigh_roots) bit for bit."""
  import os
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", "roots_eigh.npz"))
  for name in ("spec1e3_p4", "spec1e5_p2_pad", "ema_p4"):
    pad = int(g[f"{name}/pad"])
    v, m = N.matrix_inverse_pth_root_eigh(g[f"{name}/a"], int(g[f"{name}/p"]),
                                          padding_start=None if pad < 0 else pad)
    np.testing.assert_array_equal(v, g[f"{name}/root"])
    assert np.float32(m.inverse_pth_root_errors) == g[f"{name}/err"]
  v, m = N.matrix_inverse_pth_root_eigh(np.eye(8, dtype=np.float32), 4, padding_start=0)
  assert np.abs(v).sum() == 0 and m.inverse_pth_root_errors == 0  # DS:1026-1030


def test_sm3_oracle_matches_reference_golden():
  """oracle/sm3.py == the unmodified precondition/sm3.py (recorded over the numpy shim), bit for
  bit: updates of 4 steps, final accumulators and int8 momenta, three option sets."""
  from oracle import sm3 as O
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sm3.npz"))
  shapes = [(6, 4), (5,), (2, 3, 4), (3, 1, 2, 5)]
  for tag, kw in (("default", dict()),
                  ("wd_norm", dict(weight_decay=0.01, normalize_grads=True, beta1=0.8)),
                  ("beta2_one", dict(beta2=1.0, diagonal_epsilon=1e-6))):
    params = [g[f"{tag}/param{i}"] for i in range(len(shapes))]
    opt = O.sm3(0.1, **kw)
    state = opt.init(params)
    for t in range(4):
      grads = [g[f"{tag}/grad{t}_{i}"] for i in range(len(shapes))]
      u, state = opt.update(grads, state, params)
      for i in range(len(shapes)):
        assert np.array_equal(u[i], g[f"{tag}/update{t}_{i}"]), (tag, t, i)
    for i in range(len(shapes)):
      for ax, acc in enumerate(state.stats[i].diagonal_statistics):
        assert np.array_equal(acc, g[f"{tag}/acc{i}_{ax}"])
      assert np.array_equal(state.stats[i].diagonal_momentum.quantized, g[f"{tag}/momq{i}"])


def test_tearfree_oracle_matches_reference_golden():
  """oracle/tearfree.py == the unmodified precondition/tearfree (recorded over the numpy shim):
  updates of every step and the final blocked statistics / roots, for the four option sets of
  oracle.gen_golden.TEARFREE_CASES.  The eigendecompositions are the same LAPACK calls, so the
  trajectories agree to rounding of the products around them."""
  from oracle import gen_golden as G
  from oracle import tearfree as T
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", "tearfree.npz"))
  for tag in G.TEARFREE_CASES:
    params, grads, kw = G.tearfree_inputs(tag)
    kw = dict(kw)
    if kw["learning_rate"] == "schedule":
      kw["learning_rate"] = G.tearfree_schedule
    opt = T.Tearfree(params, **kw)
    for t, gr in enumerate(grads):
      outs = opt.update(gr, params)
      for i, u in enumerate(outs):
        want = g[f"{tag}/update{t}_{i}"]
        scale = np.abs(want).max() + 1e-30
        assert np.abs(u - want).max() <= 2e-4 * scale, (tag, t, i, np.abs(u - want).max() / scale)
    for i, leaf in enumerate(opt.leaves):
      if leaf is None:
        continue
      if kw.get("second_order") == "sketchy":
        for a, st in enumerate(leaf.axes):
          for name in ("eigvals", "inv_eigvals", "tail", "inv_tail"):
            want = g[f"{tag}/{name}{i}_{a}"]
            assert np.abs(st[name] - want).max() <= 2e-5 * (np.abs(want).max() + 1e-30), (
                tag, i, a, name)
        continue
      for a in range(len(leaf.shape)):
        want = g[f"{tag}/stats{i}_{a}"]  # [N, B, B]
        got = np.stack([leaf.stats[n][a] for n in range(len(leaf.slices))])
        assert np.abs(got - want).max() <= 1e-5 * (np.abs(want).max() + 1e-30), (tag, i, a)
