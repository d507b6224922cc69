"""Records golden vectors by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Build-container only: needs ``/root/reference``.  The reference sources are
imported over ``oracle/jax_shim`` (numpy array backend; JAX is not installable
here) and executed through their own public functions.  Output: small ``.npz``
fixtures under ``tests/golden/`` that travel to the GPU box.

  python -m oracle.gen_golden            # regenerate everything

Fixtures
  roots.npz      matrix_inverse_pth_root / power_iteration / mat_power cases,
                 incl. the generators of DST:341-408 (spectra, padding, all-padding)
  roots_f64.npz  the same routine with jax_enable_x64 semantics (ground truth)
  quant.npz      QuantizedValue int8 / int16 / bf16 round trips (QU:49-113)
  fd.npz         _fd_update_root, _low_rank_root, frequent_directions_update
  optimizer.npz  distributed_shampoo(...).update trajectories for 14 configs
  shapes.npz     merge_small_dims / BlockPartitioner / Preconditioner metadata
  sm3.npz        precondition/sm3.py update trajectories (rank 1-4 parameters, 3 option sets)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy.stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def gen_symmetric_matrix(rng, dim, condition_number):
  """DST:341-346."""
  u = scipy.stats.ortho_group.rvs(dim=dim, random_state=rng).astype(np.float64)
  diag = np.diag([condition_number**(-i / (dim - 1)) for i in range(dim)])
  return u @ diag @ u.T


def ema_statistics(rng, n, m, steps=20, beta2=0.999, eps=1e-6, scale=1e-2):
  """S0 = eps I; S <- beta2 S + (1-beta2) G G^T  (DS:2594, DS:2635-2636)."""
  s = eps * np.eye(n)
  for _ in range(steps):
    g = rng.standard_normal((n, m)) * scale
    s = beta2 * s + (1 - beta2) * (g @ g.T)
  return s


def root_cases():
  rng = np.random.default_rng(1234)
  cases = []
  for e in range(2, 12):  # DST:348-365
    cases.append(dict(name=f"dst_spectrum_1e{e}", a=gen_symmetric_matrix(rng, 16, 10**e),
                      p=4, ridge=1e-12, pad=None))
  for sz in (4, 32):  # DST:367-398
    ms = gen_symmetric_matrix(rng, sz, 1e3).astype(np.float32) * np.float32(1e-3)
    cases.append(dict(name=f"dst_pad_base_{sz}", a=ms, p=4, ridge=1e-3, pad=None))
    padded = np.eye(2 * sz, dtype=np.float32)
    padded[:sz, :sz] = ms
    cases.append(dict(name=f"dst_pad_padded_{sz}", a=padded, p=4, ridge=1e-3, pad=sz))
  cases.append(dict(name="dst_all_padding", a=np.eye(10, dtype=np.float32), p=4,
                    ridge=1e-3, pad=0))  # DST:400-408
  for p in (1, 2, 3, 4, 6, 8):
    cases.append(dict(name=f"p{p}_n33", a=gen_symmetric_matrix(rng, 33, 1e3), p=p,
                      ridge=1e-6, pad=None))
  cases.append(dict(name="n128_k1e4", a=gen_symmetric_matrix(rng, 128, 1e4), p=4,
                    ridge=1e-6, pad=None))
  cases.append(dict(name="n128_k1e6_p2", a=gen_symmetric_matrix(rng, 128, 1e6), p=2,
                    ridge=1e-6, pad=None))
  cases.append(dict(name="n128_ema", a=ema_statistics(rng, 128, 512), p=4,
                    ridge=1e-6, pad=None))
  cases.append(dict(name="n96_in_128_ema", a=np.pad(ema_statistics(rng, 96, 300),
                                                    ((0, 32), (0, 32))) + np.diag(
                                                        [0.0] * 96 + [1.0] * 32),
                    p=6, ridge=1e-6, pad=96))
  cases.append(dict(name="n256_k1e2", a=gen_symmetric_matrix(rng, 256, 1e2), p=4,
                    ridge=1e-6, pad=None))
  cases.append(dict(name="n256_ema_p2", a=ema_statistics(rng, 256, 64), p=2,
                    ridge=1e-6, pad=None))
  cases.append(dict(name="rank_deficient_retry", a=np.outer(np.arange(1, 25.0),
                                                            np.arange(1, 25.0)),
                    p=4, ridge=1e-12, pad=None))
  cases.append(dict(name="abs_eps", a=gen_symmetric_matrix(rng, 20, 10.0), p=4,
                    ridge=1e-4, pad=None, relative=False))
  cases.append(dict(name="n1", a=np.array([[3.0]]), p=4, ridge=1e-6, pad=None))
  cases.append(dict(name="zero_matrix", a=np.zeros((12, 12)), p=4, ridge=1e-6, pad=None))
  return cases


def run_roots(x64):
  from oracle.jax_shim import import_reference
  DS, _ = import_reference(x64=x64)
  import jax.numpy as jnp
  out = {}
  names = []
  for c in root_cases():
    a32 = np.asarray(c["a"], dtype=np.float32)
    if c["name"] == "n1":
      # reference bug DS:850-907: total_retries unassigned for 1x1; skip there.
      continue
    root, m = DS.matrix_inverse_pth_root(
        jnp.array(a32), c["p"], ridge_epsilon=c["ridge"],
        relative_matrix_epsilon=c.get("relative", True), padding_start=c["pad"])
    names.append(c["name"])
    k = c["name"]
    out[f"{k}/a"] = a32
    out[f"{k}/p"] = np.int32(c["p"])
    out[f"{k}/ridge"] = np.float64(c["ridge"])
    out[f"{k}/pad"] = np.int32(-1 if c["pad"] is None else c["pad"])
    out[f"{k}/relative"] = np.bool_(c.get("relative", True))
    out[f"{k}/root"] = np.asarray(root)
    out[f"{k}/metrics"] = np.array([
        m.inverse_pth_root_errors, m.inverse_pth_root_iters, m.final_error_ratio,
        m.max_eigen_value, m.total_retries], dtype=np.float64)
  if not x64:
    # power_iteration and mat_power on their own
    rng = np.random.default_rng(99)
    a = gen_symmetric_matrix(rng, 40, 1e2).astype(np.float32)
    v, s = DS.power_iteration(jnp.array(a))
    out["pi/a"], out["pi/v"], out["pi/s"] = a, np.asarray(v), np.asarray(s)
    v, s = DS.power_iteration(jnp.array(a), padding_start=25)
    out["pi_pad/v"], out["pi_pad/s"] = np.asarray(v), np.asarray(s)
    m = (rng.standard_normal((12, 12)) * 0.3).astype(np.float32)
    out["mp/m"] = m
    for p in (1, 2, 3, 4, 5, 6, 7, 8):
      out[f"mp/p{p}"] = np.asarray(DS.mat_power(jnp.array(m), p))
    out["pad_square/3to5"] = np.asarray(
        DS.pad_square_matrix(jnp.ones((3, 3), jnp.float32), 5))  # DST:35-55
  out["names"] = np.array(names)
  return out


def run_quant():
  from oracle.jax_shim import import_reference
  _, QU = import_reference()
  import jax.numpy as jnp
  rng = np.random.default_rng(5)
  out = {}
  mats = {
      "sym24": (lambda g: g @ g.T)(rng.standard_normal((24, 40))).astype(np.float32),
      "rect": rng.standard_normal((7, 13)).astype(np.float32),
      "zero_col": np.concatenate([rng.standard_normal((9, 5)), np.zeros((9, 2))],
                                 1).astype(np.float32),
      "halves": (np.arange(-8, 8, dtype=np.float32)[:, None] * 0.5 +
                 np.zeros((1, 3), np.float32)),
      "rank3": rng.standard_normal((4, 5, 6)).astype(np.float32),
  }
  for name, m in mats.items():
    out[f"{name}/x"] = m
    for dt, tag in ((jnp.int8, "i8"), (jnp.int16, "i16")):
      for ext in (False, True):
        if ext and (m.ndim != 2 or m.shape[0] != m.shape[1]):
          continue
        qv = QU.QuantizedValue.from_float_value(jnp.array(m), dt, ext)
        key = f"{name}/{tag}{'_diag' if ext else ''}"
        out[f"{key}/q"] = np.asarray(qv.quantized)
        out[f"{key}/bucket"] = np.asarray(qv.bucket_size)
        if ext:
          out[f"{key}/diag"] = np.asarray(qv.diagonal)
        out[f"{key}/float"] = np.asarray(qv.to_float())
    qv = QU.QuantizedValue.from_float_value(jnp.array(m), jnp.bfloat16, False)
    out[f"{name}/bf16/float"] = np.asarray(qv.to_float()).astype(np.float32)
  return out


def run_fd():
  from oracle.jax_shim import import_reference
  DS, _ = import_reference()
  import jax.numpy as jnp
  rng = np.random.default_rng(11)
  out = {}
  d, r = 24, 4
  prev = np.zeros((d, r + 2), np.float32)
  for step in range(4):
    g = (rng.standard_normal((d, 10 if step % 2 else 40)) * 0.5).astype(np.float32)
    fac = DS.frequent_directions_update(None, jnp.array(g), 0, 1.0, 1.0)
    new, m = DS._fd_update_root(fac, 4, rank=r, ridge_epsilon=1e-6, decay=0.9,
                                padding_start=d, prev=jnp.array(prev))
    out[f"fd/{step}/g"] = g
    out[f"fd/{step}/factor"] = np.asarray(fac)
    out[f"fd/{step}/prev"] = prev
    out[f"fd/{step}/new"] = np.asarray(new)
    prev = np.asarray(new)
  # padded variant
  prev = np.zeros((d, r + 2), np.float32)
  for step in range(3):
    g = np.zeros((d, d), np.float32)
    g[:17, :9] = rng.standard_normal((17, 9))
    fac = DS.frequent_directions_update(None, jnp.array(g), 0, 1.0, 1.0)
    new, _ = DS._fd_update_root(fac, 2, rank=r, ridge_epsilon=1e-6, decay=1.0,
                                padding_start=17, prev=jnp.array(prev))
    out[f"fdpad/{step}/factor"] = np.asarray(fac)
    out[f"fdpad/{step}/prev"] = prev
    out[f"fdpad/{step}/new"] = np.asarray(new)
    prev = np.asarray(new)
  a = gen_symmetric_matrix(rng, 20, 1e3).astype(np.float32)
  for cr in (3, -3):
    for pad in (20, 15):
      root, m = DS._low_rank_root(jnp.array(a), 4, compression_rank=cr,
                                  ridge_epsilon=1e-6, padding_start=pad)
      out[f"lowrank/{cr}/{pad}/root"] = np.asarray(root)
      out[f"lowrank/{cr}/{pad}/err"] = np.asarray(m.inverse_pth_root_errors)
  out["lowrank/a"] = a
  return out


OPT_SHAPES = ([12, 9], [20], [3, 3, 8, 8], [5, 1, 7])
OPT_CONFIGS = {
    "default": {},
    "two_devices": {"_D": 2},
    "three_devices": {"_D": 3},
    "quantized_int16": {"best_effort_memory_usage_reduction": True, "_D": 2},
    "quantized_reuse": {"best_effort_memory_usage_reduction": True,
                        "reuse_preconditioner": True},
    "fd": {"compression_rank": 2, "frequent_directions": True,
           "reuse_preconditioner": True},
    "fd_avg_reset": {"compression_rank": 2, "frequent_directions": True,
                     "reuse_preconditioner": True, "average_grad": True,
                     "reset_preconditioner": True, "beta2": 0.8, "_D": 2},
    "fd_every2": {"compression_rank": 2, "frequent_directions": True,
                  "reuse_preconditioner": True, "preconditioning_compute_steps": 2,
                  "statistics_compute_steps": 2, "average_grad": True},
    "lowrank_pos": {"compression_rank": 3},
    "lowrank_neg": {"compression_rank": -3, "preconditioning_compute_steps": 3},
    "sqrt_n_input": {"statistics_compute_steps": 2, "graft_type": 5,
                     "precondtioner_type": 2, "skip_preconditioning_rank_lt": 2,
                     "merge_small_dims_block_size": 16},
    "adagrad_norm_output_wd": {"graft_type": 6, "precondtioner_type": 3,
                               "skip_preconditioning_rank_lt": 2,
                               "merge_small_dims_block_size": 16,
                               "decoupled_weight_decay": True, "weight_decay": 0.1,
                               "decoupled_learning_rate": False},
    "rmsprop_clip": {"graft_type": 3, "clip_by_scaled_gradient_norm": 0.5,
                     "exponent_override": 3, "start_preconditioning_step": 2},
    "rmsprop_norm_ma": {"graft_type": 4, "weight_decay": 0.01,
                        "moving_average_for_momentum": True, "nesterov": False},
    "adagrad_rank3": {"graft_type": 2, "merge_small_dims_block_size": 16,
                      "preconditioning_compute_steps": 2},
    "none_noshape": {"graft_type": 0, "best_effort_shape_interpretation": False,
                     "skip_preconditioning_rank_lt": 2},
    "abs_eps_beta2_1": {"relative_matrix_epsilon": False, "beta2": 1.0,
                        "matrix_epsilon": 1e-4},
}
OPT_STEPS = 8


def opt_inputs():
  rng = np.random.default_rng(7)
  params = [rng.standard_normal(s).astype(np.float32) for s in OPT_SHAPES]
  grads = [[(rng.standard_normal(s) * 0.1).astype(np.float32) for s in OPT_SHAPES]
           for _ in range(OPT_STEPS)]
  return params, grads


def run_optimizer():
  from oracle.jax_shim import import_reference
  DS, _ = import_reference()
  import jax
  import jax.numpy as jnp
  params_np, grads_np = opt_inputs()
  out = {"configs": np.array(json.dumps(OPT_CONFIGS))}
  for i, p in enumerate(params_np):
    out[f"param/{i}"] = p
  for t, gs in enumerate(grads_np):
    for i, g in enumerate(gs):
      out[f"grad/{t}/{i}"] = g
  params = tuple(jnp.array(p) for p in params_np)
  for name, cfg in OPT_CONFIGS.items():
    kw = {k: v for k, v in cfg.items() if not k.startswith("_")}
    D = cfg.get("_D", 1)
    if "graft_type" in kw:
      kw["graft_type"] = DS.GraftingType(kw["graft_type"])
    if "precondtioner_type" in kw:
      kw["precondtioner_type"] = DS.PreconditionerType(kw["precondtioner_type"])
    optim = DS.distributed_shampoo(0.1, 8, batch_axis_name="batch", **kw)
    state = optim.init(params)
    for t in range(OPT_STEPS):
      grads = tuple(jnp.array(g) for g in grads_np[t])
      fn = jax.pmap(lambda _: optim.update(grads, state, params), axis_name="batch")  # pylint: disable=cell-var-from-loop
      updates, state = fn(jnp.ones([D]))
      state = jax.tree.map(lambda x: x[0], state)
      for i, u in enumerate(updates):
        out[f"{name}/update/{t}/{i}"] = np.asarray(u[0])
    for i, st in enumerate(state.stats):
      for k, (s, pc) in enumerate(zip(st.statistics, st.preconditioners)):
        s = s.to_float() if hasattr(s, "to_float") else s
        pc = pc.to_float() if hasattr(pc, "to_float") else pc
        out[f"{name}/final_stat/{i}/{k}"] = np.asarray(s)
        out[f"{name}/final_precond/{i}/{k}"] = np.asarray(pc)
      tm = st.training_metrics
      if hasattr(tm, "inverse_pth_root_errors"):
        out[f"{name}/final_metrics/{i}"] = np.stack([
            np.asarray(tm.inverse_pth_root_errors, dtype=np.float32).reshape(-1),
            np.asarray(tm.inverse_pth_root_iters, dtype=np.float32).reshape(-1),
            np.asarray(tm.final_error_ratio, dtype=np.float32).reshape(-1),
            np.asarray(tm.max_eigen_value, dtype=np.float32).reshape(-1),
            np.asarray(tm.total_retries, dtype=np.float32).reshape(-1)], 1)
  # the reference's own end-to-end goldens, DST:93-114, DST:212-258
  rng = np.random.default_rng(1234)
  shape = ([2, 5], [6, 3])

  def make(big):
    x = tuple(rng.standard_normal(size=s) for s in shape)
    if big:
      for xx in x:
        xx[..., 0] *= 100
    return tuple(np.asarray(xx, np.float32) for xx in x)

  small_p = (np.array([[1., 3.], [2., 4.]], np.float32),
             np.array([[3., 4.], [3., 4.]], np.float32))
  small_g = (np.array([[500., 5.], [500., 5.]], np.float32),
             np.array([[300., 3.], [300., 3.]], np.float32))
  large_p, large_g = make(False), make(True)
  for tag, (pp, gg) in {"dst_small": (small_p, small_g),
                        "dst_larger": (large_p, large_g)}.items():
    for q in (False, True):
      optim = DS.distributed_shampoo(
          0.1, 32, batch_axis_name="batch", preconditioning_compute_steps=2,
          best_effort_memory_usage_reduction=q)
      jp = tuple(jnp.array(x) for x in pp)
      jg = tuple(jnp.array(x) for x in gg)
      state = optim.init(jp)
      for t in range(6):
        fn = jax.pmap(lambda _: optim.update(jg, state, jp), axis_name="batch")  # pylint: disable=cell-var-from-loop
        updates, state = fn(jnp.ones([1]))
        state = jax.tree.map(lambda x: x[0], state)
        for i, u in enumerate(updates):
          out[f"{tag}{'_q' if q else ''}/update/{t}/{i}"] = np.asarray(u[0])
    for i in range(2):
      out[f"{tag}/param/{i}"] = pp[i]
      out[f"{tag}/grad/{i}"] = gg[i]
  return out


def run_shapes():
  from oracle.jax_shim import import_reference
  DS, _ = import_reference()
  import jax.numpy as jnp
  out = {}
  merge_cases = [([1, 2, 512, 1, 2048, 1, 3, 4], 1024), ([1, 2, 768, 1, 2048], 1024),
                 ([3, 3, 512, 512], 4096), ([1, 1, 1], 4096), ([7, 7, 3, 64], 4096),
                 ([1024, 16, 64], 4096), ([30522, 1024], 4096), ([5], 4096), ([], 4096)]
  res = {}
  for shape, md in merge_cases:
    res[json.dumps([shape, md])] = [int(x) for x in DS.merge_small_dims(shape, md)]
  out["merge_small_dims"] = np.array(json.dumps(res))
  pres = {}
  for shape, bs, mbs, typ in [([512, 2048], 128, 4096, 1), ([2048], 128, 4096, 1),
                              ([3, 3, 512, 512], 1024, 4096, 1), ([7, 7, 3, 64], 1024, 4096, 1),
                              ([1000, 300], 256, 4096, 2), ([1000, 300], 256, 4096, 3),
                              ([1, 1, 256, 1024], 1024, 4096, 1), ([9, 130, 5], 64, 8, 1)]:
    pre = DS.Preconditioner(jnp.zeros(shape), bs, mbs, True, DS.PreconditionerType(typ), 0)
    pres[json.dumps([shape, bs, mbs, typ])] = dict(
        shapes=[[int(a) for a in s] for s in pre.shapes_for_preconditioners()],
        exponent=int(pre.exponent_for_preconditioner()),
        dims=[bool(b) for b in pre.should_precondition_dims()])
  out["preconditioner_meta"] = np.array(json.dumps(pres))
  # partition / merge round trip ordering
  x = np.arange(11 * 7 * 5, dtype=np.float32).reshape(11, 7, 5)
  bp = DS.BlockPartitioner(jnp.array(x), 4)
  parts = bp.partition(jnp.array(x))
  out["partition/x"] = x
  out["partition/n"] = np.int32(len(parts))
  for i, p in enumerate(parts):
    out[f"partition/{i}"] = np.asarray(p)
  return out


def run_eigh_roots():
  """matrix_inverse_pth_root_eigh (DS:943-1030) of the unmodified reference."""
  from oracle.jax_shim import import_reference
  DS, _ = import_reference()
  import jax.numpy as jnp
  rng = np.random.default_rng(99)
  out = {}
  cases = [("spec1e3_p4", gen_symmetric_matrix(rng, 24, 1e3), 4, None),
           ("spec1e5_p2_pad", gen_symmetric_matrix(rng, 24, 1e5), 2, 17),
           ("ema_p4", ema_statistics(rng, 32, 64), 4, None)]
  # (padding_start = 0 cannot be generated here: the numpy-backed shim's eigh raises on the
  #  NaN matrix that case produces before DS:1026-1030 zeroes the result)
  for name, a, p, pad in cases:
    a = a.astype(np.float32)
    v, m = DS.matrix_inverse_pth_root_eigh(jnp.array(a), p, ridge_epsilon=1e-6,
                                           padding_start=pad)
    out[f"{name}/a"] = a
    out[f"{name}/p"] = np.array(p)
    out[f"{name}/pad"] = np.array(-1 if pad is None else pad)
    out[f"{name}/root"] = np.asarray(v)
    out[f"{name}/err"] = np.asarray(m.inverse_pth_root_errors)
  return out


def run_sm3():
  """Trajectories of the unmodified precondition/sm3.py (SM3:40-168): 4 steps over parameters of
  rank 1-4, three option sets; updates, final accumulators and int8 momenta."""
  from oracle import jax_shim
  _, _, ref = jax_shim.import_reference(extra=("sm3",))
  shapes = [(6, 4), (5,), (2, 3, 4), (3, 1, 2, 5)]
  out = {}
  for tag, kw in (("default", dict()),
                  ("wd_norm", dict(weight_decay=0.01, normalize_grads=True, beta1=0.8)),
                  ("beta2_one", dict(beta2=1.0, diagonal_epsilon=1e-6))):
    rng = np.random.default_rng(7)
    params = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    opt = ref.sm3(0.1, **kw)
    state = opt.init(params)
    for i, p in enumerate(params):
      out[f"{tag}/param{i}"] = p
    for t in range(4):
      g = [(rng.standard_normal(s) * 10**rng.uniform(-3, 0)).astype(np.float32) for s in shapes]
      u, state = opt.update(g, state, params)
      for i in range(len(shapes)):
        out[f"{tag}/grad{t}_{i}"] = g[i]
        out[f"{tag}/update{t}_{i}"] = np.asarray(u[i])
    for i in range(len(shapes)):
      for ax, acc in enumerate(state.stats[i].diagonal_statistics):
        out[f"{tag}/acc{i}_{ax}"] = np.asarray(acc)
      out[f"{tag}/momq{i}"] = np.asarray(state.stats[i].diagonal_momentum.quantized)
  return out


TEARFREE_CASES = {
    # tag: (shapes, steps, kwargs of oracle.tearfree.Tearfree == fields of the reference's options)
    "rmsprop": ([(6, 4), (5,), (8, 3, 2)], 4,
                dict(learning_rate=0.1, graft="rmsprop", start_preconditioning_step=1,
                     merge_dims=4, block_size=4, second_moment_decay=0.9)),
    "sgd_ema": ([(8, 8), (3, 5)], 5,
                dict(learning_rate="schedule", graft="sgd", merge_dims=8, block_size=4,
                     update_preconditioners_freq=2, second_moment_decay=0.8, ema=True,
                     nesterov=False, momentum_decay=0.5, weight_decay=0.01,
                     weight_decay_after_momentum=False, skip_preconditioning_any_dim_gt=6)),
    "none_sum": ([(4, 6), (7,)], 3,
                 dict(learning_rate=0.05, graft="none", merge_dims=8, block_size=4,
                      second_moment_decay=1.0, momentum_decay=0.0, weight_decay=0.1)),
    "sketchy": ([(12, 9), (5,), (6, 3, 4)], 4,
                dict(learning_rate=0.1, graft="rmsprop", merge_dims=4, second_order="sketchy",
                     sketchy_rank=4, sketchy_decay=0.9)),
    "sketchy_abs": ([(40, 24)], 5,
                    dict(learning_rate=0.1, graft="sgd", merge_dims=64, second_order="sketchy",
                         sketchy_rank=6, sketchy_epsilon=1e-5, sketchy_relative_epsilon=False,
                         sketchy_decay=1.0, sketchy_update_freq=2, momentum_decay=0.0)),
    "tc_blocks": ([(256, 128)], 3,
                  dict(learning_rate=0.01, graft="rmsprop", graft_decay=1.0, merge_dims=128,
                       block_size=128, second_moment_decay=0.9, update_statistics_freq=1)),
}


def tearfree_schedule(step):
  return 0.1 / (1.0 + float(step))


def tearfree_inputs(tag):
  shapes, steps, kw = TEARFREE_CASES[tag]
  rng = np.random.default_rng(sum(map(ord, tag)))
  params = [rng.standard_normal(s).astype(np.float32) for s in shapes]
  grads = [[(rng.standard_normal(s) * 10**rng.uniform(-2, 0)).astype(np.float32) for s in shapes]
           for _ in range(steps)]
  return params, grads, kw


def run_tearfree():
  """Trajectories of the unmodified precondition/tearfree/optimizer.py:tearfree (blocked Shampoo
  direction, grafting, momentum, weight decay, learning rate) over the numpy shim: the updates of
  every step and the final Shampoo statistics."""
  import importlib
  from oracle import jax_shim
  jax_shim.import_reference()
  opt = importlib.import_module("precondition.tearfree.optimizer")
  graft = importlib.import_module("precondition.tearfree.grafting")
  so = importlib.import_module("precondition.tearfree.second_order")
  sh = importlib.import_module("precondition.tearfree.shampoo")
  sk = importlib.import_module("precondition.tearfree.sketchy")
  import logging
  logging.getLogger("absl").setLevel(logging.ERROR)
  mom = importlib.import_module("precondition.tearfree.momentum")
  gtype = {"none": graft.GraftingType.NONE, "sgd": graft.GraftingType.SGD,
           "rmsprop": graft.GraftingType.RMSPROP}
  out = {}
  import contextlib, io
  for tag in TEARFREE_CASES:
    params, grads, kw = tearfree_inputs(tag)
    g = lambda k, d: kw.get(k, d)
    options = opt.TearfreeOptions(
        grafting_options=graft.Options(
            grafting_type=gtype[g("graft", "rmsprop")],
            second_moment_decay=g("graft_decay", 0.999), epsilon=g("graft_epsilon", 1e-23),
            start_preconditioning_step=g("start_preconditioning_step", 0),
            skip_preconditioning_any_dim_gt=g("skip_preconditioning_any_dim_gt", 4096),
            skip_preconditioning_rank1=g("skip_preconditioning_rank1", True)),
        second_order_options=so.Options(
            merge_dims=g("merge_dims", 1024),
            second_order_type=(so.SecondOrderType.SKETCHY if g("second_order", "") == "sketchy"
                               else so.SecondOrderType.SHAMPOO),
            sketchy_options=sk.Options(
                epsilon=g("sketchy_epsilon", 1e-7), rank=g("sketchy_rank", 128),
                relative_epsilon=g("sketchy_relative_epsilon", True),
                second_moment_decay=g("sketchy_decay", 0.999),
                update_freq=g("sketchy_update_freq", 1)),
            shampoo_options=sh.Options(
                block_size=g("block_size", 1024),
                update_preconditioners_freq=g("update_preconditioners_freq", 1),
                update_statistics_freq=g("update_statistics_freq", 1),
                second_moment_decay=g("second_moment_decay", 0.999))),
        momentum_options=mom.Options(
            ema=g("ema", False), nesterov=g("nesterov", True),
            momentum_decay=g("momentum_decay", 0.9), weight_decay=g("weight_decay", 0.0),
            weight_decay_after_momentum=g("weight_decay_after_momentum", True)))
    lr = tearfree_schedule if kw["learning_rate"] == "schedule" else kw["learning_rate"]
    tx = opt.tearfree(lr, options)
    state = tx.init(params)
    for t, gr in enumerate(grads):
      with contextlib.redirect_stdout(io.StringIO()):  # the reference prints its einsum formula
        u, state = tx.update(gr, state, params)
      for i, ui in enumerate(u):
        out[f"{tag}/update{t}_{i}"] = np.asarray(ui)
    graft_state = state[0]
    direction = graft_state if kw.get("graft") == "none" else graft_state.direction
    if kw.get("second_order") == "sketchy":
      for i, t in enumerate(direction[1].sketches):
        for a, ax in enumerate(getattr(t, "axes", [])):
          for name in ("eigvals", "inv_eigvals", "tail", "inv_tail"):
            out[f"{tag}/{name}{i}_{a}"] = np.asarray(getattr(ax, name))
          v = np.asarray(ax.eigvecs)
          out[f"{tag}/projector{i}_{a}"] = v @ v.T
      continue
    blocks = direction[1].blocks
    for i, b in enumerate(blocks):
      if hasattr(b, "stats"):
        for a, (st, rt) in enumerate(zip(b.stats, b.roots)):
          out[f"{tag}/stats{i}_{a}"] = np.asarray(st)
          out[f"{tag}/roots{i}_{a}"] = np.asarray(rt)
  return out


def main():
  os.makedirs(OUT, exist_ok=True)
  sys.path.insert(0, ROOT)
  jobs = {
      "roots.npz": lambda: run_roots(False),
      "roots_f64.npz": lambda: run_roots(True),
      "quant.npz": run_quant,
      "roots_eigh.npz": run_eigh_roots,
      "fd.npz": run_fd,
      "optimizer.npz": run_optimizer,
      "shapes.npz": run_shapes,
      "sm3.npz": run_sm3,
      "tearfree.npz": run_tearfree,
  }
  only = set(sys.argv[1:])  # optional: regenerate just the named files
  for fname, fn in jobs.items():
    if only and fname not in only:
      continue
    with np.errstate(all="ignore"):
      data = fn()
    np.savez_compressed(os.path.join(OUT, fname), **data)
    print(fname, len(data), "arrays",
          os.path.getsize(os.path.join(OUT, fname)) // 1024, "KiB")


if __name__ == "__main__":
  main()
