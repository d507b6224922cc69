"""CPU tests of the product's host-side shape logic against the reference goldens
(no CUDA calls): merge_small_dims, Preconditioner metadata, block ordering."""
import json

import numpy as np
import pytest
import torch

from precondition_b200 import distributed_shampoo as DS


def test_merge_small_dims_matches_reference(golden_shapes):
  for key, want in json.loads(str(golden_shapes["merge_small_dims"])).items():
    shape, md = json.loads(key)
    assert [int(x) for x in DS.merge_small_dims(shape, md)] == want, key


def test_preconditioner_metadata_matches_reference(golden_shapes):
  for key, want in json.loads(str(golden_shapes["preconditioner_meta"])).items():
    shape, bs, mbs, typ = json.loads(key)
    pre = DS.Preconditioner(shape, bs, mbs, True, DS.PreconditionerType(typ), 0)
    assert [[int(a) for a in s] for s in pre.shapes_for_preconditioners()] == want["shapes"]
    assert pre.exponent_for_preconditioner() == want["exponent"]
    assert pre.should_precondition_dims() == want["dims"]


def test_block_partition_order_matches_reference(golden_shapes):
  g = golden_shapes
  x = torch.as_tensor(g["partition/x"])
  bp = DS.BlockPartitioner(x, 4)
  parts = bp.partition(x)
  assert len(parts) == int(g["partition/n"])
  for i, p in enumerate(parts):
    np.testing.assert_array_equal(p.numpy(), g[f"partition/{i}"])
  np.testing.assert_array_equal(bp.merge_partitions(parts).numpy(), x.numpy())


def test_pad_square_matrix_reference_cases():
  """DST:30-72."""
  out = DS.pad_square_matrix(torch.ones(3, 3), 5).numpy()
  want = np.array([[1, 1, 1, 0, 0], [1, 1, 1, 0, 0], [1, 1, 1, 0, 0], [0, 0, 0, 1, 0],
                   [0, 0, 0, 0, 1]], np.float32)
  np.testing.assert_array_equal(out, want)
  np.testing.assert_array_equal(DS.pad_square_matrix(torch.ones(3, 3), 3).numpy(), np.ones((3, 3)))
  with pytest.raises(ValueError):
    DS.pad_square_matrix(torch.ones(3, 3), 2)
  with pytest.raises(ValueError):
    DS.pad_square_matrix(torch.ones(3, 4), 5)


def test_statistic_inventory_of_baseline_configs():
  """BASELINE config 2: MLP 512->2048->512, block 128 -> 276 statistics of 128
  (256 with p=4, 20 with p=2), SURVEY 8(a)."""
  shapes = [(512, 2048), (2048,), (2048, 512), (512,)]
  sizes, exps = [], []
  for s in shapes:
    pre = DS.Preconditioner(s, 128, 4096, True)
    for sh in pre.shapes_for_preconditioners():
      sizes.append(sh[0])
      exps.append(pre.exponent_for_preconditioner())
  assert len(sizes) == 276 and set(sizes) == {128}
  assert exps.count(4) == 256 and exps.count(2) == 20
  # 3x3 convs keep rank 3 after merging -> p = 6
  pre = DS.Preconditioner((3, 3, 512, 512), 1024, 4096, True)
  assert pre._transformed_shape == (9, 512, 512) and pre.exponent_for_preconditioner() == 6


def test_unsupported_options_fail_loudly():
  with pytest.raises(NotImplementedError):  # FD diagnostics
    DS.distributed_shampoo(0.1, 32, frequent_directions=True, compression_rank=4,
                           reuse_preconditioner=True, generate_fd_metrics=True)
  tx = DS.distributed_shampoo(0.1, 32, shard_optimizer_states=True, num_devices_for_pjit=2)
  fns = tx.init(None)  # DS:3660-3673: init hands back the sharded init / pspec / shape functions
  assert isinstance(fns, DS.InitFnState) and callable(fns.init_fn)
  shapes = [(40, 24), (8,), (3, 5)]
  n = sum(len(DS.Preconditioner(sh, 32, 4096, True).shapes_for_preconditioners()) for sh in shapes)
  shp = fns.shape_and_dtype_fn([torch.zeros(sh) for sh in shapes])
  assert shp.stats.global_stats.statistics[0] == (n + n % 2, 32, 32)  # padded to the device count


def test_preconditioning_compute_steps_schedule_matches_reference_formula():
  """DS:44-76: start + (1 - lr / lr0) * end, floored to a multiple of 10, at least 1."""
  lr = lambda step: 0.1 * (1.0 - step / 1000.0)
  f = DS.preconditioning_compute_steps_schedule
  assert f(lr, 20, 200, 0) == 20
  assert f(lr, 20, 200, 500) == 120      # 20 + 0.5 * 200
  assert f(lr, 20, 200, 1000) == 220
  assert f(lr, 1, 5, 100) == 1           # (1.5 // 10) * 10 = 0 -> max(., 1)
  assert f(lr, 25, 0, 300) == 20         # rounded DOWN to a multiple of 10


def test_partition_statistics_is_balanced_and_complete():
  """Cost-balanced partition: every statistic owned once, per-bucket counts
  differ by at most one, total cost per rank within one largest item of the mean."""
  rng = np.random.default_rng(0)
  buckets = [(1024, [4.0 * 1024**3] * 45 + [3.0 * 1024**3] * 34),
             (512, [5.0 * 512**3] * 37), (64, list(rng.uniform(1, 2, 23) * 64**3)), (9, [3.0 * 729] * 16)]
  for world in (1, 2, 3, 8):
    table = DS.partition_statistics(buckets, world)
    load = np.zeros(world)
    for key, costs in buckets:
      owned = table[key]
      assert sorted(i for o in owned for i in o) == list(range(len(costs)))
      counts = [len(o) for o in owned]
      assert max(counts) - min(counts) <= 1
      for r, o in enumerate(owned):
        load[r] += sum(costs[i] for i in o)
    assert load.max() - load.min() <= 4.0 * 1024**3 + 1e-6, (world, load)
  assert DS.newton_gemms_per_iteration(4) == 4 and DS.newton_gemms_per_iteration(2) == 3
  assert DS.newton_gemms_per_iteration(6) == 5 and DS.newton_gemms_per_iteration(8) == 5


def test_gather_layout_round_trip():
  """Packed all-gather layout + scatter indices (pc_select_scatter), emulated with numpy:
  every rank packs its rows, the concatenation is scattered back into state order."""
  owned = [[0, 3, 6], [1, 4], [2, 5]]
  cnt, row = 3, 36  # 3 x 3 fp32 roots
  lay = DS.gather_layout(owned, cnt, [("r", row), ("m", 20)])
  assert lay["lbytes"] % 16 == 0
  state = np.arange(7 * 9, dtype=np.float32).reshape(7, 9)
  recv = np.zeros((3, lay["lbytes"]), np.uint8)
  for r, o in enumerate(owned):
    ro, _ = lay["sections"]["r"]
    mo, _ = lay["sections"]["m"]
    buf = np.zeros(lay["lbytes"], np.uint8)
    rows = np.zeros((cnt, 9), np.float32)
    rows[:len(o)] = state[o]
    met = np.zeros((cnt, 5), np.float32)
    met[:len(o), 1] = o
    buf[ro:ro + cnt * row] = rows.view(np.uint8).reshape(-1)
    buf[mo:mo + cnt * 20] = met.view(np.uint8).reshape(-1)
    recv[r] = buf
  flat = recv.reshape(-1)
  out = np.full((7, 9), -1, np.float32)
  mout = np.full((7, 5), -1, np.float32)
  for j, d in enumerate(lay["dst_index"]):
    if d < 0:
      continue
    so = lay["src_offset"]["r"][j]
    out[d] = flat[so:so + row].view(np.float32)
    mout[d] = flat.view(np.float32)[lay["metrics_offset"][j]:lay["metrics_offset"][j] + 5]
  assert np.array_equal(out, state)
  assert np.array_equal(mout[:, 1], np.arange(7))


def test_option_validation():
  with pytest.raises(ValueError):
    DS.distributed_shampoo(0.1, 32, exponent_override=17)
  with pytest.raises(NotImplementedError):
    DS.distributed_shampoo(0.1, 32, frequent_directions=True, compression_rank=4,
                           reuse_preconditioner=True, generate_fd_metrics=True)
  DS.distributed_shampoo(0.1, 32, generate_fd_metrics=True)  # ignored without FD (DS:2026)
  DS.distributed_shampoo(lambda s: 0.1, 32, decay_preconditioning_compute_steps=True,
                         end_preconditioning_compute_steps=100)


def test_tearfree_blocking_and_reshaper_layout():
  """Host-side layout logic of the tearfree front-end on CPU tensors: block n of the blocked
  layout is the n-th block in the reference's order (left blocked axis major, TF/shampoo.py:302-
  362), blockify / deblockify round-trip, and merge / unmerge follow TF/reshaper.py:52-133."""
  import torch
  from oracle import tearfree as T
  from precondition_b200.tearfree import reshaper, shampoo
  opts = shampoo.Options(block_size=4)
  rng = np.random.default_rng(0)
  for shape in [(3, 8, 12, 2), (8, 3), (3, 2), (4,), (12, 8)]:
    x = rng.standard_normal(shape).astype(np.float32)
    meta = shampoo._blocks_metadata(opts, list(shape), "p")
    xb = shampoo._blockify(torch.as_tensor(x), meta)
    assert list(xb.shape) == [meta.num_blocks] + meta.block_sizes
    for n, sl in enumerate(T.blocks_of(x, 4)):
      np.testing.assert_array_equal(xb[n].numpy(), x[sl])
    np.testing.assert_array_equal(shampoo._deblockify(xb.contiguous(), meta).numpy(), x)
  ro = reshaper.Options(merge_dims=6, block_size=4)
  merge_tx, unmerge_tx = reshaper.merge(ro), reshaper.unmerge(ro)
  params = {"a": torch.as_tensor(rng.standard_normal((3, 2, 5)).astype(np.float32)),
            "b": [torch.as_tensor(rng.standard_normal((7,)).astype(np.float32))]}
  merged, _ = merge_tx.update(params, merge_tx.init(params), params)
  assert list(merged["a"].shape) == [8, 8] and list(merged["b"][0].shape) == [8]
  np.testing.assert_array_equal(merged["a"].numpy(),
                                T.merge(params["a"].numpy(), 6, 4))
  back, _ = unmerge_tx.update(merged, None, params)
  for got, want in ((back["a"], params["a"]), (back["b"][0], params["b"][0])):
    np.testing.assert_array_equal(got.numpy(), want.numpy())


def test_tearfree_chain_and_partition_specs():
  """sharded_chain threads updates and states like TF/praxis_shim.py:45-90 and the optimizer's
  init_partition_spec mirrors the nesting of its state."""
  import torch
  from precondition_b200.tearfree import optimizer, praxis_shim
  double = praxis_shim.ShardedGradientTransformation(
      lambda p: praxis_shim.EmptyState(), lambda u, s, p=None: ([2 * x for x in u], s),
      lambda p: "spec")
  chain = praxis_shim.sharded_chain(double, double)
  st = chain.init([torch.ones(2)])
  out, st2 = chain.update([torch.ones(2)], st)
  assert out[0].tolist() == [4.0, 4.0] and len(st2) == 2
  assert chain.init_partition_spec(None).inner_state == ("spec", "spec")
  with pytest.raises(ValueError, match="number of updates and states"):
    chain.update([torch.ones(2)], st[:1])
  tx = optimizer.tearfree(0.1, optimizer.TearfreeOptions())
  spec = tx.init_partition_spec({"w": praxis_shim.WeightHParams([2048, 512], None, torch.float32,
                                                              None, [-1, -1])})
  graft_spec, mom_spec, _ = spec.inner_state
  blocks = graft_spec["direction"].inner_state[1]["blocks"]["w"]
  assert [tuple(s.shape) for s in blocks["stats"]] == [(2, 1024, 1024), (2, 512, 512)]
  assert tuple(graft_spec["norm"].acc["w"].shape) == (2048, 512)
  assert tuple(mom_spec.inner_state[0].trace["w"].shape) == (2048, 512)


def test_grouped_gemm_routing_rules():
  """Which kernel family a product descriptor is sent to (host-side rules, no GPU): multiples
  of 128 and large ragged blocks with n % 4 == 0 -> tcgen05; fused (de)quantisation only for
  whole 128 x 128 tiles; k <= 4 into a large output -> streaming outer product."""
  from precondition_b200 import _lib, ops

  def desc(m, n, k, c=0x1000, c_sii=None, c_sio=0, c_in=0):
    d = _lib.GemmDesc()
    d.m, d.n, d.k = m, n, k
    d.c, d.c_in = c, c_in or None
    d.c_sii, d.c_sio = (n if c_sii is None else c_sii), c_sio
    return d

  assert ops.tc_gemm_eligible(desc(1024, 1024, 256))
  assert ops.tc_gemm_eligible(desc(128, 128, 1))              # multiples of 128: always
  assert ops.tc_gemm_eligible(desc(1000, 1000, 1024))         # ragged, large, n % 4 == 0
  assert ops.tc_gemm_eligible(desc(64, 576, 576))
  assert ops.tc_gemm_eligible(desc(576, 576, 64))
  assert not ops.tc_gemm_eligible(desc(147, 147, 64))         # n % 4 != 0
  assert not ops.tc_gemm_eligible(desc(64, 64, 256))          # ragged and too small to pay
  assert not ops.tc_gemm_eligible(desc(1000, 1000, 1))
  assert not ops.tc_gemm_eligible(desc(1, 1024, 1024))        # a vector: the streaming kernel
  assert not ops.tc_gemm_eligible(desc(1024, 1024, 256, c=0x1004))        # C not 16-byte aligned
  assert not ops.tc_gemm_eligible(desc(1024, 1024, 256, c_sii=1026))      # row stride % 4 != 0
  assert ops.tc_gemm_fused_quant_eligible(desc(2048, 2048, 1024))
  assert not ops.tc_gemm_fused_quant_eligible(desc(1000, 1000, 1024))
  assert ops.thin_outer_eligible(desc(1024, 1024, 1))
  assert ops.thin_outer_eligible(desc(64, 64, 4))
  assert not ops.thin_outer_eligible(desc(1024, 1024, 5))
  assert not ops.thin_outer_eligible(desc(8, 8, 1))           # tiny blocks stay on the tile kernel
