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
  for kw in (dict(lobpcg_topk_precondition=2), dict(shard_optimizer_states=True)):
    with pytest.raises(NotImplementedError):
      DS.distributed_shampoo(0.1, 32, **kw)
