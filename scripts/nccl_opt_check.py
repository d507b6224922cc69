"""Multi-GPU check (torchrun, NCCL): the block-sharded optimizer (batch_axis_name set: each
rank computes a contiguous chunk of the roots / sketch updates, then all-gather, DS:2841-2879)
must produce exactly the updates of the unsharded optimizer on every rank."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from precondition_b200 import distributed_shampoo as DS

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ok = True
for name, kw, shapes, block in [
    ("full roots, block 256", {}, [(512, 768), (300,), (3, 3, 32, 32)], 256),
    ("sketchy rank 8", dict(compression_rank=8, frequent_directions=True, reuse_preconditioner=True,
                            merge_small_dims_block_size=64), [(128, 192), (64, 64)], 64)]:
  rng = np.random.default_rng(0)
  params = [torch.as_tensor(rng.standard_normal(s).astype(np.float32) * 0.05).to(dev) for s in shapes]
  sharded = DS.distributed_shampoo(0.1, block, batch_axis_name="batch", start_preconditioning_step=1, **kw)
  single = DS.distributed_shampoo(0.1, block, start_preconditioning_step=1, **kw)
  s1, s2 = sharded.init(params), single.init(params)
  for t in range(3):
    grads = [torch.as_tensor((rng.standard_normal(s) * 1e-2).astype(np.float32)).to(dev) for s in shapes]
    u1, s1 = sharded.update(grads, s1, params)
    u2, s2 = single.update(grads, s2, params)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(u1, u2)):
      err = float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
      if err > 1e-6:
        ok = False
        print(f"rank {rank} {name} step {t} param {i}: sharded vs single {err:.2e}", flush=True)
  if rank == 0:
    print(f"{name}: sharded == unsharded over {world} ranks: {ok}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
