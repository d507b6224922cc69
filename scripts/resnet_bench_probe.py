"""GPU diagnostic: per-step times of bench.time_resnet50_step's loop (distinct gradients per step)."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from precondition_b200 import distributed_shampoo as DS, ops

dev = torch.device("cuda", 0)
for r in range(2):
  print("bench.time_resnet50_step:", round(bench.time_resnet50_step(dev, 1)["ms"], 2), "ms")
gen = torch.Generator(device=dev); gen.manual_seed(3)
shapes = bench.resnet50_shapes()
params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
opt = DS.distributed_shampoo(0.1, 1024, preconditioning_compute_steps=1)
state = opt.init(params)
grads = [[torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes] for _ in range(8)]
for t in range(8):
  torch.cuda.synchronize(); t0 = time.time(); l0 = ops.gpu_launches
  _, state = opt.update(grads[t], state, params)
  torch.cuda.synchronize()
  tm = torch.cat([st.training_metrics for st in state.stats if st.training_metrics is not None])
  print(f"step {t}: {(time.time() - t0) * 1e3:8.2f} ms  launches {ops.gpu_launches - l0}  "
        f"max iters {float(tm[:, 1].max()) if tm.shape[1] > 1 else -1}  max err {float(tm[:, 0].max()):.2e}")
