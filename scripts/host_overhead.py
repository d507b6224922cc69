"""GPU diagnostic: host enqueue time vs device time of one optimizer update (ResNet-50 shapes)."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from precondition_b200 import distributed_shampoo as DS
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(3)
shapes = bench.resnet50_shapes()
params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
opt = DS.distributed_shampoo(0.1, 1024, preconditioning_compute_steps=1)
state = opt.init(params)
grads = [[torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes] for _ in range(8)]
for t in range(4):
  _, state = opt.update(grads[t], state, params)
torch.cuda.synchronize()
for t in range(4, 8):
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter(); e0.record()
  _, state = opt.update(grads[t], state, params)
  e1.record(); t1 = time.perf_counter()
  torch.cuda.synchronize()
  print(f"host enqueue {1e3 * (t1 - t0):.2f} ms, device {e0.elapsed_time(e1):.2f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
_, state = opt.update(grads[0], state, params)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
