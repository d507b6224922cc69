"""GPU diagnostic: phase split of the ResNet-50 Shampoo step (BASELINE config 3, one GPU)."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from precondition_b200 import distributed_shampoo as DS

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(3)
shapes = bench.resnet50_shapes()
params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
opt = DS.distributed_shampoo(0.1, 1024, preconditioning_compute_steps=1)
state = opt.init(params)
grads = [torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes]
for _ in range(3):
  _, state = opt.update(grads, state, params)
sh = opt.init.__self__
def timed(f, n=3):
  torch.cuda.synchronize(); t0 = time.time()
  for _ in range(n): f()
  torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
def stage():
  torch._foreach_copy_(sh._gviews, [g.reshape(-1) for g in grads])
def graft():
  sh._transform_all(10, 0.1)
print(f"total {timed(lambda: opt.update(grads, state, params)):.2f} ms | stage {timed(stage):.2f} stats {timed(sh._update_statistics):.2f} "
      f"roots {timed(lambda: sh._compute_preconditioners(10)):.2f} apply {timed(sh._apply_preconditioners):.2f} graft {timed(graft):.2f}")
for s, bk in sorted(sh.buckets.items()):
  t = timed(lambda: sh._newton_roots([bk], 1, 0), 2)
  print(f"   bucket {s:5d} x {bk.count:3d} (solved as {bk.job.sp}): roots alone {t:7.2f} ms")
from torch.profiler import profile, ProfilerActivity
for name, f in (("stats", sh._update_statistics), ("apply", sh._apply_preconditioners)):
  torch.cuda.synchronize()
  t0 = time.time()
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    f(); torch.cuda.synchronize()
  evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
  agg = {}
  for e in evs:
    a = agg.setdefault(e.name[:48], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
  span = max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)
  print(name, f"device span {span / 1e3:.2f} ms, busy {sum(v[1] for v in agg.values()) / 1e3:.2f} ms")
  for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:5]:
    print(f"    {k:48s} {v[0]:4d} {v[1] / 1e3:8.2f} ms")
