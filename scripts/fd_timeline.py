"""GPU diagnostic: kernel breakdown of one Sketchy update at 4096 x 4096, rank 256, batch 2."""
import sys, torch
sys.path.insert(0, ".")
from precondition_b200 import ops
from torch.profiler import profile, ProfilerActivity
d, rank, batch = 4096, 256, 2
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(5)
xs = torch.randn((batch, d, d), generator=g, device=dev) * 0.01
prev = torch.zeros((batch, d, rank + 2), device=dev)
for _ in range(2):
  prev, _ = ops.fd_update_root_batched(xs, prev, [4] * batch, rank, decay=0.999)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  ops.fd_update_root_batched(xs, prev, [4] * batch, rank, decay=0.999)
  torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in evs:
  a = agg.setdefault(e.name[:60], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
span = max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)
print(f"span {span / 1e3:.1f} ms")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
  print(f"  {k:60s} {v[0]:4d} {v[1] / 1e3:9.2f} ms  avg {v[1] / v[0] / 1e3:8.3f} ms")
