"""GPU diagnostic: kernel timeline of one solver call (torch.profiler / CUPTI) --
per-kernel totals, idle gaps between kernels, and the largest gaps with their neighbours."""
import sys
import torch
sys.path.insert(0, ".")
from precondition_b200 import ops, _lib  # noqa: E402
import bench  # noqa: E402

B, n = int(sys.argv[1]) if len(sys.argv) > 1 else 74, 1024
dev = torch.device("cuda", 0)
xs = bench.make_statistics_torch(B, n, seed=1000, device=dev)
ps = torch.full((B,), 4, dtype=torch.int32, device=dev)
roots = torch.empty_like(xs)
for _ in range(3):
  ops.matrix_inverse_pth_root_batched(xs, ps, None, out=roots)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
  ops.matrix_inverse_pth_root_batched(xs, ps, None, out=roots)
  torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
print(f"span {(t1 - t0) / 1e3:.3f} ms, {len(evs)} device activities")
agg = {}
for e in evs:
  k = e.name[:50]
  a = agg.setdefault(k, [0, 0.0])
  a[0] += 1
  a[1] += e.time_range.end - e.time_range.start
busy = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
  print(f"  {k:50s} {v[0]:5d} {v[1] / 1e3:9.3f} ms")
print(f"busy {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms")
gaps = []
for a, b in zip(evs[:-1], evs[1:]):
  g = b.time_range.start - a.time_range.end
  gaps.append((g, a.name[:30], b.name[:30], (a.time_range.end - t0) / 1e3))
gaps.sort(reverse=True)
for g in gaps[:12]:
  print(f"  gap {g[0]:8.1f} us after {g[1]:30s} before {g[2]:30s} at {g[3]:.3f} ms")
