"""GPU diagnostic: kernel breakdown of one tearfree (blocked Shampoo, block 256) update."""
import sys
import torch
sys.path.insert(0, ".")
from torch.profiler import profile, ProfilerActivity
from precondition_b200.tearfree import optimizer, second_order, shampoo
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(6)
shapes = [(1024, 1024)] * 4 + [(1024, 4096), (4096, 1024), (1024,), (4096,), (1024,)]
params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
so = second_order.Options(merge_dims=1024, shampoo_options=shampoo.Options(block_size=int(sys.argv[1]) if len(sys.argv) > 1 else 256))
tx = optimizer.tearfree(0.1, optimizer.TearfreeOptions(second_order_options=so))
state = tx.init(params)
for _ in range(3):
  g = [torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes]
  _, state = tx.update(g, state, params)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  _, state = tx.update(g, state, params)
  torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in evs:
  a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
span = max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)
print(f"device span {span / 1e3:.2f} ms, summed kernel time {sum(v[1] for v in agg.values()) / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:14]:
  print(f"  {k:70s} {v[0]:5d} {v[1] / 1e3:9.2f} ms")
