import sys, torch
sys.path.insert(0, ".")
from precondition_b200 import distributed_shampoo as DS, _lib
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
p = [torch.randn((4096, 4096), device=dev) * 0.05]
opt = DS.distributed_shampoo(0.1, 1024, preconditioning_compute_steps=1000)
st = opt.init(p); sh = opt.init.__self__
sh.gbuf.copy_(torch.randn_like(sh.gbuf) * 1e-2)
for f in (sh._update_statistics, sh._apply_preconditioners):
  for _ in range(3): f()
  torch.cuda.synchronize()
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    f(); torch.cuda.synchronize()
  evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
  print(f.__name__, "span %.1f us" % (evs[-1].time_range.end - evs[0].time_range.start))
  for e in evs: print("   %-60s %8.1f us" % (e.name[:60], e.time_range.end - e.time_range.start))
