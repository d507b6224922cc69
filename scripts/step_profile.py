"""GPU diagnostic: kernel breakdown (torch.profiler) of one optimizer update for a named config:
  python scripts/step_profile.py resnet|bert|sketchy|mlp"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from precondition_b200 import distributed_shampoo as DS
from torch.profiler import profile, ProfilerActivity

name = sys.argv[1] if len(sys.argv) > 1 else "resnet"
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(3)
if name == "resnet":
  shapes = bench.resnet50_shapes(); opt = DS.distributed_shampoo(0.1, 1024, preconditioning_compute_steps=1)
elif name == "bert":
  shapes = bench.bert_large_shapes()
  opt = DS.distributed_shampoo(0.1, 2048, preconditioning_compute_steps=1,
                               best_effort_memory_usage_reduction=True, batch_axis_name="batch")
elif name == "sketchy":
  shapes = [(4096, 4096)] * 8
  opt = DS.distributed_shampoo(0.1, 4096, compression_rank=256, frequent_directions=True,
                               reuse_preconditioner=True)
else:
  shapes = [(512, 2048), (2048,), (2048, 512), (512,)]; opt = DS.distributed_shampoo(0.1, 128, graft_type=DS.GraftingType.SGD)
params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
state = opt.init(params)
grads = [torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes]
for _ in range(7):
  _, state = opt.update(grads, state, params)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); _, state = opt.update(grads, state, params); e1.record(); torch.cuda.synchronize()
print(f"{name}: step {e0.elapsed_time(e1):.2f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  _, state = opt.update(grads, state, params)
  torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in evs:
  a = agg.setdefault(e.name[:70], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
span = max(e.time_range.end for e in evs) - min(e.time_range.start for e in evs)
print(f"device span {span / 1e3:.2f} ms, summed kernel time {sum(v[1] for v in agg.values()) / 1e3:.2f} ms")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:18]:
  print(f"  {k:70s} {v[0]:5d} {v[1] / 1e3:9.2f} ms")
