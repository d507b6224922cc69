"""GPU diagnostic: time every grouped-GEMM launch list of one optimizer (statistics, the apply
passes; tcgen05 lists and the CUDA-core size classes) separately, with the shapes in each class:
  python scripts/gemm_classes_probe.py resnet|bert"""
import sys
import torch
sys.path.insert(0, ".")
import bench
from precondition_b200 import distributed_shampoo as DS, ops

name = sys.argv[1] if len(sys.argv) > 1 else "resnet"
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(3)
if name == "resnet":
  shapes = bench.resnet50_shapes(); opt = DS.distributed_shampoo(0.1, 1024, preconditioning_compute_steps=1)
else:
  shapes = bench.bert_large_shapes()
  opt = DS.distributed_shampoo(0.1, 2048, preconditioning_compute_steps=1,
                               best_effort_memory_usage_reduction=True, batch_axis_name="batch")
params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
state = opt.init(params)
grads = [torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes]
for _ in range(7):
  _, state = opt.update(grads, state, params)
torch.cuda.synchronize()
sh = opt.export_state.__self__


def timed(fn, reps=5):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps


def describe(descs):
  from collections import Counter
  c = Counter((d.m, d.n, d.k) for d in descs)
  fl = sum(2.0 * d.m * d.n * d.k for d in descs)
  by = sum(4.0 * (d.m * d.k + d.n * d.k + d.m * d.n) for d in descs)
  top = ", ".join(f"{k}x{v}" for k, v in c.most_common(6))
  return fl, by, top


def show(tag, lists_tc, lists_simt):
  if lists_tc is not None and lists_tc.count:
    t = timed(lists_tc.run)
    fl, by, top = describe(list(lists_tc.arr)[:lists_tc.count])
    print(f"{tag} tcgen05   {lists_tc.count:4d} descs {t*1e3:8.1f} us  {fl/t/1e9:8.1f} TF/s  {top}")
  if lists_simt is not None:
    for (devd, count, mm, mn), descs in zip(lists_simt.groups, lists_simt.group_descs):
      t = timed(lambda: ops.grouped_gemm(devd, count, mm, mn))
      fl, by, top = describe(descs)
      print(f"{tag} simt      {count:4d} descs {t*1e3:8.1f} us  {fl/t/1e9:8.2f} TF/s "
            f"{by/t/1e6:8.1f} GB/s  {top}")
    for (devd, count, mm, mn, kind), descs in zip(lists_simt.thin, lists_simt.thin_descs):
      t = timed(lambda: ops._lib.load().pc_grouped_gemm_thin(
          ops._ptr(devd), count, mm, mn, kind, __import__("ctypes").c_void_p(ops._stream())))
      fl, by, top = describe(descs)
      print(f"{tag} thin {kind}    {count:4d} descs {t*1e3:8.1f} us  {fl/t/1e9:8.2f} TF/s "
            f"{by/t/1e6:8.1f} GB/s  {top}")
    for (devd, count, mm, mn, splits, ws), descs in zip(lists_simt.splitk, lists_simt.splitk_descs):
      t = timed(lambda: (ops._lib.load().pc_grouped_gemm_splitk(
          ops._ptr(devd), count, mm, mn, splits, ops._ptr(ws), ws.numel(),
          __import__("ctypes").c_void_p(ops._stream()))))
      fl, by, top = describe(descs)
      print(f"{tag} split-K   {count:4d} descs {t*1e3:8.1f} us  {fl/t/1e9:8.2f} TF/s "
            f"{by/t/1e6:8.1f} GB/s  {top}")


show("stats  ", sh._stat_tc, sh._stat_simt)
for j, (tc, simt) in enumerate(zip(sh._apply_tc, sh._apply_simt)):
  show(f"apply {j}", tc, simt)

# kernel-level view of the CUDA-core lists (torch.profiler)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for lst in [sh._stat_simt] + list(sh._apply_simt):
    if lst is not None:
      lst.run()
  torch.cuda.synchronize()
agg = {}
for e in prof.events():
  if e.device_type == torch.autograd.DeviceType.CUDA:
    a = agg.setdefault(e.name[:60], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
  print(f"  {k:60s} {v[0]:4d} {v[1]:9.1f} us")

# rank-1 statistics updates (k = 1) vs the rest of the tcgen05 statistics list
if sh._stat_tc is not None:
  arr = list(sh._stat_tc.arr)[:sh._stat_tc.count]
  for tag, sub in (("k == 1", [d for d in arr if d.k == 1]), ("k > 1", [d for d in arr if d.k > 1])):
    if sub:
      lst = ops.TcGemmList(sub, dev)
      t = timed(lst.run)
      fl, by, top = describe(sub)
      print(f"stats tcgen05 {tag}: {len(sub)} descs {t*1e3:8.1f} us  {top}")
