"""GPU diagnostic: achieved HBM bandwidth of the grafting + momentum tail (north_star item 4) on one
large parameter, and of the power iteration inside the solver (via torch.profiler durations)."""
import sys, torch
sys.path.insert(0, ".")
from precondition_b200 import ops
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
n = 1 << 26  # 64 Mi elements = 256 MB per tensor
g, pg = torch.randn(n, device=dev), torch.randn(n, device=dev)
mom, dmom, upd = torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.empty(n, device=dev)
diag = torch.zeros(n, device=dev)
# (graft_type, needs diagonal statistics, algorithmic bytes / element over all passes)
for name, gt, dg, nbytes in (("SGD", 1, None, 28 + 8), ("RMSPROP_NORMALIZED", 4, diag, 36 + 12 + 4)):
  opt = ops.make_graft_options(beta1=0.9, beta2=0.999, graft_type=gt, diagonal_epsilon=1e-10, weight_decay=0.0,
                               learning_rate=0.1, nesterov=1, moving_average_for_momentum=0,
                               decoupled_learning_rate=1, decoupled_weight_decay=0, run_shampoo=1,
                               clip_by_scaled_gradient_norm=0.0)
  for _ in range(2):
    ops.graft_momentum(g, None, pg, dg, dmom, mom, upd, opt)
  torch.cuda.synchronize()
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
      ops.graft_momentum(g, None, pg, dg, dmom, mom, upd, opt)
    torch.cuda.synchronize()
  agg = {}
  for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
      a = agg.setdefault(e.name[:50], [0, 0.0]); a[0] += 1; a[1] += e.time_range.end - e.time_range.start
  tot = 0.0
  for k, v in agg.items():
    print(f"  {k:50s} x{v[0]}  avg {v[1] / v[0]:9.1f} us"); tot += v[1] / 3
  print(f"{name} graft + momentum tail, {n} elements: {tot:.1f} us per call -> {nbytes * n / tot / 1e3:.0f} GB/s "
        f"({nbytes} algorithmic B/element over the norm and apply passes)")
