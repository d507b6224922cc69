"""GPU diagnostic: grouped vs per-parameter grafting kernel, per buffer and segment."""
import sys
import torch
sys.path.insert(0, ".")
from precondition_b200 import ops
dev = torch.device("cuda", 0)
for graft in range(7):
  gen = torch.Generator(device=dev).manual_seed(graft)
  numels = [1, 37, 8192, 8193, 70001, 4]
  skip = [False, True, False, False, False, True]
  offs, total = [], 0
  for n_ in numels:
    offs.append(total); total += (n_ + 31) // 32 * 32
  mk = lambda scale=1.0: torch.randn(total, generator=gen, device=dev) * scale
  grad, param, pg = mk(1e-2), mk(0.1), mk(3.0)
  diag, dmom, mom = mk().abs(), mk(1e-2), mk(1e-2)
  opt = ops.make_graft_options(
      beta1=0.9, beta2=0.999, graft_type=graft, diagonal_epsilon=1e-10, weight_decay=1e-3,
      learning_rate=0.1, nesterov=1, moving_average_for_momentum=int(graft % 2),
      decoupled_learning_rate=int(graft % 3 != 0), decoupled_weight_decay=int(graft % 2 == 0),
      run_shampoo=1, clip_by_scaled_gradient_norm=0.5 if graft in (3, 4) else 0.0)
  want = [t.clone() for t in (diag, dmom, mom)]
  want_u = torch.zeros(total, device=dev)
  for o, n_, sk in zip(offs, numels, skip):
    sl = slice(o, o + n_)
    ops.graft_momentum(grad[sl], param[sl], None if sk else pg[sl], want[0][sl], want[1][sl],
                       want[2][sl], want_u[sl], opt)
  got = [t.clone() for t in (diag, dmom, mom)]
  got_u = torch.zeros(total, device=dev)
  group = ops.GraftGroup([(o, n_, not sk) for o, n_, sk in zip(offs, numels, skip)], dev)
  group.run(grad, param, pg, got[0], got[1], got[2], got_u, opt)
  torch.cuda.synchronize()
  for o, n_ in zip(offs, numels):
    sl = slice(o, o + n_)
    errs = []
    for a, b in zip(got + [got_u], want + [want_u]):
      errs.append(float(((a[sl] - b[sl]).abs() / b[sl].abs().clamp_min(1e-12)).max()))
    print(graft, n_, " ".join(f"{e:.2e}" for e in errs))
