"""GPU diagnostic: Sketchy configs of the product optimizer vs the oracle, statistic by
statistic (Gram vs factor, sketch operator, scalar slots)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import optimizer as O
from oracle import numerics as N
from oracle.gen_golden import OPT_CONFIGS, OPT_STEPS, opt_inputs
from precondition_b200 import distributed_shampoo as DS

def op(packed, r):
  vecs, inv, const, skip = N.low_rank_unpack(packed.astype(np.float64), r)
  d = packed.shape[0]
  return np.eye(d) if skip else const * (np.eye(d) - vecs @ vecs.T) + (vecs * inv) @ vecs.T

rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
names = sys.argv[1:] or ["fd"]
params, grads = opt_inputs()
for name in names:
  cfg = {k: v for k, v in OPT_CONFIGS[name].items() if not k.startswith("_")}
  oo = O.distributed_shampoo(0.1, 8, batch_axis_name="batch", **cfg)
  ost = oo.init(params)
  po = DS.distributed_shampoo(0.1, 8, batch_axis_name="batch", **cfg)
  tp = [torch.as_tensor(p).cuda() for p in params]
  pst = po.init(tp)
  r = cfg["compression_rank"]
  print("==", name)
  for t in range(OPT_STEPS):
    with np.errstate(all="ignore"):
      ou, ost = oo.update(grads[t], ost, params)
    pu, pst = po.update([torch.as_tensor(g).cuda() for g in grads[t]], pst, tp)
    torch.cuda.synchronize()
    for i in range(len(params)):
      line = f" t={t} p{i}: upd {rel(pu[i].cpu().numpy(), ou[i]):.1e}"
      for k, (a, b, pa, pb) in enumerate(zip(pst.stats[i].statistics, ost.stats[i].statistics,
                                              pst.stats[i].preconditioners, ost.stats[i].preconditioners)):
        a, pa = a.cpu().numpy(), pa.cpu().numpy()
        if pb.shape[0] != pb.shape[1]:
          line += f" | s{k}[fd {pb.shape[0]}] gram {rel(a, b @ b.T):.0e} op {rel(op(pa, r), op(pb, r)):.0e} slots {rel(pa[:, r:], pb[:, r:]):.0e} hz {pa[-1,-2]:.0f}/{pb[-1,-2]:.0f}"
        else:
          line += f" | s{k}[{pb.shape[0]}] stat {rel(a, b):.0e} prec {rel(pa, pb):.0e}"
      print(line, flush=True)
    if t == 5:
      i = 0
      for k, (pa, pb) in enumerate(zip(pst.stats[i].preconditioners, ost.stats[i].preconditioners)):
        if pb.shape[0] != pb.shape[1]:
          print("   ours slots", np.round(pa.cpu().numpy()[:, r:].T, 5).tolist())
          print("   ref  slots", np.round(pb[:, r:].T, 5).tolist())
          break
