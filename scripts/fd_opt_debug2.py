import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import optimizer as O
from oracle import numerics as N
from oracle.gen_golden import OPT_CONFIGS, OPT_STEPS, opt_inputs
from precondition_b200 import distributed_shampoo as DS
np.set_printoptions(precision=5, suppress=True, linewidth=220)
def op(packed, r):
  vecs, inv, const, skip = N.low_rank_unpack(packed.astype(np.float64), r)
  d = packed.shape[0]
  return np.eye(d) if skip else const * (np.eye(d) - vecs @ vecs.T) + (vecs * inv) @ vecs.T
params, grads = opt_inputs()
cfg = {k: v for k, v in OPT_CONFIGS["fd"].items() if not k.startswith("_")}
oo = O.distributed_shampoo(0.1, 8, batch_axis_name="batch", **cfg)
ost = oo.init(params)
po = DS.distributed_shampoo(0.1, 8, batch_axis_name="batch", **cfg)
tp = [torch.as_tensor(p).cuda() for p in params]
pst = po.init(tp)
sh = po.init.__self__
for t in range(6):
  ou, ost = oo.update(grads[t], ost, params)
  pu, pst = po.update([torch.as_tensor(g).cuda() for g in grads[t]], pst, tp)
torch.cuda.synchronize()
bk = sh.buckets[8]
dense = bk.precs.cpu().numpy(); packed = bk.packed.cpu().numpy()
errs = [np.abs(dense[i] - op(packed[i], 2)).max() / np.abs(op(packed[i], 2)).max() for i in range(bk.count)]
print("dense vs op(packed): max rel", max(errs), "count", bk.count)
# preconditioned grad of param 0
plan = sh.plans[0]
pg = sh.pgbuf[plan.offset:plan.offset + plan.numel].cpu().numpy()
osh = oo.init.__self__
for pi in (1, 0):
  plan = sh.plans[pi]
  pg = sh.pgbuf[plan.offset:plan.offset + plan.numel].cpu().numpy()
  want = osh._preconditioner(params[pi]).preconditioned_grad(grads[5][pi], ost.stats[pi].preconditioners).reshape(-1)
  print("param", pi, "precond grad rel err", np.abs(pg - want).max() / np.abs(want).max())
  print("ours", pg[:24]); print("want", want[:24])
  print("upd ours", pu[pi].cpu().numpy().reshape(-1)[:8], "want", ou[pi].reshape(-1)[:8])
  print("mom ours", pst.stats[pi].momentum.to_float().cpu().numpy().reshape(-1)[:6], "want", ost.stats[pi].momentum.to_float().reshape(-1)[:6])
  print("dmom ours", pst.stats[pi].diagonal_momentum.to_float().cpu().numpy().reshape(-1)[:6], "want", ost.stats[pi].diagonal_momentum.to_float().reshape(-1)[:6])
want = want
print("precond grad rel err", np.abs(pg - want).max() / np.abs(want).max())
print("ours", pg[:16]); print("want", want[:16])
g = grads[5][0].reshape(-1)
print("grad", g[:16])
print("stat refs p0", plan.stat_refs[:4], "skip", plan.skip, "tshape", plan.tshape)
print("dense[0] @ g[:8]", dense[plan.stat_refs[0][1]].T @ g[:8])
