"""GPU diagnostic: product optimizer vs oracle, state by state."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import optimizer as O
from oracle.gen_golden import OPT_CONFIGS, OPT_STEPS, opt_inputs
from precondition_b200 import distributed_shampoo as DS

names = sys.argv[1:] or ["adagrad_rank3", "adagrad_norm_output_wd", "sqrt_n_input", "quantized_int16"]
params, grads = opt_inputs()
for name in names:
  cfg = {k: v for k, v in OPT_CONFIGS[name].items() if not k.startswith("_")}
  okw, pkw = dict(cfg), dict(cfg)
  if "graft_type" in cfg:
    okw["graft_type"] = O.GraftingType(cfg["graft_type"]); pkw["graft_type"] = DS.GraftingType(cfg["graft_type"])
  if "precondtioner_type" in cfg:
    okw["precondtioner_type"] = O.PreconditionerType(cfg["precondtioner_type"])
    pkw["precondtioner_type"] = DS.PreconditionerType(cfg["precondtioner_type"])
  oo = O.distributed_shampoo(0.1, 8, batch_axis_name="batch", **okw)
  ost = oo.init(params)
  po = DS.distributed_shampoo(0.1, 8, batch_axis_name="batch", **pkw)
  tp = [torch.as_tensor(p).cuda() for p in params]
  pst = po.init(tp)
  print("==", name)
  for t in range(OPT_STEPS):
    with np.errstate(all="ignore"):
      ou, ost = oo.update(grads[t], ost, params)
    pu, pst = po.update([torch.as_tensor(g).cuda() for g in grads[t]], pst, tp)
    torch.cuda.synchronize()
    for i in range(len(params)):
      f = lambda x: x.to_float() if hasattr(x, "to_float") else x
      rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
      es = [rel(f(a).cpu().numpy(), f(b)) for a, b in zip(pst.stats[i].statistics, ost.stats[i].statistics)]
      ep = [rel(f(a).cpu().numpy(), f(b)) for a, b in zip(pst.stats[i].preconditioners, ost.stats[i].preconditioners)]
      em = rel(pst.stats[i].momentum.to_float().cpu().numpy(), ost.stats[i].momentum.to_float())
      eu = rel(pu[i].cpu().numpy(), ou[i])
      tm = pst.stats[i].training_metrics
      its = "" if tm is None or not len(es) else f" iters {tm[:,1].cpu().numpy().astype(int).tolist()} vs {ost.stats[i].training_metrics[:,1].astype(int).tolist()}"
      print(f" t={t} p{i}: upd {eu:.1e} mom {em:.1e} stat {max(es) if es else 0:.1e} prec {max(ep) if ep else 0:.1e} "
            f"worst_prec_idx {int(np.argmax(ep)) if ep else -1}{its if t in (0, 4, 5) else ''}", flush=True)
