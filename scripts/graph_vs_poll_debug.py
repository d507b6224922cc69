"""GPU diagnostic: graph-mode vs host-polled solver, per-matrix differences."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_round2 import _spd_batch
from precondition_b200 import ops
for n in (256, 384):
  xs = torch.as_tensor(_spd_batch(n, 6, n)).cuda()
  xs[5] = 0.0
  ps = [4, 2, 6, 8, 4, 4]
  pads = [n, n, n - 3, n, 0, n]
  res = {}
  for mode in ("graph", "poll", "graph", "poll"):
    if mode == "poll": os.environ["PC_ROOT_MODE"] = "poll"
    else: os.environ.pop("PC_ROOT_MODE", None)
    r, m = ops.matrix_inverse_pth_root_batched(xs, ps, pads)
    torch.cuda.synchronize()
    if mode in res:
      a = res[mode]
      print(n, mode, "repeatable:", [bool(torch.equal(a[0][b].nan_to_num(0, 1, -1), r[b].nan_to_num(0, 1, -1))) for b in range(6)])
    res[mode] = (r.clone(), m.clone())
  os.environ.pop("PC_ROOT_MODE", None)
  for b in range(6):
    a, c = res["graph"][0][b], res["poll"][0][b]
    d = (a.nan_to_num(0, 1, -1) - c.nan_to_num(0, 1, -1)).abs().max()
    print(n, b, "maxdiff", float(d), "nan", int(a.isnan().sum()), int(c.isnan().sum()),
          "metrics", res["graph"][1][b].tolist(), res["poll"][1][b].tolist())
