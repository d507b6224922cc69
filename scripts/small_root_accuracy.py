"""GPU diagnostic: float64 residual of the persistent small-block solver vs the fp32 oracle and the
CUDA-core engine (ratio ours / oracle per matrix), six- vs eight-term split products."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import numerics as N
from oracle.gen_golden import ema_statistics, gen_symmetric_matrix
from precondition_b200 import _lib, ops

def residual64(root, a, p, eps):
  x = torch.as_tensor(root).double().cuda()
  n = a.shape[0]
  d = torch.as_tensor(a).double().cuda() + eps * torch.eye(n, dtype=torch.float64, device="cuda")
  return float((torch.linalg.matrix_power(x, p) @ d - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max())

for n in (128, 64, 32):
  rng = np.random.default_rng(n)
  count = 16
  xs = np.stack([gen_symmetric_matrix(rng, n, 10.0 ** (1 + i % 4)) if i % 2 else
                 ema_statistics(rng, n, max(2 * n, 8)) for i in range(count)]).astype(np.float32)
  ps = [4, 2] * (count // 2)
  ref = []
  for b in range(count):
    want, wm = N.matrix_inverse_pth_root(xs[b], ps[b])
    ref.append((residual64(want, xs[b], ps[b], 1e-6 * wm.max_eigen_value), wm))
  for name, env, eng in (("small6", {}, _lib.PC_ENGINE_TC_SMALL), ("small8", {"PC_SMALL_TERMS": "8"}, _lib.PC_ENGINE_TC_SMALL),
                         ("simt", {}, _lib.PC_ENGINE_SIMT_FP32)):
    os.environ.pop("PC_SMALL_TERMS", None); os.environ.update(env)
    r, m = ops.matrix_inverse_pth_root_batched(torch.as_tensor(xs).cuda(), ps, engine=eng)
    torch.cuda.synchronize()
    ratios = []
    for b in range(count):
      ours = residual64(r[b].cpu().numpy(), xs[b], ps[b], 1e-6 * ref[b][1].max_eigen_value)
      ratios.append(ours / max(ref[b][0], 1e-12))
    its = [int(m[b, 1]) - int(ref[b][1].inverse_pth_root_iters) for b in range(count)]
    print(n, name, "residual ratio ours/oracle: max %.2f median %.2f" % (max(ratios), float(np.median(ratios))),
          "iter diffs", its)
