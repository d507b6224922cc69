"""GPU diagnostic: times pc_fd_update_batched at Sketchy scale (BASELINE config 5:
4096 x 4096 blocks, rank 256) and checks the result against torch.linalg.eigh of the
same covariance (checker only)."""
import sys
import time
import numpy as np
import torch
sys.path.insert(0, ".")
from precondition_b200 import ops  # noqa: E402
from oracle import numerics as N  # noqa: E402


def case(d, rank, batch, iters, m=None):
  m = m or d
  g = torch.Generator(device="cuda").manual_seed(d + rank)
  # gradient blocks with a decaying spectrum + noise floor
  u = torch.linalg.qr(torch.randn(d, d, generator=g, device="cuda"))[0]
  spec = torch.cat([torch.logspace(0, -1.5, rank + 64, device="cuda"),
                    torch.full((d - rank - 64,), 0.01, device="cuda")])
  xs = torch.stack([(u * spec) @ torch.randn(d, m, generator=g, device="cuda") / m**0.5
                    for _ in range(batch)]).contiguous()
  prev = torch.zeros((batch, d, rank + 2), device="cuda")
  ps = [4] * batch
  out = prev
  for step in range(3):  # chained steps: the sketch warms up
    torch.cuda.synchronize(); t0 = time.time()
    out, _ = ops.fd_update_root_batched(xs, out, ps, rank, decay=0.999, subspace_iters=iters)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"d={d} rank={rank} batch={batch} m={m} iters={iters} step {step}: {dt * 1e3:.1f} ms", flush=True)
  # check step 3 against eigh of its covariance
  prev3 = out
  out, _ = ops.fd_update_root_batched(xs, prev3, ps, rank, decay=0.999, subspace_iters=iters)
  torch.cuda.synchronize()
  for b in range(min(batch, 1)):
    pk = prev3[b].double()
    vecs, lam, tail = pk[:, :rank], pk[-rank:, -1], pk[1, -1]
    ridge = 1e-6 * max(float(lam[0]), 1e-6)
    half = vecs * torch.sqrt(0.999 * (lam + ridge))
    c = half @ half.T + xs[b].double() @ xs[b].double().T
    s = torch.linalg.eigvalsh(c).flip(0)
    got = out[b].double()
    gv, ge, gt = got[:, :rank], got[-rank:, -1], got[1, -1]
    ev_err = float(((ge + s[rank]) - s[:rank]).abs().max() / s[0])
    tail_err = float(abs(gt - (0.999 * tail + s[rank])) / (0.999 * tail + s[rank]))
    orth = float((gv.T @ gv - torch.eye(rank, device="cuda", dtype=torch.float64)).abs().max())
    resid = float((c @ gv - gv * (ge + s[rank])).norm(dim=0).max() / s[0])
    print(f"   check: eig err {ev_err:.2e} tail err {tail_err:.2e} orth {orth:.2e} resid {resid:.2e} "
          f"has_zeros {float(got[-1, -2])}", flush=True)


if __name__ == "__main__":
  case(1024, 64, 4, 8)
  case(4096, 256, 2, 8)
  case(4096, 256, 2, 4)
