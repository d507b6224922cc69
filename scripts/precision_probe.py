"""GPU experiment (not product code): which split-precision scheme lets the
coupled Newton iteration reach the reference's 1e-6 stopping rule?

bf16 planes are held as fp32 tensors whose values are exactly bf16; a TF32
cuBLAS GEMM on them forms exact products and accumulates in the tensor core's
fp32 accumulator, i.e. the numerics of tcgen05 kind::f16 with fp32 accumulate.
Prints one JSON line per (matrix class, scheme, form).
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle.gen_golden import ema_statistics, gen_symmetric_matrix  # noqa: E402
from oracle import numerics as N  # noqa: E402


def split(x, k):
  planes, r = [], x
  for _ in range(k):
    b = r.to(torch.bfloat16).to(torch.float32)
    planes.append(b)
    r = r - b
  return planes


def mm_split(a, b, nplanes, nprod):
  """sum of the nprod largest plane products, smallest first."""
  pa, pb = split(a, nplanes), split(b, nplanes)
  pairs = sorted([(i + j, i, j) for i in range(nplanes) for j in range(nplanes)])[:nprod]
  torch.backends.cuda.matmul.allow_tf32 = True
  acc = None
  for _, i, j in reversed(pairs):
    t = pa[i] @ pb[j]
    acc = t if acc is None else acc + t
  torch.backends.cuda.matmul.allow_tf32 = False
  return acc


def mm_fp32(a, b):
  torch.backends.cuda.matmul.allow_tf32 = False
  return a @ b


def mm_tf32x3(a, b):
  def sp(x):
    hi = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)
    return hi, x - hi
  ah, al = sp(a)
  bh, bl = sp(b)
  torch.backends.cuda.matmul.allow_tf32 = True
  out = al @ bh + ah @ bl + ah @ bh
  torch.backends.cuda.matmul.allow_tf32 = False
  return out


SCHEMES = {
    "fp32": mm_fp32,
    "bf16x3": lambda a, b: mm_split(a, b, 2, 3),
    "bf16x6": lambda a, b: mm_split(a, b, 3, 6),
    "bf16x9": lambda a, b: mm_split(a, b, 3, 9),
    "tf32x3": mm_tf32x3,
}


def newton(a, p, mm, form, ridge=1e-6, tol=1e-6, iters=100):
  n = a.shape[0]
  eye = torch.eye(n, device=a.device)
  lam = torch.linalg.eigvalsh(a.double())[-1].float()
  damped = a + ridge * lam * eye
  z = (1 + p) / (2 * torch.linalg.norm(damped))
  m = damped * z
  h = eye * z**(1.0 / p)
  err = (m - eye).abs().max()
  trace = []
  if form == "standard":
    alpha = -1.0 / p
    i, ratio, old_h = 0, 1.0, h
    while i < iters and err > tol and ratio < 1.2:
      mi = (1 - alpha) * eye + alpha * m
      pw = mi
      for _ in range(int(np.log2(p))):
        pw = mm(pw, pw)
      new_m = mm(pw, m)
      new_h = mm(h, mi)
      new_err = (new_m - eye).abs().max()
      ratio = float(new_err / err)
      m, old_h, h, err = new_m, h, new_h, new_err
      i += 1
      trace.append(float(err))
    h = h if ratio < 1.2 else old_h
  else:
    d = (eye - m) / p
    i, ratio, old_h = 0, 1.0, h
    while i < iters and err > tol and ratio < 1.2:
      q = d
      for _ in range(int(np.log2(p))):
        q = 2 * q + mm(q, q)
      new_d = mm(q, d) - q / p + d
      new_h = h + mm(h, d)
      new_err = p * new_d.abs().max()
      ratio = float(new_err / err)
      d, old_h, h, err = new_d, h, new_h, new_err
      i += 1
      trace.append(float(err))
    h = h if ratio < 1.2 else old_h
  return h, i, float(err), ratio, trace, float(lam)


def main():
  rng = np.random.default_rng(0)
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
  mats = {
      "spectrum_1e4": gen_symmetric_matrix(rng, n, 1e4),
      "ema": ema_statistics(rng, n, 2 * n),
      "spectrum_1e6": gen_symmetric_matrix(rng, n, 1e6),
  }
  for name, a64 in mats.items():
    a = torch.as_tensor(a64.astype(np.float32)).cuda()
    for p in (4, 2):
      for scheme, mm in SCHEMES.items():
        for form in ("standard", "deviation"):
          h, it, err, ratio, trace, lam = newton(a, p, mm, form)
          eps = 1e-6 * lam
          exact = N.exact_inverse_pth_root(a64.astype(np.float32), p, eps)
          hn = h.cpu().numpy()
          rel = float(np.linalg.norm(hn - exact) / np.linalg.norm(exact))
          res = N.root_residual(hn, a64.astype(np.float32), p, eps)
          print(json.dumps(dict(matrix=name, n=n, p=p, scheme=scheme, form=form, iters=it,
                                err=err, ratio=ratio, rel_fro_vs_f64=rel, residual=res,
                                tail=trace[-3:])), flush=True)


if __name__ == "__main__":
  main()
