"""CPU emulation of split-precision schemes for the coupled Newton iteration
(DS:836-885): which operand splits keep the reference's iteration counts and
residuals?  Products of planes are formed exactly (float64), rounded to fp32
per K-chunk of 64 and chunk sums are accumulated in fp32, like the tcgen05
engine does.  Usage: python scripts/split_probe_cpu.py [n]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import numerics as N
from oracle.gen_golden import gen_symmetric_matrix, ema_statistics

CH = int(__import__("os").environ.get("PROBE_CHUNK", "64"))

def bf16(x):
  x = np.asarray(x, np.float32)
  b = x.view(np.uint32).astype(np.uint64)
  r = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16) << 16
  return r.astype(np.uint32).view(np.float32)

def split_bf16(x, k):
  planes, r = [], x.astype(np.float32)
  for _ in range(k):
    p = bf16(r); planes.append(p); r = (r - p).astype(np.float32)
  return planes

def split_fp16(x, k, shift=11):
  planes, r, sc = [], x.astype(np.float32), 1.0
  for i in range(k):
    p = (r * np.float32(sc)).astype(np.float16).astype(np.float32)
    planes.append((p, sc))
    r = (r - p / np.float32(sc)).astype(np.float32)
    sc *= 2.0**shift
  return planes

def chunked(pairs, n):
  """pairs: list of (A_plane64, B_plane64, scale). returns fp32 product with chunked accumulation."""
  out = np.zeros((n, n), np.float32)
  for k0 in range(0, n, CH):
    acc = np.zeros((n, n), np.float64)
    for a, b, s in pairs:
      acc += (a[:, k0:k0 + CH] @ b[k0:k0 + CH, :]) * s
    out = (out + acc.astype(np.float32)).astype(np.float32)
  return out

def mm(scheme):
  def f(a, b):
    n = a.shape[0]
    if scheme == "fp32":
      return a @ b
    if scheme.startswith("bf16x"):
      terms = int(scheme[5:]); k = 3
      A = [p.astype(np.float64) for p in split_bf16(a, k)]
      B = [p.astype(np.float64) for p in split_bf16(b, k)]
      lim = {3: 1, 6: 2, 9: 4}[terms]
      pairs = [(A[i], B[j], 1.0) for i in range(k) for j in range(k) if i + j <= lim]
    else:  # fp16xT
      terms = int(scheme[5:]); k = 2 if terms <= 4 else 3
      A = split_fp16(a, k); B = split_fp16(b, k)
      sel = {3: [(0, 0), (0, 1), (1, 0)], 4: [(0, 0), (0, 1), (1, 0), (1, 1)],
             6: [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]}[terms]
      pairs = [(A[i][0].astype(np.float64), B[j][0].astype(np.float64), 1.0 / (A[i][1] * B[j][1]))
               for i, j in sel]
    c = chunked(pairs, n)
    c = np.tril(c) + np.tril(c, -1).T  # engine: lower triangle authoritative
    return c.astype(np.float32)
  return f

def stored(scheme, x):
  """value actually kept between GEMMs (operand planes re-summed)"""
  if scheme == "fp32": return x
  if scheme.startswith("bf16x"): return x  # 3 bf16 planes are exact
  k = 2 if int(scheme[5:]) <= 4 else 3
  return sum((p / np.float32(s)).astype(np.float64) for p, s in split_fp16(x, k)).astype(np.float32)

def root(a, p, scheme, ridge_eps=1e-6):
  f = np.float32; n = a.shape[0]; a = a.astype(f); I = np.eye(n, dtype=f)
  _, ev = N.power_iteration(a, 100, 1e-6, None)
  ridge = f(ridge_eps) * max(ev, f(1e-25)); alpha = f(-1.0 / p)
  d = a + ridge * I
  z = f(1 + p) / (f(2) * np.linalg.norm(d).astype(f))
  M = d * z; H = I * np.power(z, f(1.0 / p)); err = np.max(np.abs(M - I)); ratio = f(1); i = 0
  g = mm(scheme); oldH = H
  while i < 100 and err > 1e-6 and ratio < 1.2:
    Mi = ((f(1) - alpha) * I + alpha * M).astype(f)
    if p == 4:
      q = g(Mi, Mi); q = g(q, q); newM = g(q, M)
    elif p == 2:
      q = g(Mi, Mi); newM = g(q, M)
    newH = g(H, Mi)
    nerr = np.max(np.abs(newM - I)); ratio = nerr / err
    M, oldH, H, err = newM, H, newH, nerr; i += 1
  Hout = H if ratio < 1.2 else oldH
  return Hout, i, float(err), float(ridge)

def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
  rng = np.random.default_rng(0)
  mats = {"spec1e2": gen_symmetric_matrix(rng, n, 1e2), "spec1e4": gen_symmetric_matrix(rng, n, 1e4),
          "spec1e6": gen_symmetric_matrix(rng, n, 1e6), "ema": ema_statistics(rng, n, n // 4),
          "ema_full": ema_statistics(rng, n, 2 * n)}
  schemes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp32", "bf16x6", "fp16x3", "fp16x4", "fp16x6"]
  print("| matrix | p | scheme | iters | err | relF vs f64 | resid |")
  for name, a in mats.items():
    for p in (4, 2):
      for s in schemes:
        H, it, err, ridge = root(a, p, s)
        a64 = a.astype(np.float32).astype(np.float64)
        exact = N.exact_inverse_pth_root(a64, p, ridge)
        rel = np.linalg.norm(H - exact) / np.linalg.norm(exact)
        res = N.root_residual(H, a64, p, ridge)
        print(f"| {name} | {p} | {s} | {it} | {err:.2e} | {rel:.2e} | {res:.2e} |", flush=True)

main()
