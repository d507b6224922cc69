"""GPU diagnostic: how much does a concurrent H2D / D2H / D2D copy stream slow the solver, and
which part of it (GEMM launches timed by the library vs the rest)?"""
import ctypes, sys, threading, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from precondition_b200 import _lib, ops
dev = torch.device("cuda", 0)
B, n, p = 74, 1024, 4
xs = bench.make_statistics_torch(B, n, seed=1000, device=dev)
ps = torch.full((B,), p, dtype=torch.int32, device=dev)
ps_host = [p] * B
host = torch.empty((B, n, n), dtype=torch.float32).pin_memory()
buf = torch.empty_like(xs); buf2 = torch.empty_like(xs)
out = torch.empty_like(xs); met = torch.empty((B, 5), device=dev)
ws = torch.empty(ops.root_workspace_bytes(B, n) + 256, dtype=torch.uint8, device=dev)
side = torch.cuda.Stream(dev)
lib = _lib.load()

def solve():
  ops.matrix_inverse_pth_root_batched(xs, ps, None, out=out, metrics_out=met, workspace=ws, ps_host=ps_host)

def run(kind, copies_per_step):
  solve(); torch.cuda.synchronize()
  stats = _lib.Stats()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(5):
    with torch.cuda.stream(side):
      for _ in range(copies_per_step):
        if kind == "h2d": buf.copy_(host, non_blocking=True)
        elif kind == "d2h": host.copy_(buf, non_blocking=True)
        elif kind == "d2d": buf2.copy_(buf, non_blocking=True)
    solve()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / 5

for kind, c in (("none", 0), ("h2d", 1), ("h2d", 3), ("h2d", 6), ("d2h", 1), ("d2h", 5), ("d2d", 20), ("none", 0)):
  print(f"{kind:5s} x{c}: {run(kind, c):7.2f} ms/step", flush=True)
