"""GPU diagnostic: where does the end-to-end loop of bench.py lose time against the device-only loop?
Variants: solve only / + H2D / + D2H / both, same double-buffered structure."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from precondition_b200 import ops
dev = torch.device("cuda", 0)
B, n, p = 74, 1024, 4
xs = bench.make_statistics_torch(B, n, seed=1000, device=dev)
ps = torch.full((B,), p, dtype=torch.int32, device=dev)
ps_host = [p] * B
host_in = torch.empty((B, n, n), dtype=torch.float32).pin_memory(); host_in.copy_(xs.cpu())
host_out = torch.empty((B, n, n), dtype=torch.float32).pin_memory()
dev_ins = [xs.clone(), xs.clone()]
dev_outs = [torch.empty_like(xs), torch.empty_like(xs)]
met = [torch.empty((B, 5), device=dev) for _ in range(2)]
ws = torch.empty(ops.root_workspace_bytes(B, n) + 256, dtype=torch.uint8, device=dev)
cur = torch.cuda.current_stream(dev)
s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

def loop(steps, h2d, d2h):
  ev_in = [torch.cuda.Event() for _ in range(2)]; ev_solved = [torch.cuda.Event() for _ in range(2)]
  ev_out = [torch.cuda.Event() for _ in range(2)]
  s_h2d.wait_stream(cur); s_d2h.wait_stream(cur)
  def upload(i):
    k = i % 2
    with torch.cuda.stream(s_h2d):
      s_h2d.wait_event(ev_solved[k])
      if h2d: dev_ins[k].copy_(host_in, non_blocking=True)
      ev_in[k].record(s_h2d)
  upload(0)
  for i in range(steps):
    k = i % 2
    if i + 1 < steps: upload(i + 1)
    cur.wait_event(ev_in[k]); cur.wait_event(ev_out[k])
    ops.matrix_inverse_pth_root_batched(dev_ins[k], ps, None, out=dev_outs[k], metrics_out=met[k], workspace=ws, ps_host=ps_host)
    ev_solved[k].record(cur)
    with torch.cuda.stream(s_d2h):
      s_d2h.wait_event(ev_solved[k])
      if d2h: host_out.copy_(dev_outs[k], non_blocking=True)
      ev_out[k].record(s_d2h)
  cur.wait_stream(s_h2d); cur.wait_stream(s_d2h)

for name, h, d in (("solve only", False, False), ("+h2d", True, False), ("+d2h", False, True), ("both", True, True), ("solve only", False, False)):
  loop(3, h, d); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter(); e0.record(); loop(10, h, d); e1.record(); th = time.perf_counter() - t0
  torch.cuda.synchronize()
  print(f"{name:12s} {e0.elapsed_time(e1) / 10:7.2f} ms/step (host enqueue {th * 100:.2f} ms/step)")
# raw copy speeds
for name, f in (("h2d", lambda: dev_ins[0].copy_(host_in, non_blocking=True)), ("d2h", lambda: host_out.copy_(dev_outs[0], non_blocking=True))):
  f(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); f(); e1.record(); torch.cuda.synchronize()
  print(f"{name}: {e0.elapsed_time(e1):.2f} ms for {B * n * n * 4 / 1e6:.0f} MB = {B * n * n * 4 / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
