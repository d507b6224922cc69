"""GPU diagnostic: Shampoo step (BASELINE config 2) with each root engine, and its phase split."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from precondition_b200 import distributed_shampoo as DS, _lib, ops

dev = torch.device("cuda", 0)
for engine in (0, 1, 4, 2):
  gen = torch.Generator(device=dev); gen.manual_seed(0)
  shapes = [(512, 2048), (2048,), (2048, 512), (512,)]
  params = [torch.randn(s, generator=gen, device=dev) * 0.05 for s in shapes]
  opt = DS.distributed_shampoo(0.1, 128, graft_type=DS.GraftingType.SGD, engine=engine)
  state = opt.init(params)
  grads = [[torch.randn(s, generator=gen, device=dev) * 1e-2 for s in shapes] for _ in range(16)]
  for t in range(6):
    _, state = opt.update(grads[t], state, params)
  torch.cuda.synchronize()
  t0 = time.time()
  for t in range(6, 16):
    upd, state = opt.update(grads[t], state, params)
  torch.cuda.synchronize()
  ms = (time.time() - t0) * 100
  sh = opt.init.__self__
  # phase split
  def timed(f, n=10):
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
  print(f"engine {engine}: step {ms:.3f} ms | stats {timed(sh._update_statistics):.3f} roots {timed(sh._compute_preconditioners):.3f} "
        f"apply {timed(sh._apply_preconditioners):.3f} ms", flush=True)
