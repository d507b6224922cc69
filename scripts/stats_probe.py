"""GPU diagnostic: statistics update / preconditioner application time, fp32 CUDA-core
grouped GEMM vs the tcgen05 grouped GEMM, on block_size-1024 and -2048 parameters."""
import sys, time
import torch
sys.path.insert(0, ".")
from precondition_b200 import distributed_shampoo as DS, _lib

dev = torch.device("cuda", 0)
def timed(f, n=5):
  f(); torch.cuda.synchronize(); t0 = time.time()
  for _ in range(n): f()
  torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
for shape, block in [((4096, 4096), 1024), ((4096, 4096), 2048), ((1024, 4096), 1024)]:
  for engine, name in ((_lib.PC_ENGINE_SIMT_FP32, "simt"), (_lib.PC_ENGINE_AUTO, "tcgen05")):
    p = [torch.randn(shape, device=dev) * 0.05]
    opt = DS.distributed_shampoo(0.1, block, engine=engine, preconditioning_compute_steps=1000)
    st = opt.init(p)
    sh = opt.init.__self__
    sh.gbuf.copy_(torch.randn_like(sh.gbuf) * 1e-2)
    nb = (shape[0] // block) * (shape[1] // block)
    gf = nb * 2 * (2 * block**3) / 1e9
    ts, ta = timed(sh._update_statistics), timed(sh._apply_preconditioners)
    print(f"{shape} block {block} {name:8s}: stats {ts:8.3f} ms ({gf / ts:7.1f} TFLOP/s)  apply {ta:8.3f} ms ({gf / ta:7.1f} TFLOP/s)", flush=True)
