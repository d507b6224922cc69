"""GPU diagnostic: interleaved A/B timing of the Newton-chain kernel variants in ONE process
(same thermal / power state): whole solver call (graph mode) and GEMM-launch time (host-polled
mode with CUDA events around the GEMM launches)."""
import ctypes, os, sys
import torch
sys.path.insert(0, ".")
import bench
from precondition_b200 import _lib, ops

dev = torch.device("cuda", 0)
B, n, p = int(os.environ.get("B", 74)), 1024, 4
xs = bench.make_statistics_torch(B, n, seed=1000, device=dev)
ps = torch.full((B,), p, dtype=torch.int32, device=dev)
roots = torch.empty_like(xs)
met = torch.empty((B, 5), device=dev)
lib = _lib.load()
variants = {"ws": {}, "ws2": {"PC_TC_WS2": "1"}, "pair256": {"PC_TC_PAIR256": "1"},
            "pair256_chunk2": {"PC_TC_PAIR256": "1", "PC_TC_CHUNK": "2"},
            "ws_chunk2": {"PC_TC_CHUNK": "2"}}
names = sys.argv[1:] or list(variants)
KEYS = ("PC_TC_WS2", "PC_TC_PAIR256", "PC_TC_CHUNK")
def run(env):
  for k in KEYS: os.environ.pop(k, None)
  os.environ.update(env)
  ops.matrix_inverse_pth_root_batched(xs, ps, None, out=roots, metrics_out=met, ps_host=[p] * B)
for nm in names:  # warm every variant
  run(variants[nm]); run(variants[nm])
torch.cuda.synchronize()
res = {nm: {"call": [], "gemm": []} for nm in names}
st = _lib.Stats()
for rep in range(int(os.environ.get("REPS", 4))):
  for nm in names:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run(variants[nm])
    e1.record(); torch.cuda.synchronize()
    res[nm]["call"].append(e0.elapsed_time(e1) / 3)
    lib.pc_stats_reset(1)
    run(variants[nm]); torch.cuda.synchronize()
    lib.pc_stats_get(ctypes.byref(st)); lib.pc_stats_reset(0)
    res[nm]["gemm"].append(st.gemm_ms)
for nm in names:
  c, g = res[nm]["call"], res[nm]["gemm"]
  print(f"{nm:16s} call ms {' '.join(f'{x:6.2f}' for x in c)} | gemm ms {' '.join(f'{x:6.2f}' for x in g)} | "
        f"min call {min(c):.2f} min gemm {min(g):.2f} roots/s {B / min(c) * 1e3:.0f} iters {float(met[:,1].mean()):.1f} err {float(met[:,0].max()):.2e}")
