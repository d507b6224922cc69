"""GPU diagnostic: one Newton-chain GEMM phase (74 x 1024^2, scaled-fp16 engine) through the
debug hook, timed with CUDA events; run under PC_TC_ABLATE=0/1/2/3 to see what the phase costs
without the epilogue stores (bit 0) and without the TMEM chunk pull (bit 1)."""
import sys, torch
sys.path.insert(0, ".")
from precondition_b200 import ops
b, n = 74, 1024
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn((b, n, n), generator=g, device="cuda")
a = (a + a.transpose(1, 2)).contiguous()
for _ in range(2):
  ops.debug_tc_gemm(a, a, -3)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for _ in range(3):
    ops.debug_tc_gemm(a, a, -3)
  torch.cuda.synchronize()
ts = [e.time_range.end - e.time_range.start for e in prof.events()
      if e.device_type == torch.autograd.DeviceType.CUDA and "tc_phase" in e.name]
print("tc_phase launches", len(ts), "us each", [round(t, 1) for t in ts])
