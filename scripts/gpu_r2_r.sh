#!/bin/bash
# racecheck / synccheck of the solver kernels on a small problem (n = 256: symmetric power iteration
# with clusters + tcgen05 Newton chain), and of the tearfree tail
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/rc_probe.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from precondition_b200 import ops
from oracle.gen_golden import gen_symmetric_matrix
rng = np.random.default_rng(0)
xs = torch.as_tensor(np.stack([gen_symmetric_matrix(rng, 256, 1e2) for _ in range(3)]).astype(np.float32)).cuda()
r, m = ops.matrix_inverse_pth_root_batched(xs, [4, 2, 4], [256, 250, 256])
torch.cuda.synchronize()
print("iters", m[:, 1].tolist(), "err", m[:, 0].tolist())
PY
for tool in racecheck synccheck; do
  PC_ROOT_MODE=poll timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/rc_probe.py > gpurun_out/r2r_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "iters|ERROR SUMMARY|hazard|Error" gpurun_out/r2r_$tool.log | head -8
done
