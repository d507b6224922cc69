#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "eigh_roots" 2>&1 | grep -v "Warning\|numerics.py" | tail -25
timeout 300 python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
from precondition_b200 import ops
from oracle.gen_golden import ema_statistics
rng = np.random.default_rng(0)
for d, b in ((1024, 4), (2048, 2)):
  xs = torch.as_tensor(np.stack([ema_statistics(rng, d, 2 * d) for _ in range(b)]).astype(np.float32)).cuda()
  ops.matrix_inverse_pth_root_eigh_batched(xs, [4] * b); torch.cuda.synchronize()
  t0 = time.time(); r, m = ops.matrix_inverse_pth_root_eigh_batched(xs, [4] * b); torch.cuda.synchronize()
  print(f"eigh root d={d} batch={b}: {(time.time() - t0) * 1e3:.0f} ms, error metric {m[:, 0].tolist()}")
PY
