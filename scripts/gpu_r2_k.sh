#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_baseline_configs.py tests/test_gpu_optimizer.py -q -m gpu -x 2>&1 | grep -v "Warning\|numerics.py\|^$\|nv = v\|v_out\|z = f\|mat_m =\|mat_h =\|h = conv" | tail -8
timeout 600 python - <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
import bench
r = bench.time_bert_large_step(torch.device("cuda", 0), 1)
print({k: r[k] for k in ("ms", "ms_per_step_list", "sampled_root_rel_frobenius_vs_oracle")})
PY
