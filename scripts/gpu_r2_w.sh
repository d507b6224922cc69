#!/bin/bash
# tile-based first initialisation: parity tests, A/B against the strip kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_root.py tests/test_gpu_full_size.py tests/test_gpu_baseline_configs.py tests/test_gpu_round2.py -q -m gpu -x 2>&1 | tail -3
for rep in 1 2; do
for s in 1 0; do
  PC_INIT_STRIP=$s timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-step --no-big > gpurun_out/r2w_$s.json 2>/dev/null
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2w_$s.json").read().strip().splitlines()[-1])
print("strip=$s", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "iters", d["run_info"]["newton_iters_mean"], "maxerr", d["run_info"]["max_error"])
PY
done
done
