#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "== small tests"; timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "small_root" 2>&1 | tail -40 | tee gpurun_out/r2e_pytest_small.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-big > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 400 gpurun_out/r2e_bench.err
python - <<PY
import json
try:
  d = json.loads(open("gpurun_out/r2e_bench.json").read().strip().splitlines()[-1])
  for k in ("value", "small_block_roots", "shampoo_step", "shampoo_step_resnet50"):
    print(k, json.dumps(d.get(k))[:1200])
except Exception as e:
  print("bench FAILED", e)
PY
echo "== all tests"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r2e_pytest_all.log
