#!/bin/bash
# Round 2, GPU session A: smoke, GPU tests, default bench, kernel-variant benches, ResNet probe.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest new"; timeout 900 python -m pytest tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2a_pytest_new.log
echo "== pytest all"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r2a_pytest_all.log
echo "== bench default"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; tail -c 600 gpurun_out/r2a_bench_default.err
for v in "PC_ROOT_MODE=poll" "PC_TC_PAIR256=1" "PC_TC_PAIR256=1 PC_TC_CHUNK=2" "PC_TC_WS2=1" "PC_TC_CHUNK=2"; do
  tag=$(echo "$v" | tr ' =' '__')
  echo "== bench $v"
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-step --no-cpu-baseline > gpurun_out/r2a_bench_$tag.json 2> gpurun_out/r2a_bench_$tag.err
  python - <<PY
import json
try:
  d = json.loads(open("gpurun_out/r2a_bench_$tag.json").read().strip().splitlines()[-1])
  print("$v", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 4), "gemm_ms", round(d["roofline"]["gemm_ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "iters", d["run_info"]["newton_iters_mean"], "err", d["run_info"]["max_error"])
except Exception as e:
  print("$v", "FAILED", e)
PY
done
echo "== resnet probe"; timeout 600 python scripts/resnet_step_probe.py 2>&1 | tail -40 | tee gpurun_out/r2a_resnet_probe.log
python - <<PY
import json
d = json.loads(open("gpurun_out/r2a_bench_default.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "roofline", "parity", "shampoo_step", "shampoo_step_resnet50", "shampoo_step_bert_large", "sketchy_step", "sketchy_update", "cpu_baseline", "run_info", "gpu_launches", "clocks"):
  print(k, json.dumps(d.get(k))[:900])
PY
