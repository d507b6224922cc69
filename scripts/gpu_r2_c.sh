#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/graph_vs_poll_debug.py 2>&1 | tail -40 | tee gpurun_out/r2c_graph_vs_poll.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
tail -5 gpurun_out/r2c_bench_n2.err
python - <<PY
import json
try:
  d = json.loads(open("gpurun_out/r2c_bench_n2.json").read().strip().splitlines()[-1])
  for k in ("value", "ms_per_step", "e2e", "shampoo_step_resnet50", "shampoo_step_bert_large", "sketchy_step", "run_info", "gpu_launches"):
    print(k, json.dumps(d.get(k))[:1200])
except Exception as e:
  print("N=2 bench FAILED", e)
PY
