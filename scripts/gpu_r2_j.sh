#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2j_bench_n$N.json 2> gpurun_out/r2j_bench_n$N.err
tail -3 gpurun_out/r2j_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
python - <<PY
import json
try:
  d = json.loads(open("gpurun_out/r2j_bench_n$N.json").read().strip().splitlines()[-1])
  print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d["run_info"]["gather_overlap"][:90], "launches", d["gpu_launches"])
  for k in ("shampoo_step_resnet50", "shampoo_step_bert_large", "sketchy_step"):
    v = d.get(k) or {}
    print(k, v.get("ms"), v.get("sharded_vs_single_max_rel"))
except Exception as e:
  print("FAILED", e)
PY
