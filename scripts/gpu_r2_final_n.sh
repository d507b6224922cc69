#!/bin/bash
# final multi-GPU bench line of round 2: N=${N:-8} ranks on one node
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2k_bench_n$N.json 2> gpurun_out/r2k_bench_n$N.err
tail -3 gpurun_out/r2k_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
  d = json.loads(open("gpurun_out/r2k_bench_n$N.json").read().strip().splitlines()[-1])
  print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
  for k in ("shampoo_step_resnet50", "shampoo_step_bert_large", "sketchy_step"):
    v = d.get(k) or {}
    print(k, v.get("ms"), v.get("sharded_vs_single_max_rel"), (v.get("shard_optimizer_states") or {}).get("ms"))
except Exception as e:
  print("FAILED", e)
PY
