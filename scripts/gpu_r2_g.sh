#!/bin/bash
# 8-GPU: full bench (peer gather), then NCCL gather for comparison (headline only)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err
tail -3 gpurun_out/r2g_bench_n$N.err
PC_GATHER=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-step --no-cpu-baseline > gpurun_out/r2g_bench_n${N}_nccl.json 2> gpurun_out/r2g_bench_n${N}_nccl.err
python - <<PY
import json
for tag in ("", "_nccl"):
  try:
    d = json.loads(open("gpurun_out/r2g_bench_n$N%s.json" % tag).read().strip().splitlines()[-1])
    print(tag or "peer", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d["run_info"]["gather_overlap"][:70])
    if not tag:
      for k in ("shampoo_step_resnet50", "shampoo_step_bert_large", "sketchy_step"):
        print(k, json.dumps(d.get(k))[:1000])
  except Exception as e:
    print(tag, "FAILED", e)
PY
