#!/bin/bash
# 2-GPU: weak-scaling step with NCCL vs copy-engine peer gather, 1 / 2 / 4 sub-batches; then full bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # env... -- args
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-step --no-cpu-baseline $EXTRA > gpurun_out/r2d_$tag.json 2> gpurun_out/r2d_$tag.err
  python - <<PY
import json
try:
  d = json.loads(open("gpurun_out/r2d_$tag.json").read().strip().splitlines()[-1])
  print("$tag", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d["run_info"]["gather_overlap"][:60])
except Exception as e:
  print("$tag FAILED", e); print(open("gpurun_out/r2d_$tag.err").read()[-1500:])
PY
}
EXTRA="--split 1" run nccl_s1 PC_GATHER=nccl
EXTRA="--split 2" run nccl_s2 PC_GATHER=nccl
EXTRA="--split 1" run peer_s1 PC_GATHER=peer
EXTRA="--split 2" run peer_s2 PC_GATHER=peer
EXTRA="--split 4" run peer_s4 PC_GATHER=peer
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --no-big > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err
tail -3 gpurun_out/r2d_bench_n2.err
python - <<PY
import json
try:
  d = json.loads(open("gpurun_out/r2d_bench_n2.json").read().strip().splitlines()[-1])
  for k in ("value", "ms_per_step", "shampoo_step_resnet50", "run_info"):
    print(k, json.dumps(d.get(k))[:900])
except Exception as e:
  print("N=2 bench FAILED", e)
PY
