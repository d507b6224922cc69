#!/bin/bash
# end-to-end loop with / without binding the process to the GPU's NUMA node
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -i "numa\|socket\|^CPU(s)" | head
for numa in 0 1 0 1; do
  PC_BENCH_NUMA=$numa timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-step --no-big > gpurun_out/r2p_numa$numa.json 2> gpurun_out/r2p_numa$numa.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2p_numa$numa.json").read().strip().splitlines()[-1])
print("numa=$numa", d["run_info"]["host_numa_binding"], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 2))
PY
done
