#!/bin/bash
# power-iteration shapes: tests, ResNet-50 step with shape A only vs automatic
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_root.py -q -m gpu -x -k "power_iteration or golden or mixed or graph_mode" 2>&1 | tail -3
for s in A auto A auto; do
  PC_PI_SHAPE=$s timeout 300 python - <<'PY'
import os, sys, torch
sys.path.insert(0, ".")
import bench
r = bench.time_resnet50_step(torch.device("cuda", 0), 1, steps=4, warm=3)
print("shape", os.environ.get("PC_PI_SHAPE"), "resnet50 step", round(r["ms"], 2), r["ms_per_step_list"])
PY
done
