#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/small_root_prof.py 2>&1 | grep -v Warning | tail -4
timeout 600 python scripts/small_root_accuracy.py 2>&1 | grep -v Warning | grep -v small8 | tail -8
timeout 900 python -m pytest tests/test_gpu_round2.py -q -m gpu -k "small_root" 2>&1 | grep -v "Warning\|numerics.py\|^$\|nv = v\|v_out\|z = f\|mat_m =\|mat_h =\|h = conv" | tail -12
