#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/graft_group_debug.py 2>&1 | tail -50 | tee gpurun_out/r2b_graft_debug.log
timeout 600 python scripts/variant_ab_probe.py 2>&1 | tail -12 | tee gpurun_out/r2b_variant_ab.log
timeout 600 python -m pytest tests/test_gpu_round2.py -q -m gpu 2>&1 | tail -40 | tee gpurun_out/r2b_pytest_new.log
