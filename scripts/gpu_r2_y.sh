#!/bin/bash
# factored low-rank application: optimizer / FD tests with it on and off, Sketchy step profile
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_optimizer.py tests/test_gpu_fd.py tests/test_gpu_baseline_configs.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -3
PC_LOWRANK_APPLY=0 timeout 900 python -m pytest tests/test_gpu_optimizer.py -q -m gpu -x -k "fd or sketchy or lowrank" 2>&1 | tail -2
timeout 300 python scripts/step_profile.py sketchy 2>&1 | grep -v Warn | head -9
PC_LOWRANK_APPLY=0 timeout 300 python scripts/step_profile.py sketchy 2>&1 | grep -v Warn | head -3
