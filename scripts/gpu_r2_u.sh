#!/bin/bash
# FD rework check: tests, accuracy/timing of the 4096 / rank 256 update with the old and new orthonormalisation, step profile
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_fd.py tests/test_gpu_full_size.py tests/test_gpu_tearfree.py -q -m gpu -x 2>&1 | tail -3
echo "--- new (inverse + GEMM)"; timeout 300 python scripts/fd_bench.py 2>&1 | tail -4
echo "--- old (PC_FD_TRSM=1)"; PC_FD_TRSM=1 timeout 300 python scripts/fd_bench.py 2>&1 | tail -4
timeout 300 python scripts/step_profile.py sketchy 2>&1 | grep -v Warn | head -16
