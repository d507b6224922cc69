#!/bin/bash
# full GPU suite, then compute-sanitizer memcheck over the kernels added late in round 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | grep -v "Warning\|numerics.py\|^$\|nv = v\|v_out\|z = f\|mat_m =\|mat_h =\|h = conv" | tail -6
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tearfree.py -q -m gpu -x -k "golden or pinv_root or stand_alone or momentum" > gpurun_out/r2q_sanitizer_tearfree.log 2>&1
echo "sanitizer tearfree rc=$?"; tail -4 gpurun_out/r2q_sanitizer_tearfree.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_root.py -q -m gpu -x -k "golden or padding or power" > gpurun_out/r2q_sanitizer_root.log 2>&1
echo "sanitizer root rc=$?"; tail -4 gpurun_out/r2q_sanitizer_root.log
