#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tail -12
import sys; sys.path.insert(0, "scripts"); sys.path.insert(0, ".")
import fd_bench
fd_bench.case(4096, 256, 2, 6)
fd_bench.case(1024, 64, 4, 6)
PY
timeout 600 python scripts/fd_timeline.py 2>&1 | tail -16
echo "== fd tests"; timeout 900 python -m pytest tests/test_gpu_fd.py tests/test_gpu_full_size.py tests/test_gpu_optimizer.py -q -m gpu 2>&1 | tail -8
