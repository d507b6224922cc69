mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fd.py -q 2>&1 | tail -30
timeout 600 python scripts/fd_bench.py 2>&1 | tail -20
