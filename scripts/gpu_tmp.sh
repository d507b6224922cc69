timeout 600 python -m pytest tests/test_gpu_fd.py tests/test_gpu_optimizer.py tests/test_gpu_full_size.py tests/test_gpu_root.py -q -x 2>&1 | tail -2
timeout 300 python scripts/fd_bench.py 2>&1 | grep -E "step 2|check" | head -6
timeout 300 python scripts/fd_timeline.py 2>&1 | grep -v -i warn | tail -12
