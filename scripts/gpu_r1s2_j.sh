mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_optimizer.py -q -x 2>&1 | tail -8
