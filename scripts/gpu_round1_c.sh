set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_optimizer.py > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python -m pytest tests/test_gpu_optimizer.py -q > gpurun_out/pytest_opt.log 2>&1; tail -40 gpurun_out/pytest_opt.log
