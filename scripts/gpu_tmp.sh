timeout 900 python -m pytest tests/test_gpu_root.py tests/test_gpu_optimizer.py tests/test_gpu_baseline_configs.py -q -x 2>&1 | tail -2
timeout 300 python scripts/step_engine_probe.py 2>&1 | tail -4
