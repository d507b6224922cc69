timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_optimizer.py tests/test_gpu_baseline_configs.py -q -x 2>&1 | tail -2
timeout 600 python scripts/resnet_step_probe.py 2>&1 | grep -v -i warn | grep -E "total|stats|apply|simt|tc_|splitk"
