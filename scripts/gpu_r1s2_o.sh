timeout 600 python -m pytest tests/test_gpu_fd.py tests/test_gpu_optimizer.py -q -x 2>&1 | tail -2
timeout 300 python scripts/fd_bench.py 2>&1 | grep -E "step 2|check" | head -6
for mb in 0 64 96 112; do
  PC_PI_GROUP_MB=$mb timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-step 2>&1 | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); print('pi_group_mb=$mb value', round(l['value'],1), 'ms', round(l['ms_per_step'],2), 'gemm', round(l['roofline']['gemm_ms_per_step'],2))"
done
