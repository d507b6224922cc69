mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_root.py -q 2>&1 | tail -3
for cfg in "1 32" "1 64" "0 64"; do
  set -- $cfg
  PC_TC_WS=$1 timeout 200 python bench.py --steps 2 --warmup 3 --batch $2 --engine tc6 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
print('ws=$1 batch=$2 ms_per_step', round(l['ms_per_step'],2), 'iters', l['config']['newton_iters_mean'], 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'roots/s', round(l['value'],1), 'e2e', round(l['e2e']['value'],1))"
done
timeout 300 python bench.py --steps 2 --warmup 3 --batch 64 --n 128 --no-cpu-baseline | cut -c1-260
