mkdir -p gpurun_out
for cfg in "0 0" "1 0" "2 0" "0 1" "1 1" "2 1"; do
  set -- $cfg
  PC_TC_DEBUG=$1 PC_TC_2CTA=$2 timeout 200 python bench.py --steps 2 --warmup 3 --batch 32 --engine tc6 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
print('dbg=$1 2cta=$2 ms_per_step', round(l['ms_per_step'],2), 'iters', l['config']['newton_iters_mean'], 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2))"
done
