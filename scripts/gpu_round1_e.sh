mkdir -p gpurun_out
timeout 180 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; tail -14 gpurun_out/tc_check.log | cut -c1-200
PC_TC_2CTA=0 timeout 180 python scripts/tc_check.py 2>&1 | grep -E "gemm n=1024|gemm n=256 batch=3|root n=1024" | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_root.py -q 2>&1 | tail -4
for cfg in "0 0" "0 1" "2 0" "2 1"; do
  set -- $cfg
  PC_TC_DEBUG=$1 PC_TC_2CTA=$2 timeout 200 python bench.py --steps 2 --warmup 3 --batch 32 --engine tc6 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
print('dbg=$1 2cta=$2 ms_per_step', round(l['ms_per_step'],2), 'iters', l['config']['newton_iters_mean'], 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'roots/s', round(l['value'],1))"
done
