mkdir -p gpurun_out
timeout 180 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; grep -E "gemm|root n=1024|root n=128" gpurun_out/tc_check.log | cut -c1-190
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_root.py -q 2>&1 | tail -4
for cfg in "0 1" "0 0" "2 1" "1 1"; do
  set -- $cfg
  PC_TC_DEBUG=$1 PC_TC_WS=$2 timeout 200 python bench.py --steps 2 --warmup 3 --batch 32 --engine tc6 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
print('dbg=$1 ws=$2 ms_per_step', round(l['ms_per_step'],2), 'iters', l['config']['newton_iters_mean'], 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'roots/s', round(l['value'],1))"
done
