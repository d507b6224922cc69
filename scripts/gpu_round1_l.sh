mkdir -p gpurun_out
PC_TC_PAIR128=1 timeout 180 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; grep -E "gemm n=1024|gemm n=256 batch=3|root n=1024" gpurun_out/tc_check.log | cut -c1-190
PC_TC_PAIR128=1 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_root.py -q 2>&1 | tail -2
for cfg in "1 74" "0 74" "1 37"; do
  set -- $cfg
  PC_TC_PAIR128=$1 timeout 200 python bench.py --steps 2 --warmup 3 --batch $2 --engine tc6 --no-cpu-baseline --no-step 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
its=l['config']['newton_iters_mean']; n_it = its if its<100 else 600
print('pair128=$1 batch=$2 ms_per_step', round(l['ms_per_step'],2), 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'iters', its, 'ms_per_iter', round(l['roofline']['gemm_ms_per_step']/n_it,3), 'roots/s', round(l['value'],1), 'frac', round(l['roofline']['frac'],3))"
done
