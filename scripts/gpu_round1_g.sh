mkdir -p gpurun_out
timeout 180 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; grep -E "gemm n=1024|gemm n=256 batch=3|root n=1024|root n=128" gpurun_out/tc_check.log | cut -c1-190
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_root.py -q 2>&1 | tail -3
for cfg in "0 1 32" "0 1 64" "0 0 32"; do
  set -- $cfg
  PC_TC_DEBUG=$1 PC_TC_WS=$2 timeout 200 python bench.py --steps 2 --warmup 3 --batch $3 --engine tc6 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
print('dbg=$1 ws=$2 batch=$3 ms_per_step', round(l['ms_per_step'],2), 'iters', l['config']['newton_iters_mean'], 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'roots/s', round(l['value'],1), 'e2e', round(l['e2e']['value'],1))"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_phase|root_|power_iteration|simt|select|quant' -c 2000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch 32 --engine tc6 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv
