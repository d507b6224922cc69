mkdir -p gpurun_out
timeout 300 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; grep -E "passes=-3|engine=4" gpurun_out/tc_check.log | cut -c1-230
PC_TC_PAIR256=0 timeout 300 python scripts/tc_check.py 2>&1 | grep -E "passes=-3|engine=4" | cut -c1-230
for cfg in "1 2" "0 2" "1 1" "0 1"; do
  set -- $cfg
  PC_TC_PAIR256=$1 PC_TC_CHUNK=$2 timeout 200 python bench.py --steps 3 --warmup 3 --batch 74 --no-cpu-baseline --no-step 2>&1 | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
its=l['config']['newton_iters_mean']
print('pair256=$1 chunk=$2 ms_per_step', round(l['ms_per_step'],2), 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'iters', its, 'roots/s', round(l['value'],1), 'frac', round(l['roofline']['frac'],3), 'maxerr', l['config']['max_error'], 'clk', l['clocks']['sm_mhz'], l['clocks']['reasons'])"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
