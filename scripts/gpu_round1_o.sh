mkdir -p gpurun_out
timeout 180 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; grep -E "gemm n=1024|gemm n=256 batch=3|root n=1024" gpurun_out/tc_check.log | cut -c1-190
PC_TC_PAIR256=0 timeout 180 python scripts/tc_check.py 2>&1 | grep -E "gemm n=1024|gemm n=256 batch=3|gemm n=512|root n=1024" | cut -c1-150
for cfg in "1 1 74" "1 0 74" "0 1 74" "0 0 74" "0 1 72" "1 1 70"; do
  set -- $cfg
  PC_TC_PAIR256=$1 PC_TC_SYNC=$2 timeout 200 python bench.py --steps 2 --warmup 3 --batch $3 --engine tc6 --no-cpu-baseline --no-step 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
its=l['config']['newton_iters_mean']; n_it = its if its<100 else 600
print('pair256=$1 sync=$2 batch=$3 ms_per_step', round(l['ms_per_step'],2), 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'iters', its, 'ms_per_iter', round(l['roofline']['gemm_ms_per_step']/n_it,3), 'roots/s', round(l['value'],1), 'frac', round(l['roofline']['frac'],3))"
done
PC_TC_PAIR256=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_phase -s 40 -c 2 -o gpurun_out/prof_ws_sync -f python bench.py --steps 1 --warmup 3 --batch 74 --engine tc6 --no-cpu-baseline --no-step > gpurun_out/ncu_full_ws.log 2>&1
tail -2 gpurun_out/ncu_full_ws.log | cut -c1-200
