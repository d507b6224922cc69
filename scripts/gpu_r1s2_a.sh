mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_default.json
for cfg in "1 74" "0 74"; do
  set -- $cfg
  PC_TC_PAIR256=$1 timeout 200 python bench.py --steps 2 --warmup 3 --batch $2 --engine tc6 --no-cpu-baseline --no-step 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
its=l['config']['newton_iters_mean']
print('pair256=$1 batch=$2 ms_per_step', round(l['ms_per_step'],2), 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'iters', its, 'roots/s', round(l['value'],1), 'frac', round(l['roofline']['frac'],3))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_phase -s 40 -c 2 -o gpurun_out/prof_p256 -f python bench.py --steps 1 --warmup 3 --batch 74 --engine tc6 --no-cpu-baseline --no-step > gpurun_out/ncu_full_p256.log 2>&1
tail -2 gpurun_out/ncu_full_p256.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-step > /dev/null 2>&1
tail -1 gpurun_out/launches_default.csv | cut -c1-200
