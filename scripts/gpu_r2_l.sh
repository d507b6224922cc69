#!/bin/bash
# symmetric-half power iteration: parity tests, A/B of the bench call, launch list + L2 hit rate
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_root.py tests/test_gpu_full_size.py tests/test_gpu_baseline_configs.py tests/test_gpu_round2.py -q -m gpu -x 2>&1 | grep -v "Warning\|numerics.py\|^$\|nv = v\|v_out\|z = f\|mat_m =\|mat_h =\|h = conv" | tail -8
for rep in 1 2; do
for sym in 0 1; do
  PC_PI_SYM=$sym timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-step --no-big > gpurun_out/r2l_sym$sym.json 2> gpurun_out/r2l_sym$sym.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2l_sym$sym.json").read().strip().splitlines()[-1])
print("sym=$sym", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "gemm_share", round(d["roofline"]["gemm_share_of_step"], 3), "iters", d["run_info"]["newton_iters_mean"])
PY
done
done
PC_ROOT_MODE=poll timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum --clock-control none -k regex:'power_iteration' -c 6 --csv --log-file gpurun_out/r2l_pi_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-step --no-big > /dev/null 2>&1
grep -v "^==" gpurun_out/r2l_pi_ncu.csv | python -c "
import csv, sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:40], r['Metric Name'], r['Metric Value'], r['Metric Unit'])
" | tail -9
