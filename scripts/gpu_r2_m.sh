#!/bin/bash
# power-iteration kernel alone under ncu (durations are shares, not bench values) + root tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_root.py tests/test_gpu_baseline_configs.py -q -m gpu -x 2>&1 | tail -2
PC_ROOT_MODE=poll timeout 600 ncu --metrics gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum --clock-control none -k regex:'power_iteration' -c 4 --csv --log-file gpurun_out/r2m_pi_ncu.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-step --no-big > /dev/null 2>&1
grep -v "^==" gpurun_out/r2m_pi_ncu.csv | python -c "
import csv, sys
for r in csv.DictReader(sys.stdin):
    print(' ', r['Kernel Name'][:44], r['Metric Name'], r['Metric Value'], r['Metric Unit'])
" | tail -3
