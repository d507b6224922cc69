mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "0"; do
  PC_TC_PAIR256=$cfg timeout 200 python bench.py --steps 3 --warmup 3 --batch 74 --no-cpu-baseline --no-step 2>&1 | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
print('pair256=$cfg ms_per_step', round(l['ms_per_step'],2), 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'iters', l['config']['newton_iters_mean'], 'roots/s', round(l['value'],1), 'frac', round(l['roofline']['frac'],3), 'maxerr', l['config']['max_error'], 'e2e', round(l['e2e']['ms_per_step'],2))"
done
timeout 300 python scripts/timeline.py 74 2>&1 | grep -E "span|power_iter|root_init|tc_phase|busy"
