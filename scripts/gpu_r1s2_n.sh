timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-step 2>&1 | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); print('value', round(l['value'],1), 'ms', round(l['ms_per_step'],2), 'e2e', round(l['e2e']['value'],1), 'frac', round(l['roofline']['frac'],3), 'iters', l['config']['newton_iters_mean'])"
timeout 300 python scripts/timeline.py 74 2>&1 | grep -E "span|power_iter|root_|tc_phase|busy"
