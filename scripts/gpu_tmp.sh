timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python scripts/step_engine_probe.py 2>&1 | tail -4
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-step 2>&1 | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); print('value', round(l['value'],1), 'ms', round(l['ms_per_step'],2), 'gemm', round(l['roofline']['gemm_ms_per_step'],2), 'frac', round(l['roofline']['frac'],3), 'launches', l['gpu_launches'])"
