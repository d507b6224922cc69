mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in "0 auto"; do
  set -- $cfg
  PC_TC_PAIR256=$1 timeout 200 python bench.py --steps 3 --warmup 3 --batch 74 --engine $2 --no-cpu-baseline --no-step 2>&1 | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
its=l['config']['newton_iters_mean']
print('pair256=$1 engine=$2 ms_per_step', round(l['ms_per_step'],2), 'gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'iters', its, 'roots/s', round(l['value'],1), 'frac', round(l['roofline']['frac'],3), 'e2e', l['e2e']['ms_per_step'], 'launches', l['gpu_launches'])"
done
