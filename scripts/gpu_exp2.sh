for cfg in 0 16 32 2 1; do
  PC_TC_DEBUG=$cfg timeout 200 python bench.py --steps 2 --warmup 3 --batch 32 --engine tc6 --no-cpu-baseline --no-step 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readline())
its=l['config']['newton_iters_mean']; n_it = its if its<100 else 600
print('dbg=$cfg gemm_ms', round(l['roofline']['gemm_ms_per_step'],2), 'iters', its, 'ms_per_iter', round(l['roofline']['gemm_ms_per_step']/n_it,3))"
done
