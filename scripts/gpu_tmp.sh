timeout 600 python bench.py --no-cpu-baseline 2>gpurun_out/b.err | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); print('value', round(l['value'],1)); print(l['shampoo_step_resnet50']['ms'], l['shampoo_step']['ms'], l['sketchy_update']['ms'])"
