N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}_s2.json 2> gpurun_out/bench_n${N}_s2.err; echo "bench n$N rc=$?"
grep '^{' gpurun_out/bench_n${N}_s2.json | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); print('n_gpus', l['n_gpus'], 'value', round(l['value'],1), 'ms', round(l['ms_per_step'],2), 'e2e', round(l['e2e']['value'],1), 'frac', round(l['roofline']['frac'],3), l['clocks'])"
tail -3 gpurun_out/bench_n${N}_s2.err | cut -c1-300
