mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/nccl_opt_check.py 2>&1 | grep -E "sharded ==|rank" | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_s2.json 2> gpurun_out/bench_n2_s2.err; echo "bench n2 rc=$?"
grep '^{' gpurun_out/bench_n2_s2.json | python -c "
import sys,json
l=json.loads(sys.stdin.readline()); print('n_gpus', l['n_gpus'], 'value', round(l['value'],1), 'e2e', round(l['e2e']['value'],1), 'frac', round(l['roofline']['frac'],3), l['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | grep '^{' | cut -c1-200
