mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/nccl_opt_check.py 2>&1 | grep -v -i "warn\|NCCL version\|^$" | tail -8; echo "nccl_opt_check rc=${PIPESTATUS[0]}"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_s2.json 2> gpurun_out/bench_n2_s2.err; echo "bench n2 rc=$?"; grep '^{' gpurun_out/bench_n2_s2.json | cut -c1-1200
