set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python scripts/precision_probe.py 1024 > gpurun_out/precision_probe.jsonl 2> gpurun_out/precision_probe.err; tail -2 gpurun_out/precision_probe.err
timeout 600 python bench.py --steps 2 --warmup 3 --batch 16 > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; cat gpurun_out/bench_simt.json; tail -3 gpurun_out/bench_simt.err
timeout 300 python bench.py --steps 2 --warmup 3 --batch 64 --n 128 --no-cpu-baseline > gpurun_out/bench_n128.json 2>> gpurun_out/bench_simt.err; cat gpurun_out/bench_n128.json
