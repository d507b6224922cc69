set -x
mkdir -p gpurun_out
timeout 240 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; tail -25 gpurun_out/tc_check.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 2 --warmup 3 --batch 16 --engine tc6 --no-cpu-baseline > gpurun_out/bench_tc6_b16.json 2> gpurun_out/bench_tc6.err; cat gpurun_out/bench_tc6_b16.json; tail -3 gpurun_out/bench_tc6.err
timeout 300 python bench.py --steps 2 --warmup 3 --batch 64 --engine tc6 --no-cpu-baseline > gpurun_out/bench_tc6_b64.json 2>> gpurun_out/bench_tc6.err; cat gpurun_out/bench_tc6_b64.json
