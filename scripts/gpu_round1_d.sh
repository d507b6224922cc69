set -x
mkdir -p gpurun_out
timeout 180 python scripts/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "tc_check rc=$?"; tail -16 gpurun_out/tc_check.log
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_root.py -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 2 --warmup 3 --batch 64 --engine tc6 --no-cpu-baseline > gpurun_out/bench_tc6_b64.json 2> gpurun_out/bench_tc6.err; cut -c1-1500 gpurun_out/bench_tc6_b64.json; tail -3 gpurun_out/bench_tc6.err
PC_TC_2CTA=0 timeout 300 python bench.py --steps 2 --warmup 3 --batch 64 --engine tc6 --no-cpu-baseline > gpurun_out/bench_tc6_b64_1cta.json 2>> gpurun_out/bench_tc6.err; cut -c1-400 gpurun_out/bench_tc6_b64_1cta.json
