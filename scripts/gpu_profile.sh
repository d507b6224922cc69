set -x
mkdir -p gpurun_out
B=${1:-16}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_phase -s 30 -c 4 -o gpurun_out/prof_tc -f python bench.py --steps 1 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/
