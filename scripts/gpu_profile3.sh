mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_phase -s 40 -c 3 -o gpurun_out/prof_p256 -f python bench.py --steps 1 --warmup 3 --batch 32 --engine tc6 --no-cpu-baseline --no-step > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
