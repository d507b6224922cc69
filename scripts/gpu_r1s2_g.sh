mkdir -p gpurun_out
PC_TC_PAIR256=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_phase -s 40 -c 3 -o gpurun_out/prof_fp16_ws -f python bench.py --steps 1 --warmup 3 --batch 74 --no-cpu-baseline --no-step > gpurun_out/ncu_full_ws.log 2>&1
tail -2 gpurun_out/ncu_full_ws.log | cut -c1-200
PC_TC_PAIR256=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_phase -s 40 -c 3 -o gpurun_out/prof_fp16_p256 -f python bench.py --steps 1 --warmup 3 --batch 74 --no-cpu-baseline --no-step > gpurun_out/ncu_full_p256.log 2>&1
tail -2 gpurun_out/ncu_full_p256.log | cut -c1-200
