mkdir -p gpurun_out
PC_TC_PAIR256=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_phase|root_|power_iteration|simt|select|quant' -c 1200 --csv --log-file gpurun_out/launches_fp16x3.csv python bench.py --steps 1 --warmup 3 --batch 74 --no-cpu-baseline --no-step > /dev/null 2>&1
tail -1 gpurun_out/launches_fp16x3.csv | cut -c1-200
