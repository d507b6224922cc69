mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tc_phase|root_|power_iteration|simt|select|quant' -c 1500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-step > /dev/null 2>&1
tail -1 gpurun_out/launches_final.csv | cut -c1-160
