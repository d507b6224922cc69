mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/step_engine_probe.py 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_ggemm -c 2 -o gpurun_out/prof_ggemm -f python scripts/stats_timeline.py > gpurun_out/ncu_ggemm.log 2>&1
tail -2 gpurun_out/ncu_ggemm.log | cut -c1-200
