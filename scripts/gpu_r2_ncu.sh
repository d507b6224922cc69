#!/bin/bash
# Round 2 ncu evidence: launch list of the default bench (host-polled driver: same kernels as the graph),
# ncu --set full of the persistent small-block solver, the grouped grafting kernels, the Newton GEMM phases.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# 1. launch list (the graph driver hides the loop body from per-launch listing: use the polled driver)
PC_ROOT_MODE=poll timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:'tc_phase|root_|power_iteration|simt|select|quant|small_root' -c 1500 --csv \
  --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-step > /dev/null 2>&1
echo "launch list rows: $(wc -l < gpurun_out/r02_launches_bench.csv)"
# 2. small_root_kernel, full set, one launch of the C1 batch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:small_root -s 3 -c 1 \
  -o gpurun_out/r02_small_root -f python scripts/small_root_prof_noenv.py > gpurun_out/r02_ncu_small.log 2>&1
ncu -i gpurun_out/r02_small_root.ncu-rep --page raw --csv > gpurun_out/r02_small_root_raw.csv 2>/dev/null
# 3. grouped grafting kernels on the ResNet-50 step
timeout 600 ncu --set full --clock-control none -k regex:graft_group -s 12 -c 3 \
  -o gpurun_out/r02_graft_group -f python scripts/step_profile.py resnet > gpurun_out/r02_ncu_graft.log 2>&1
ncu -i gpurun_out/r02_graft_group.ncu-rep --page raw --csv > gpurun_out/r02_graft_group_raw.csv 2>/dev/null
# 4. Newton GEMM phases (one iteration = 3 launches) of the headline call
PC_ROOT_MODE=poll timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_phase -s 40 -c 3 \
  -o gpurun_out/r02_tc_phase -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-step > gpurun_out/r02_ncu_tc.log 2>&1
ncu -i gpurun_out/r02_tc_phase.ncu-rep --page raw --csv > gpurun_out/r02_tc_phase_raw.csv 2>/dev/null
ls -la gpurun_out | grep r02_
