#!/bin/bash
# ncu --set full of the power-iteration kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PC_ROOT_MODE=poll timeout 900 ncu --set full --import-source on --clock-control none -k regex:'power_iteration' -s 2 -c 1 -f -o gpurun_out/r2n_pi python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-step --no-big > /dev/null 2>&1
ls -la gpurun_out/r2n_pi.ncu-rep
