#!/bin/bash
# one `ncu --set full` capture of ~one pipeline step (every kernel type of the hot path), summarised on the box:
# the .ncu-rep of 130 kernels is too large to bring back, the raw CSV is not
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -s 700 -c 130 -f -o /tmp/full_step python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/full_step.log 2>&1; echo "ncu full exit $?"
ncu -i /tmp/full_step.ncu-rep --page raw --csv > gpurun_out/full_step_raw.csv 2> gpurun_out/full_step_raw.err; echo "export exit $?"; ls -la /tmp/full_step.ncu-rep gpurun_out/full_step_raw.csv
# the dominant kernel alone, with source correlation, small enough to bring back
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 60 -c 1 -f -o gpurun_out/k_integrate_r1_v4 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/k_integrate_ncu.log 2>&1; echo "ncu k_integrate exit $?"
ls -la gpurun_out/*.ncu-rep
