#!/bin/bash
# final evidence of a round: full bench line (with the CPU baseline), reference arm, launch list, ORB kernels under ncu --set full
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu launches exit $?"
timeout 400 ncu --set full --clock-control none -k "regex:k_fast_cells|k_distribute|k_orient_describe|k_blur|k_resize_level|k_compact|k_pack_selected" -s 40 -c 16 -f -o /tmp/orb_step python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/orb_ncu.log 2>&1; echo "ncu orb exit $?"
ncu -i /tmp/orb_step.ncu-rep --page raw --csv > gpurun_out/orb_step_raw.csv 2>/dev/null; ls -la gpurun_out/orb_step_raw.csv
