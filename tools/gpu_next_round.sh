#!/bin/bash
# first GPU call of the next round: the tests that never met a GPU (shown as real passes / failures), then the timings of the widened rows, then
# ncu over the new kernels.  Usage: gpurun --timeout 900 -- 'bash tools/gpu_next_round.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_widened.py -q -rA --runxfail -p no:cacheprovider > gpurun_out/unverified_tests.log 2>&1; echo "unverified tests exit $?"; tail -15 gpurun_out/unverified_tests.log
timeout 300 python tools/bench_extras.py > gpurun_out/bench_extras.jsonl 2> gpurun_out/bench_extras.err; echo "extras exit $?"; cat gpurun_out/bench_extras.jsonl
timeout 300 ncu --set full --clock-control none -k "regex:k_mesh_|k_bow_|k_init_|k_compact_queries|k_undistort|k_depth_u16" -c 24 -f -o /tmp/extras python tools/bench_extras.py > gpurun_out/extras_ncu.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/extras.ncu-rep --page raw --csv > gpurun_out/extras_raw.csv 2>/dev/null; ls -la gpurun_out/extras_raw.csv
# A/B of the SM-sharing knobs: same bench, one JSON line each
for cfg in "" "PLVS_TSDF_CTAS_PER_SM=1 PLVS_MATCH_RESOLVE_SMEM=0" "PLVS_MATCH_RESOLVE_SMEM=0" "PLVS_TSDF_SM_RESERVE=16" "PLVS_TSDF_SM_RESERVE=32"; do
  tag=$(echo "${cfg:-default}" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_${tag}.json 2> gpurun_out/bench_ab_${tag}.err; echo "bench [$cfg] exit $?"; cut -c1-260 gpurun_out/bench_ab_${tag}.json
done
# the open FAST min/max-tree issue: score-map mismatches of the tree path on the device, with the ring values of the first few
PLVS_FAST_TREE=1 PLVS_ORB_DEBUG=1 timeout 120 python tools/fast_tree_probe.py > gpurun_out/fast_tree_probe.log 2>&1; echo "fast tree probe exit $?"; tail -12 gpurun_out/fast_tree_probe.log
