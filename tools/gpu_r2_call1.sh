#!/bin/bash
# round 2, first GPU call: the whole GPU suite (new strict bench-config / shim-runtime / colour-merge tests), the baseline bench line,
# the SM-sharing A/B and the FAST min/max-tree probe.   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call1.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r2_pytest_gpu.log
for cfg in "" "PLVS_TSDF_SM_RESERVE=8" "PLVS_TSDF_SM_RESERVE=16" "PLVS_TSDF_CTAS_PER_SM=1 PLVS_MATCH_RESOLVE_SMEM=0" "PLVS_MATCH_RESOLVE_SMEM=0"; do
  tag=$(echo "${cfg:-default}" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_ab_${tag}.json 2> gpurun_out/r2_bench_ab_${tag}.err; echo "bench [$cfg] exit $?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_ab_${tag}.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["stage_busy_ms_per_step"])
except Exception as e:
    print("no line", e)
PY
done
PLVS_FAST_TREE=1 PLVS_ORB_DEBUG=1 timeout 120 python tools/fast_tree_probe.py > gpurun_out/r2_fast_tree_probe.log 2>&1; echo "fast tree probe exit $?"; tail -12 gpurun_out/r2_fast_tree_probe.log
nvidia-smi --query-gpu=name,memory.total --format=csv; nproc; free -g | head -2
