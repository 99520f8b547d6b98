#!/bin/bash
# round 2, third GPU call: A/B of the bounded-lifetime k_integrate CTAs, the claim-resolution variants, the cv::cornerScore-style FAST score, and the C3 line
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 400 python bench.py --no-cpu-baseline --no-latency --repeats 3 > gpurun_out/r2c3_${tag}.json 2> gpurun_out/r2c3_${tag}.err; echo "bench $tag exit $?"
}
run default
run persistent PLVS_TSDF_ITEMS_PER_CTA=0
run items4w4 PLVS_TSDF_ITEMS_PER_CTA=4 PLVS_TSDF_GRID_WAVES=4
run items16w2 PLVS_TSDF_ITEMS_PER_CTA=16 PLVS_TSDF_GRID_WAVES=2
run cluster PLVS_MATCH_RESOLVE=cluster
run lpq0 PLVS_MATCH_RESOLVE_LPQ=0
run lpq2 PLVS_MATCH_RESOLVE_LPQ=2
PLVS_FAST_TREE=2 PLVS_ORB_DEBUG=1 timeout 120 python tools/fast_tree_probe.py > gpurun_out/r2c3_fast_tree2_probe.log 2>&1; echo "fast tree2 probe exit $?"; tail -3 gpurun_out/r2c3_fast_tree2_probe.log
PLVS_FAST_TREE=2 timeout 300 python -m pytest tests/test_gpu_orb.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c3_orb_tree2.log 2>&1; echo "orb tests with tree2 exit $?"; tail -3 gpurun_out/r2c3_orb_tree2.log
run fasttree2 PLVS_FAST_TREE=2
timeout 600 python bench.py --config c3 --no-cpu-baseline --steps 10 --repeats 3 > gpurun_out/r2c3_c3.json 2> gpurun_out/r2c3_c3.err; echo "bench c3 exit $?"; tail -3 gpurun_out/r2c3_c3.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c3_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"], d.get("latency"))
    except Exception as e:
        print(f, "no line", e)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r2c3_launches.csv python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency > gpurun_out/r2c3_launches.log 2>&1; echo "ncu launches exit $?"
