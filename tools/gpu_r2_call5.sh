#!/bin/bash
# round 2, fifth GPU call: full GPU suite; bench (warp-cooperative resolve walks, tile run-ahead, faster distribute); c3; ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2c5_pytest.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --no-cpu-baseline --repeats 3 > gpurun_out/r2c5_${tag}.json 2> gpurun_out/r2c5_${tag}.err; echo "bench $tag exit $?"; }
run default
run cluster PLVS_MATCH_RESOLVE=cluster
timeout 600 python bench.py --config c3 --no-cpu-baseline --steps 10 --repeats 3 > gpurun_out/r2c5_c3.json 2> gpurun_out/r2c5_c3.err; echo "bench c3 exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c5_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"], d.get("latency"))
    except Exception as e:
        print(f, "no line", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_resolve_cta|k_candidates|k_build_grid" -s 40 -c 6 -f -o gpurun_out/r2c5_match_kernels python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency > gpurun_out/r2c5_match_ncu.log 2>&1; echo "ncu matcher exit $?"
timeout 600 ncu --set full --clock-control none -k "regex:k_distribute|k_fast_cells|k_compact|k_orient|k_resize|k_blur|k_pack" -s 20 -c 24 -f -o gpurun_out/r2c5_orb_kernels python bench.py --steps 2 --warmup 1 --repeats 1 --batch 1 --no-cpu-baseline --no-latency > gpurun_out/r2c5_orb_ncu.log 2>&1; echo "ncu orb exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_integrate|k_classify|k_depth_tiles|k_commit" -s 60 -c 10 -f -o gpurun_out/r2c5_tsdf_kernels python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency > gpurun_out/r2c5_tsdf_ncu.log 2>&1; echo "ncu tsdf exit $?"
ls -la gpurun_out/*.ncu-rep
