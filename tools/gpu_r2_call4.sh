#!/bin/bash
# round 2, fourth GPU call: matcher tests, the bench with the incremental resolve + FAST sequence, ncu of the matcher kernels, launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_match.py tests/test_gpu_orb.py tests/test_gpu_bench_config.py -m gpu -q -p no:cacheprovider > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2c4_pytest.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --no-cpu-baseline --repeats 3 > gpurun_out/r2c4_${tag}.json 2> gpurun_out/r2c4_${tag}.err; echo "bench $tag exit $?"; }
run default
run items0 PLVS_TSDF_ITEMS_PER_CTA=0
run cluster PLVS_MATCH_RESOLVE=cluster
timeout 600 python bench.py --config c3 --no-cpu-baseline --steps 10 --repeats 3 > gpurun_out/r2c4_c3.json 2> gpurun_out/r2c4_c3.err; echo "bench c3 exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c4_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"], d.get("latency"))
    except Exception as e:
        print(f, "no line", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_resolve_cta|k_candidates" -s 40 -c 6 -f -o gpurun_out/r2c4_match_kernels python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency > gpurun_out/r2c4_match_ncu.log 2>&1; echo "ncu matcher exit $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 700 --csv --log-file gpurun_out/r2c4_launches.csv python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency > gpurun_out/r2c4_launches.log 2>&1; echo "ncu launches exit $?"
