#!/bin/bash
# round 2, twelfth GPU call: parallel std::sort emulation in the distributor, two frame-construction threads, resolve phase counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c12_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2c12_pytest.log
timeout 300 python tools/resolve_phases.py 5 > gpurun_out/r2c12_resolve_phases.log 2>&1; echo "phases exit $?"; cat gpurun_out/r2c12_resolve_phases.log | tail -12
run() { tag=$1; shift; env "$@" timeout 500 python bench.py --no-cpu-baseline --repeats 5 --no-latency > gpurun_out/r2c12_${tag}.json 2> gpurun_out/r2c12_${tag}.err; echo "bench $tag exit $?"; }
timeout 500 python bench.py --no-cpu-baseline --repeats 5 > gpurun_out/r2c12_default.json 2> gpurun_out/r2c12_default.err; echo "bench default exit $?"
run ex2 PLVS_PIPELINE_EXTRACTORS=2
run ex6 PLVS_PIPELINE_EXTRACTORS=6
run default_again
timeout 600 python bench.py --config c3 --steps 10 --no-cpu-baseline --repeats 3 --no-latency > gpurun_out/r2c12_c3.json 2> gpurun_out/r2c12_c3.err; echo "bench c3 exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c12_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), d["e2e"].get("passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"], d.get("latency"))
    except Exception as e:
        print(f, "no line", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_resize|k_fast|k_compact|k_distribute|k_pack|k_blur|k_orient|k_build_grid" -s 60 -c 60 --csv --log-file gpurun_out/r2c12_launches_b1.csv python bench.py --steps 2 --warmup 1 --repeats 1 --batch 1 --no-cpu-baseline --no-latency --driver python > gpurun_out/r2c12_launches_b1.log 2>&1; echo "ncu launches b1 exit $?"
