#!/bin/bash
# round 2, call 15: final-code validation on one GPU -- full suite, smoke, c2 / c3 bench lines, launch lists
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c15_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2c15_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c15_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2c15_smoke.log
timeout 600 python bench.py > gpurun_out/r2c15_c2.json 2> gpurun_out/r2c15_c2.err; echo "bench c2 exit $?"
timeout 600 python bench.py --no-cpu-baseline --no-latency > gpurun_out/r2c15_c2_again.json 2> gpurun_out/r2c15_c2_again.err; echo "bench c2 again exit $?"
timeout 600 python bench.py --config c3 --steps 10 --no-cpu-baseline > gpurun_out/r2c15_c3.json 2> gpurun_out/r2c15_c3.err; echo "bench c3 exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c15_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), d["e2e"].get("passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"], d.get("latency"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "no line", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_resize|k_fast|k_compact|k_distribute|k_pack|k_blur|k_orient|k_build_grid" -s 60 -c 60 --csv --log-file gpurun_out/r2c15_launches_b1.csv python bench.py --steps 2 --warmup 1 --repeats 1 --batch 1 --no-cpu-baseline --no-latency --driver python > gpurun_out/r2c15_launches_b1.log 2>&1; echo "ncu launches b1 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/r2c15_launches.csv python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency --driver python > gpurun_out/r2c15_launches.log 2>&1; echo "ncu launches exit $?"
