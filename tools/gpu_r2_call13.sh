#!/bin/bash
# round 2, call 13: warp-aggregated walk queue; stream-order A/B (TSDF first); reference arm; smoke
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c13_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2c13_pytest.log
timeout 300 python tools/resolve_phases.py 3 > gpurun_out/r2c13_resolve_phases.log 2>&1; tail -7 gpurun_out/r2c13_resolve_phases.log
run() { tag=$1; shift; env "$@" timeout 500 python bench.py --no-cpu-baseline --repeats 5 --no-latency > gpurun_out/r2c13_${tag}.json 2> gpurun_out/r2c13_${tag}.err; echo "bench $tag exit $?"; }
run default
run tsdf_first PLVS_STREAM_ORDER=tsdf
run default_again
run tsdf_first_again PLVS_STREAM_ORDER=tsdf
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c13_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), d["e2e"].get("passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
    except Exception as e:
        print(f, "no line", e)
PY
timeout 900 python bench.py --impl reference > gpurun_out/r2c13_reference.json 2> gpurun_out/r2c13_reference.err; echo "reference arm exit $?"; tail -c 1500 gpurun_out/r2c13_reference.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2c13_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2c13_smoke.log
