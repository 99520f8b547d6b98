#!/bin/bash
# round 2, call 14: ORB tests after the distributor / compact changes, single-frame launch list, ncu --set full captures for profiles/
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_orb.py tests/test_gpu_bench_config.py -m gpu -q -p no:cacheprovider > gpurun_out/r2c14_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2c14_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_resize|k_fast|k_compact|k_distribute|k_pack|k_blur|k_orient|k_build_grid" -s 60 -c 60 --csv --log-file gpurun_out/r2c14_launches_b1.csv python bench.py --steps 2 --warmup 1 --repeats 1 --batch 1 --no-cpu-baseline --no-latency --driver python > gpurun_out/r2c14_launches_b1.log 2>&1; echo "ncu launches b1 exit $?"
timeout 500 python bench.py --no-cpu-baseline --repeats 5 > gpurun_out/r2c14_default.json 2> gpurun_out/r2c14_default.err; echo "bench default exit $?"
timeout 600 python bench.py --config c3 --steps 10 --no-cpu-baseline --repeats 3 > gpurun_out/r2c14_c3.json 2> gpurun_out/r2c14_c3.err; echo "bench c3 exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c14_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), d["e2e"].get("passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"], d.get("latency"))
    except Exception as e:
        print(f, "no line", e)
PY
# captures for profiles/: every kernel of a step once (python driver: deterministic launch order), full sections
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_integrate|k_bind|k_classify|k_commit|k_depth_tiles" -s 20 -c 14 -f -o gpurun_out/r2c14_tsdf python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency --driver python > gpurun_out/r2c14_tsdf_ncu.log 2>&1; echo "ncu tsdf exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_resolve_cta|k_candidates|k_triangulate|k_tri_finish" -s 16 -c 12 -f -o gpurun_out/r2c14_match python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency --driver python > gpurun_out/r2c14_match_ncu.log 2>&1; echo "ncu match exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_resize|k_fast|k_compact|k_distribute|k_pack|k_blur|k_orient|k_build_grid" -s 14 -c 14 -f -o gpurun_out/r2c14_orb_b1 python bench.py --steps 2 --warmup 1 --repeats 1 --batch 1 --no-cpu-baseline --no-latency --driver python > gpurun_out/r2c14_orb_ncu.log 2>&1; echo "ncu orb exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 700 --csv --log-file gpurun_out/r2c14_launches.csv python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency --driver python > gpurun_out/r2c14_launches.log 2>&1; echo "ncu launches exit $?"
ls -la gpurun_out/r2c14_*.ncu-rep
