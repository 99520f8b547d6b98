#!/bin/bash
# round 2, second GPU call: the whole GPU suite (no -x), the new bench line (native stage threads, one-CTA resolve), the Python-driver A/B,
# the launch list and a full ncu capture of the matcher kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_native.json 2> gpurun_out/r2_bench_native.err; echo "bench native exit $?"; tail -3 gpurun_out/r2_bench_native.err
timeout 400 python bench.py --no-cpu-baseline --no-latency --driver python --repeats 3 > gpurun_out/r2_bench_python.json 2> gpurun_out/r2_bench_python.err; echo "bench python exit $?"
PLVS_TSDF_SM_RESERVE=8 timeout 400 python bench.py --no-cpu-baseline --no-latency --repeats 3 > gpurun_out/r2_bench_native_reserve8.json 2> gpurun_out/r2_bench_native_reserve8.err; echo "bench reserve8 exit $?"
PLVS_MATCH_RESOLVE=cluster timeout 400 python bench.py --no-cpu-baseline --no-latency --repeats 3 > gpurun_out/r2_bench_native_cluster.json 2> gpurun_out/r2_bench_native_cluster.err; echo "bench cluster-resolve exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"], d.get("latency"))
        print("   ", d["kernel_ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d["e2e"]["d2h_bytes_per_step"])
    except Exception as e:
        print(f, "no line", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency > gpurun_out/r2_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_resolve_cta|k_candidates|k_build_grid" -s 40 -c 8 -f -o gpurun_out/r2_match_kernels python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency > gpurun_out/r2_match_ncu.log 2>&1; echo "ncu matcher exit $?"
ls -la gpurun_out/*.ncu-rep
