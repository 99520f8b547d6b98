#!/bin/bash
# round 2, ninth GPU call: TMA-staged FAST cells (16-byte aligned box) -- tests, A/B at VGA and 1080p, ncu of k_fast_cells both ways
mkdir -p gpurun_out
PLVS_ORB_TMA=1 timeout 600 python -m pytest tests/test_gpu_orb.py tests/test_gpu_bench_config.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c9_pytest_tma.log 2>&1; echo "pytest (TMA on) exit $?"; tail -4 gpurun_out/r2c9_pytest_tma.log
timeout 600 python -m pytest tests/test_gpu_orb.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c9_pytest.log 2>&1; echo "pytest (TMA off) exit $?"; tail -2 gpurun_out/r2c9_pytest.log
run() { tag=$1; shift; env "$@" timeout 500 python bench.py --no-cpu-baseline --repeats 5 --no-latency > gpurun_out/r2c9_${tag}.json 2> gpurun_out/r2c9_${tag}.err; echo "bench $tag exit $?"; }
run c2_plain
run c2_tma PLVS_ORB_TMA=1
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --config c3 --steps 10 --no-cpu-baseline --repeats 3 --no-latency > gpurun_out/r2c9_${tag}.json 2> gpurun_out/r2c9_${tag}.err; echo "bench $tag exit $?"; }
run c3_plain
run c3_tma PLVS_ORB_TMA=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c9_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), round(d["roofline"]["frac"], 3), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"])
    except Exception as e:
        print(f, "no line", e)
PY
for t in 0 1; do
PLVS_ORB_TMA=$t timeout 600 ncu --set full --clock-control none -k "regex:k_fast_cells|k_compact" -s 2 -c 4 -f -o gpurun_out/r2c9_fast_c3_tma$t python bench.py --config c3 --steps 2 --warmup 1 --repeats 1 --batch 1 --no-cpu-baseline --no-latency > gpurun_out/r2c9_fast_ncu_$t.log 2>&1; echo "ncu fast c3 tma=$t exit $?"
done
