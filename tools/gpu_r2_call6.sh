#!/bin/bash
# round 2, sixth GPU call: full suite (TMA-staged FAST, counting distributor, Deform / world cloud on the device), bench c2 + c3, TMA A/B, ncu of the extractor and matcher
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c6_pytest.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2c6_pytest.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --no-cpu-baseline --repeats 3 > gpurun_out/r2c6_${tag}.json 2> gpurun_out/r2c6_${tag}.err; echo "bench $tag exit $?"; }
run default
run notma PLVS_ORB_TMA=0
timeout 600 python bench.py --config c3 --no-cpu-baseline --steps 10 --repeats 3 > gpurun_out/r2c6_c3.json 2> gpurun_out/r2c6_c3.err; echo "bench c3 exit $?"
PLVS_ORB_TMA=0 timeout 600 python bench.py --config c3 --no-cpu-baseline --steps 10 --repeats 3 --no-latency > gpurun_out/r2c6_c3_notma.json 2> gpurun_out/r2c6_c3_notma.err; echo "bench c3 notma exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c6_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"], d.get("latency"))
    except Exception as e:
        print(f, "no line", e)
PY
timeout 600 ncu --set full --clock-control none -k "regex:k_distribute|k_fast_cells|k_compact|k_orient|k_blur|k_pack" -s 12 -c 12 -f -o gpurun_out/r2c6_orb_kernels python bench.py --steps 2 --warmup 1 --repeats 1 --batch 1 --no-cpu-baseline --no-latency > gpurun_out/r2c6_orb_ncu.log 2>&1; echo "ncu orb exit $?"
timeout 600 ncu --set full --clock-control none -k "regex:k_distribute|k_fast_cells" -s 4 -c 4 -f -o gpurun_out/r2c6_orb_c3_kernels python bench.py --config c3 --steps 2 --warmup 1 --repeats 1 --batch 1 --no-cpu-baseline --no-latency > gpurun_out/r2c6_orb_c3_ncu.log 2>&1; echo "ncu orb c3 exit $?"
timeout 600 ncu --set full --clock-control none -k "regex:k_resolve_cta" -s 8 -c 4 -f -o gpurun_out/r2c6_resolve python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency > gpurun_out/r2c6_resolve_ncu.log 2>&1; echo "ncu resolve exit $?"
