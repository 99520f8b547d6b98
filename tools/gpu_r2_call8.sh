#!/bin/bash
# round 2, eighth GPU call: TMA probe (one configuration per process), integrate-CTA lifetime A/B with 5 passes each, ncu of resolve / extractor
mkdir -p gpurun_out
for cfg in "1 16 19 0 0" "1 53 19 0 0" "1 48 19 0 0" "1 16 60 0 0" "1 16 19 1 0" "1 16 19 0 5" "1 32 60 1 5" "1 53 60 1 5" "1 64 41 2 9" "0 53 19 0 0"; do
  timeout 30 tools/tma_probe.bin $cfg 2>&1 | grep "^mode"
done > gpurun_out/r2c8_tma_probe.log; cat gpurun_out/r2c8_tma_probe.log
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --no-cpu-baseline --repeats 5 --no-latency > gpurun_out/r2c8_${tag}.json 2> gpurun_out/r2c8_${tag}.err; echo "bench $tag exit $?"; }
run persistent
run b8w2 PLVS_TSDF_ITEMS_PER_CTA=8 PLVS_TSDF_GRID_WAVES=2
run b4w4 PLVS_TSDF_ITEMS_PER_CTA=4 PLVS_TSDF_GRID_WAVES=4
run persistent_again
run b8w2_again PLVS_TSDF_ITEMS_PER_CTA=8 PLVS_TSDF_GRID_WAVES=2
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c8_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), d["e2e"].get("passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"])
    except Exception as e:
        print(f, "no line", e)
PY
timeout 600 ncu --set full --clock-control none -k "regex:k_resolve_cta" -s 8 -c 4 -f -o gpurun_out/r2c8_resolve python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-latency > gpurun_out/r2c8_resolve_ncu.log 2>&1; echo "ncu resolve exit $?"
timeout 600 ncu --set full --clock-control none -k "regex:k_distribute|k_fast_cells|k_compact|k_orient|k_blur|k_pack" -s 12 -c 12 -f -o gpurun_out/r2c8_orb_kernels python bench.py --steps 2 --warmup 1 --repeats 1 --batch 1 --no-cpu-baseline --no-latency > gpurun_out/r2c8_orb_ncu.log 2>&1; echo "ncu orb exit $?"
