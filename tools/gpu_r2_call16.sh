#!/bin/bash
# round 2, call 16 (gpurun --gpus 8): the bench at N = 1, 2, 4, 8 back to back on one box (one camera stream per rank, weak scaling), the 2-GPU tests
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() { n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then timeout 400 python bench.py --gpus 1 --no-cpu-baseline --no-latency --repeats 3 "$@" > gpurun_out/r2c16_${tag}.json 2> gpurun_out/r2c16_${tag}.err
  else timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --no-cpu-baseline --no-latency --repeats 3 "$@" > gpurun_out/r2c16_${tag}.json 2> gpurun_out/r2c16_${tag}.err; fi
  echo "bench N=$n $tag exit $?"; }
run 1 n1
run 2 n2
run 4 n4
run 8 n8
python - <<'PY'
import json, glob
base = {}
for f in sorted(glob.glob("gpurun_out/r2c16_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), d["e2e"].get("passes"), round(d["roofline"]["frac"], 3), d["stage_busy_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "no line", e)
PY
