#!/bin/bash
# round 2, call 19 (gpurun --gpus 8): the final bench.py at N = 1 and N = 8 on one box (every rank bound to its GPU's socket, every rank the same stream)
mkdir -p gpurun_out
run() { n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then timeout 300 python bench.py --gpus 1 --no-cpu-baseline --no-latency --repeats 3 "$@" > gpurun_out/r2c19_${tag}.json 2> gpurun_out/r2c19_${tag}.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --no-cpu-baseline --no-latency --repeats 3 "$@" > gpurun_out/r2c19_${tag}.json 2> gpurun_out/r2c19_${tag}.err; fi
  echo "bench N=$n $tag exit $?"; }
run 1 n1
run 8 n8
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c19_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), d["e2e"].get("passes"), d["stage_busy_ms_per_step"], d.get("per_rank_ms_per_step"), d["config"].get("host_affinity"), d["config"].get("rank_streams"))
    except Exception as e:
        print(f, "no line", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
