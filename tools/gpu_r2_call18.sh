#!/bin/bash
# round 2, call 18 (gpurun --gpus 8): does binding every rank to its GPU's socket straighten the scaling curve?  N = 1 and N = 8 bound, N = 8 unbound, same box
mkdir -p gpurun_out
lscpu | grep -E "Socket|NUMA|Model name|^CPU\(s\)" | head -8
for g in 0 4; do cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $g | tr 'A-Z' 'a-z' | sed 's/^0000//')/local_cpulist 2>/dev/null; done
run() { n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then env "$@" timeout 400 python bench.py --gpus 1 --no-cpu-baseline --no-latency --repeats 3 > gpurun_out/r2c18_${tag}.json 2> gpurun_out/r2c18_${tag}.err
  else env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --no-cpu-baseline --no-latency --repeats 3 > gpurun_out/r2c18_${tag}.json 2> gpurun_out/r2c18_${tag}.err; fi
  echo "bench N=$n $tag exit $?"; }
run 1 n1_bound PLVS_BENCH_NUMA=1
run 8 n8_bound PLVS_BENCH_NUMA=1
run 8 n8_unbound PLVS_BENCH_NUMA=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c18_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), d["e2e"].get("passes"), d["stage_busy_ms_per_step"], d["config"].get("host_affinity"))
    except Exception as e:
        print(f, "no line", e)
PY
