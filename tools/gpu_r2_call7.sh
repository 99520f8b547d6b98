#!/bin/bash
# round 2, seventh GPU call, TWO GPUs: TMA probe, stereo eye split (test + bench), colour merge over NCCL with its bandwidth, the bench at N=2, priority A/B at N=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
for m in 1 0; do timeout 60 tools/tma_probe.bin $m > gpurun_out/r2c7_tma_probe_$m.log 2>&1; echo "tma probe mode $m exit $?"; cat gpurun_out/r2c7_tma_probe_$m.log; done
timeout 600 python -m pytest tests/test_gpu_stereo_split.py tests/test_gpu_widened.py tests/test_gpu_orb.py -m gpu -q -p no:cacheprovider > gpurun_out/r2c7_pytest.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2c7_pytest.log
timeout 300 python tools/bench_stereo_split.py > gpurun_out/r2c7_stereo_split.jsonl 2> gpurun_out/r2c7_stereo_split.err; echo "stereo bench exit $?"; cat gpurun_out/r2c7_stereo_split.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/merge_2gpu_check.py > gpurun_out/r2c7_merge_2gpu.log 2>&1; echo "merge exit $?"; tail -5 gpurun_out/r2c7_merge_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --repeats 3 --no-latency > gpurun_out/r2c7_bench_2gpu.json 2> gpurun_out/r2c7_bench_2gpu.err; echo "bench 2gpu exit $?"
run() { tag=$1; shift; env "$@" timeout 400 python bench.py --no-cpu-baseline --repeats 3 --no-latency > gpurun_out/r2c7_${tag}.json 2> gpurun_out/r2c7_${tag}.err; echo "bench $tag exit $?"; }
run default
run orbfirst PLVS_STREAM_ORDER=orb
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c7_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), d.get("value_passes"), round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 4), d["stage_busy_ms_per_step"])
        print("   ", d["kernel_ms_per_step"])
    except Exception as e:
        print(f, "no line", e)
PY
