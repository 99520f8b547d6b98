#!/bin/bash
# one gpurun call: GPU parity tests, smoke, a short bench line, and (arg "ncu") the ncu launch list of the same bench command
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log )
tail -6 gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log ); tail -3 gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_busy_frac','clocks','stage_busy_ms_per_step')}); print(d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['e2e'])
PY
if [ "$1" = "ncu" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu exit $?"
fi
