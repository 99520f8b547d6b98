#!/bin/bash
# round 2, call 23: the last GPU seconds of the round -- one short c2 run of the final build
mkdir -p gpurun_out
timeout 100 python bench.py --no-cpu-baseline --no-latency --repeats 3 > gpurun_out/r2c23_c2_final.json 2> gpurun_out/r2c23_c2_final.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r2c23_c2_final.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['e2e']['value']), d['value_passes'], d['e2e']['passes'], round(d['roofline']['frac'],3), d['stage_busy_ms_per_step'])"
