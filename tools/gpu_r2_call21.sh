#!/bin/bash
# round 2, call 21: compute-sanitizer racecheck (shared-memory hazards) over the kernels with hand-made shared-memory protocols
mkdir -p gpurun_out
SAN="compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 20"
timeout 240 $SAN --log-file gpurun_out/r2c21_racecheck_match.log python -m pytest tests/test_gpu_match.py -m gpu -q -x -p no:cacheprovider -k "resolve_variants or projection_last_competition or projection_map_claims" > gpurun_out/r2c21_match.out 2>&1; echo "racecheck match exit $?"; tail -2 gpurun_out/r2c21_match.out; tail -3 gpurun_out/r2c21_racecheck_match.log
timeout 240 $SAN --log-file gpurun_out/r2c21_racecheck_orb.log python -m pytest tests/test_gpu_orb.py tests/test_gpu_widened.py -m gpu -q -x -p no:cacheprovider -k "distributor_state or line_knn2" > gpurun_out/r2c21_orb.out 2>&1; echo "racecheck orb exit $?"; tail -2 gpurun_out/r2c21_orb.out; tail -3 gpurun_out/r2c21_racecheck_orb.log
