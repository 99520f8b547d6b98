#!/bin/bash
# round 2, call 22 (last): the full GPU suite on the final build; racecheck of the claim resolution after the __syncwarp fix
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c22_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2c22_pytest.log
timeout 120 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 10 --log-file gpurun_out/r2c22_racecheck_match.log python -m pytest tests/test_gpu_match.py -m gpu -q -x -p no:cacheprovider -k "resolve_variants or projection_last_competition or projection_map_claims" > gpurun_out/r2c22_match.out 2>&1; echo "racecheck match exit $?"; tail -1 gpurun_out/r2c22_match.out; tail -2 gpurun_out/r2c22_racecheck_match.log
