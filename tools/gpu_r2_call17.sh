#!/bin/bash
# round 2, call 17 (gpurun --gpus 2): the 2-GPU tests (stereo eye split, NCCL merge with colour) on the final build
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stereo_split.py tests/test_gpu_merge.py -m gpu -q -p no:cacheprovider > gpurun_out/r2c17_pytest_2gpu.log 2>&1; echo "pytest 2gpu exit $?"; tail -3 gpurun_out/r2c17_pytest_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/merge_2gpu_check.py > gpurun_out/r2c17_merge_2gpu.log 2>&1; echo "merge check exit $?"; grep -E "PASS|FAIL|merge_exchange" gpurun_out/r2c17_merge_2gpu.log | cut -c1-300
