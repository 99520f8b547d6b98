#!/bin/bash
# round 2, call 20: compute-sanitizer memcheck over the kernels that are new or rewritten this round (claim resolution, k_bind / cull split, distributor,
# compact, deform / world cloud, line k-NN, frame grid at extraction), through their GPU tests
mkdir -p gpurun_out
SAN="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20"
timeout 200 $SAN --log-file gpurun_out/r2c20_memcheck_widened.log python -m pytest tests/test_gpu_widened.py -m gpu -q -x -p no:cacheprovider -k "line_knn2 or deform or world_cloud or keyframe_ids or search_local_points or search_for_initialization" > gpurun_out/r2c20_widened.out 2>&1; echo "memcheck widened exit $?"; tail -2 gpurun_out/r2c20_widened.out; tail -2 gpurun_out/r2c20_memcheck_widened.log
timeout 200 $SAN --log-file gpurun_out/r2c20_memcheck_match.log python -m pytest tests/test_gpu_match.py -m gpu -q -x -p no:cacheprovider -k "resolve_variants or projection_map or projection_last or grid_cache" > gpurun_out/r2c20_match.out 2>&1; echo "memcheck match exit $?"; tail -2 gpurun_out/r2c20_match.out; tail -2 gpurun_out/r2c20_memcheck_match.log
timeout 240 $SAN --log-file gpurun_out/r2c20_memcheck_orb_tsdf.log python -m pytest tests/test_gpu_orb.py tests/test_gpu_tsdf.py -m gpu -q -x -p no:cacheprovider -k "distributor_state or orb_vga_2000 or batch_matches_single or scan_sequence or scan_color_sequence or carving_moves or nan_zero or many_scans" > gpurun_out/r2c20_orb_tsdf.out 2>&1; echo "memcheck orb+tsdf exit $?"; tail -2 gpurun_out/r2c20_orb_tsdf.out; tail -2 gpurun_out/r2c20_memcheck_orb_tsdf.log
