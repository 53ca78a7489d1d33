#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -5
for k in conv128 res32 conv256 conv512 conv320 conv640 conv1280 conv1280s; do echo -n "$k: "; timeout 120 python tools/prof_kernels.py $k 5 2>&1 | tail -1; done | tee gpurun_out/halo_speed.txt
