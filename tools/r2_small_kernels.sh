#!/bin/bash
# softmax / LayerNorm kernel rework: correctness first, then timings (L2-cold-ish: inputs are 0.1-1 GB)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "softmax or ln_ or attn_lse or vae" 2>&1 | tail -5 > gpurun_out/small_kernels_tests.log
cat gpurun_out/small_kernels_tests.log
for k in softmax ln320 ln640 ln1280; do timeout 120 python tools/prof_kernels.py $k 20; done 2>&1 | tee gpurun_out/small_kernels_timing.log
