#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 60 tools/micro/umma_rowoffset_test 2>&1 | tee gpurun_out/umma_rowoffset.txt
HALO="conv_halo_128 conv_halo_res_f32_stats conv_halo_shortcut_320 conv_halo_edges_768 conv_halo_ragged"
echo "== halo checks (base offset from address)"; timeout 600 python tools/gpu_kernel_check.py $HALO 2>&1 | tail -8
echo "== halo checks (base offset 0)"; B200_DEBUG_FLAGS=32 timeout 600 python tools/gpu_kernel_check.py $HALO 2>&1 | tail -8
for k in conv128 res32 conv256 conv512 conv320; do echo -n "$k: "; timeout 120 python tools/prof_kernels.py $k 5 2>&1 | tail -1; done | tee gpurun_out/halo_speed.txt
