#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/gpu_kernel_check.py linear_mn_w linear_mn_w_noswap linear_mn_aw linear_mn_aw_noswap linear_mn_a linear_mn_aw_batched 2>&1 | tail -9
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu7.log
timeout 600 python tools/train_step_timing.py --batch 2 --height 768 --width 768 --steps 3 --warmup 2 --breakdown --out gpurun_out/train_r2b.json > gpurun_out/train_r2b.log 2>&1
python - <<'P'
import json
try:
    d=json.load(open("gpurun_out/train_r2b.json")); print("train", d["ms_per_step"], d["forward_ms"], d["backward_optimizer_ms"], d["peak_mem_gb"], list(d["breakdown_ms"].items())[:14])
except Exception as e: print("ERR", e); print(open("gpurun_out/train_r2b.log").read()[-1500:])
P
