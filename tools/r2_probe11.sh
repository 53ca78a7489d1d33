#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -6
timeout 900 python -m pytest tests/test_engine_gpu.py -q -x -k "backward or training or checkpoint or wgrad" 2>&1 | tail -5
timeout 600 python tools/train_step_timing.py --batch 2 --height 768 --width 768 --steps 4 --warmup 3 --breakdown --out gpurun_out/train_r2c.json > gpurun_out/train_r2c.log 2>&1
timeout 600 python tools/train_step_timing.py --batch 2 --height 512 --width 640 --steps 4 --warmup 3 --out gpurun_out/train_r2c_512x640.json > gpurun_out/train_r2c_512.log 2>&1
python - <<'P'
import json
for f in ("train_r2c", "train_r2c_512x640"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["forward_ms"], d["backward_optimizer_ms"], d["peak_mem_gb"], list(d["breakdown_ms"].items())[:12])
    except Exception as e: print("ERR", f, e)
P
tail -3 gpurun_out/train_r2c.log
