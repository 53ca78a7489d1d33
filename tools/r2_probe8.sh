#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -4
timeout 1200 python bench.py --steps 10 --warmup 3 --dump-shapes > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["unet_fwd_ms"], d["roofline"]["achieved"], d["fast_mode"], d["clocks"], d["breakdown_ms_eager_step"])
print(d.get("train_step")); print(d.get("cpu_baseline"))
P
head -16 gpurun_out/conv_shapes.txt
