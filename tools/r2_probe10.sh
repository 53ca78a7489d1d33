#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -4
for f in 0 64; do for k in conv128 res32 conv256 conv512 conv320; do echo -n "flags=$f $k: "; B200_STATS=1 timeout 120 python tools/prof_kernels.py $k 5 $f 2>&1 | tail -1; done; done
for k in linear lin320; do echo -n "$k: "; timeout 120 python tools/prof_kernels.py $k 5 2>&1 | tail -1; done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --dump-shapes > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench10.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["unet_fwd_ms"], d["roofline"]["achieved"], d["fast_mode"], d["clocks"], d["breakdown_ms_eager_step"])
P
head -8 gpurun_out/conv_shapes.txt; head -10 gpurun_out/linear_shapes.txt
