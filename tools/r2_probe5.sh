#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for k in conv128 conv256 conv512; do for f in 0 2 4 6 22; do
  echo -n "halo $k flags=$f: "; timeout 120 python tools/prof_kernels.py $k 5 $f 2>&1 | tail -1
done; done | tee gpurun_out/isolate_halo.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --dump-shapes > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench5.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["unet_fwd_ms"], d["roofline"]["achieved"], d["fast_mode"], d["clocks"], d["breakdown_ms_eager_step"])
P
head -40 gpurun_out/conv_shapes.txt
