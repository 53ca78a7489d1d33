#!/bin/bash
# round-2 GPU call #3 (1 GPU): rest of the suite with the vectorised epilogue, load/MMA isolation experiments, parity modes, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -12 gpurun_out/pytest_gpu3.log
for k in conv128 res32 conv256 conv512; do for f in 0 2 4 6 22; do
  echo -n "$k flags=$f: "; timeout 120 python tools/prof_kernels.py $k 5 $f 2>&1 | tail -1
done; done | tee gpurun_out/isolate.txt
timeout 900 python tools/parity_modes.py 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --dump-shapes > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench3.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["unet_fwd_ms"], d["roofline"]["achieved"], d["fast_mode"], d["clocks"], d["breakdown_ms_eager_step"])
P
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --no-fast --stream mixed > gpurun_out/bench3_mixed.json 2> gpurun_out/bench3_mixed.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench3_mixed.json").read().strip().splitlines()[-1])
print("mixed", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["breakdown_ms_eager_step"])
P
