#!/bin/bash
# round-2 end-state check on one B200: full GPU suite, smoke(), the default bench line, the reference arm
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_final.log
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value","ms_per_step","unet_fwd_ms","unet_tensor_frac","step_tensor_frac","gpu_launches")}, d["e2e"], d["roofline"]["achieved"], d["roofline"]["frac"], d["fast_mode"], d["clocks"])
print(d["breakdown_ms_eager_step"]); print(d.get("train_step")); print(d.get("cpu_baseline"))
P
