#!/bin/bash
# round-2 GPU call #2 (1 GPU): full -m gpu suite, default bench line (incl. training child), wgrad variants timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -15 gpurun_out/pytest_gpu2.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
for v in "0 0" "1 0" "0 296" "1 296"; do
  set -- $v
  B200_WGRAD_PADDED=$1 B200_WGRAD_SPLIT_K=$2 timeout 600 python tools/train_step_timing.py --batch 2 --height 768 --width 768 --steps 3 --warmup 2 --breakdown --out gpurun_out/train_p$1_s$2.json > gpurun_out/train_p$1_s$2.log 2>&1
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/train_p$1_s$2.json")); print("wgrad padded=$1 split=$2", d["ms_per_step"], d["forward_ms"], d["backward_optimizer_ms"], d["peak_mem_gb"], list(d["breakdown_ms"].items())[:8])
except Exception as e: print("ERR $1 $2", e)
P
done
