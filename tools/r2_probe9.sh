#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== tool standalone"; timeout 600 python tools/train_step_timing.py --batch 2 --height 768 --width 768 --steps 3 --warmup 2 --out gpurun_out/t1.json 2>&1 | tail -1 | cut -c1-400
echo "== bench --workload train standalone"; timeout 600 python bench.py --workload train --steps 3 --warmup 2 2>gpurun_out/t2.err | cut -c1-900
echo "== bench default (train child after inference), no cpu baseline"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/t3.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['train_step'])"
nvidia-smi --query-gpu=clocks.sm,power.draw,memory.used --format=csv
