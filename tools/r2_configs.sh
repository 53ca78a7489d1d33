#!/bin/bash
# BASELINE.json configs 4 and 5 through bench.py on one B200
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out; : > gpurun_out/extra_configs_r02.jsonl
for r in 384 512 768 1024; do
  timeout 600 python bench.py --workload normals --res $r --steps 5 --warmup 3 --no-cpu-baseline --no-fast 2>>gpurun_out/extra.err | tail -1 >> gpurun_out/extra_configs_r02.jsonl
done
timeout 600 python bench.py --workload geowizard --steps 5 --warmup 3 --no-cpu-baseline --no-fast 2>>gpurun_out/extra.err | tail -1 >> gpurun_out/extra_configs_r02.jsonl
python - <<'P'
import json
for l in open("gpurun_out/extra_configs_r02.jsonl"):
    try:
        d=json.loads(l); print(d["metric"], d["config"]["workload"][:60], round(d["value"],1), "img/s", round(d["ms_per_step"],1), "ms", "e2e", round(d["e2e"]["value"],1), "conv", round(d["roofline"]["achieved"]), "TF/s")
    except Exception as e: print("ERR", e, l[:200])
P
tail -3 gpurun_out/extra.err
