#!/bin/bash
# round-2 GPU call #1: full -m gpu suite, never-on-GPU checks, error budget, bench (fp32 / fp16 stream), ncu captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python tools/pending_gpu_checks.py > gpurun_out/pending.log 2>&1; echo "pending rc=$?" >> gpurun_out/pending.log
grep -E "PASS|FAIL|rc=" gpurun_out/pending.log
timeout 600 python tools/error_budget.py 24 > gpurun_out/error_budget_24.txt 2>&1
tail -12 gpurun_out/error_budget_24.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-shapes > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
cp gpurun_out/conv_shapes.txt gpurun_out/conv_shapes_fp32.txt; cp gpurun_out/linear_shapes.txt gpurun_out/linear_shapes_fp32.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --stream fp16 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
python - <<'P'
import json
for f in ("bench_fp32","bench_fp16"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["unet_fwd_ms"], d["roofline"]["frac"], d["breakdown_ms_eager_step"])
    except Exception as e:
        print(f, "ERR", e)
P
for k in gn32 gn attn lin320 conv1280s conv128; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gn_apply|attention_d64|gemm_conv" -s 2 -c 1 -o gpurun_out/ncu_$k -f python tools/prof_kernels.py $k 3 > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/ncu_$k.ncu-rep --page raw --csv > gpurun_out/ncu_$k.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$k.ncu-rep --page details --csv > gpurun_out/ncu_$k.details.csv 2>/dev/null
done
du -sh gpurun_out
# keep the merge-back under the 64 MiB cap: CSV pages are enough for all but the two kernels being tuned
rm -f gpurun_out/ncu_gn.ncu-rep gpurun_out/ncu_conv128.ncu-rep gpurun_out/ncu_conv1280s.ncu-rep
ls -la gpurun_out | head -40
