#!/bin/bash
# round-2 ncu captures for profiles/: final kernels, --set full (+ source for the conv)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for k in conv128 conv256 gn32 attn lin320 conv1280s; do
  src=""; [ "$k" = "conv128" ] && src="--import-source on"
  timeout 300 ncu --set full --clock-control none $src -k regex:"gn_apply|attention_d64|gemm_conv" -s 2 -c 1 -o gpurun_out/r2_$k -f python tools/prof_kernels.py $k 3 > gpurun_out/r2_$k.log 2>&1
  ncu -i gpurun_out/r2_$k.ncu-rep --page raw --csv > gpurun_out/r2_$k.raw.csv 2>/dev/null
done
ncu -i gpurun_out/r2_conv128.ncu-rep --page source --csv > gpurun_out/r2_conv128.source.csv 2>/dev/null
# launch list of one bench step (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train --no-fast > gpurun_out/r2_launches.log 2>&1
rm -f gpurun_out/r2_gn32.ncu-rep gpurun_out/r2_attn.ncu-rep gpurun_out/r2_lin320.ncu-rep gpurun_out/r2_conv1280s.ncu-rep gpurun_out/r2_conv256.ncu-rep
du -sh gpurun_out; ls gpurun_out | grep r2_
