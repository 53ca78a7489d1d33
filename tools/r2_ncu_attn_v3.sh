#!/bin/bash
# ncu --set full of the default flash kernel (attention_d64_v3_kernel) at the production shape, one launch
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none -k regex:"attention_d64" -s 2 -c 1 -o gpurun_out/r2_attn_v3 -f python tools/prof_kernels.py attn 3 > gpurun_out/r2_attn_v3.log 2>&1
ncu -i gpurun_out/r2_attn_v3.ncu-rep --page raw --csv > gpurun_out/r2_attn_v3.raw.csv 2>/dev/null
rm -f gpurun_out/r2_attn_v3.ncu-rep
python - <<'P'
import csv
rows = list(csv.reader(open("gpurun_out/r2_attn_v3.raw.csv")))
hdr, unit, val = rows[0], rows[1], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum"]
for i, h in enumerate(hdr):
    if h in want or "shared" in h and "pct" in h or "pipe_xu" in h or "tmem" in h.lower() or "uniform" in h and "pct" in h:
        print(f"{h:95s} {val[i]:>18s} {unit[i]}")
P
