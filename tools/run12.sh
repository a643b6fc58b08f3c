#!/bin/bash
# round-2 run 12 (1 GPU): ncu --set full of the one-pass conversion kernel (ec(8,2), parts 1 and 4 lost -> all ec(3,2) parts)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:fused_convert -s 1 -c 1 -o gpurun_out/r12_prof_conv python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --rec-variants full --steps 1 --warmup 1 --out gpurun_out/r12_tmp.md > gpurun_out/r12_ncu_conv.log 2>&1
python tools/ncu_summary.py gpurun_out/r12_prof_conv.ncu-rep "one-pass slice conversion ec(8,2) (parts 1, 4 lost) -> ec(3,2), 64 chunks" > gpurun_out/r12_prof_conv.md 2>/dev/null
ncu -i gpurun_out/r12_prof_conv.ncu-rep --page source --csv > gpurun_out/r12_prof_conv_source.csv 2>/dev/null
ls -la gpurun_out | tail -8
head -40 gpurun_out/r12_prof_conv.md
