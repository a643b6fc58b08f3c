#!/bin/bash
# round-2 run 14 (1 GPU): the final library — full GPU test suite, smoke, sanitizer over the new kernels, full sweep, bench (both arms), ncu of the conversion kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r14_pytest.log 2>&1; tail -2 gpurun_out/r14_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r14_smoke.log 2>&1; tail -1 gpurun_out/r14_smoke.log
TOOLS="memcheck synccheck" bash tools/sanitize.sh > gpurun_out/r14_sanitize.log 2>&1; tail -6 gpurun_out/r14_sanitize.log
timeout 900 python bench.py > gpurun_out/r14_bench.json 2> gpurun_out/r14_bench.err; cut -c1-400 gpurun_out/r14_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/r14_bench_ref.json 2> gpurun_out/r14_bench_ref.err; cut -c1-300 gpurun_out/r14_bench_ref.json
timeout 900 python tools/sweep.py --out gpurun_out/r14_sweep.md > /dev/null 2> gpurun_out/r14_sweep.err; grep -c "^|" gpurun_out/r14_sweep.md
NCU="ncu --set full --clock-control none"
$NCU -k regex:fused_convert -s 1 -c 1 -o gpurun_out/r14_prof_conv python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --rec-variants full --steps 1 --warmup 1 --out gpurun_out/r14_tmp.md > gpurun_out/r14_ncu_conv.log 2>&1
python tools/ncu_summary.py gpurun_out/r14_prof_conv.ncu-rep "one-pass slice conversion ec(8,2) (parts 1, 4 lost) -> ec(3,2), 64 chunks" "python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --steps 1 --warmup 1" > gpurun_out/r14_prof_conv.md 2>/dev/null
rm -f gpurun_out/r14_prof_conv.ncu-rep
ls gpurun_out | wc -l
