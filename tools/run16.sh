#!/bin/bash
# round-2 run 16 (1 GPU): the final library — full GPU test suite, smoke, bench (with the conversion entry), sweep, ncu of the adopted conversion kernel, launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r16_pytest.log 2>&1; tail -2 gpurun_out/r16_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r16_smoke.log 2>&1; tail -1 gpurun_out/r16_smoke.log
timeout 900 python bench.py --steps 20 > gpurun_out/r16_bench.json 2> gpurun_out/r16_bench.err; cut -c1-300 gpurun_out/r16_bench.json; tail -3 gpurun_out/r16_bench.err
timeout 900 python tools/sweep.py --out gpurun_out/r16_sweep.md > /dev/null 2> gpurun_out/r16_sweep.err; grep "ec(3,2): all\|xor3: all" gpurun_out/r16_sweep.md
NCU="ncu --set full --clock-control none"
$NCU -k regex:fused_convert -s 1 -c 1 -o gpurun_out/r16_prof_conv python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --rec-variants full --steps 1 --warmup 1 --out gpurun_out/r16_tmp.md > gpurun_out/r16_ncu_conv.log 2>&1
python tools/ncu_summary.py gpurun_out/r16_prof_conv.ncu-rep "one-pass slice conversion ec(8,2) (parts 1, 4 lost) -> ec(3,2), 64 chunks" "python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --steps 1 --warmup 1" > gpurun_out/r16_prof_conv.md 2>/dev/null
rm -f gpurun_out/r16_prof_conv.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r16_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r16_ncu_bench.log 2>&1
TOOLS="memcheck" bash tools/sanitize.sh > gpurun_out/r16_sanitize.log 2>&1; tail -3 gpurun_out/r16_sanitize.log
ls gpurun_out | wc -l
