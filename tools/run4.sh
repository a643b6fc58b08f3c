#!/bin/bash
# round-2 run 4: the final library — full GPU test suite, smoke, the complete sweep, bench (both arms), ncu launch list, sanitizer
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4_pytest.log 2>&1; tail -3 gpurun_out/r4_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1; tail -1 gpurun_out/r4_smoke.log
python tools/sweep.py --out gpurun_out/r4_sweep_full.md > /dev/null 2> gpurun_out/r4_sweep_full.err
python tools/sweep.py --full-size-only --sections enc,rec --goals 'ec(4,2);ec(6,2);ec(6,3);ec(8,3);ec(4,4);ec(6,4);ec(21,4);ec(8,6);ec(4,5);ec(31,4);ec(31,3);ec(16,8);ec(32,32)' --rec 'ec(5,3):1,3;ec(8,4):0,2,5,7;ec(8,4):1,6;ec(8,2):0' --out gpurun_out/r4_sweep_more.md > /dev/null 2> gpurun_out/r4_sweep_more.err
timeout 900 python bench.py --steps 20 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; tail -c 400 gpurun_out/r4_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r4_bench_ref.json 2> gpurun_out/r4_bench_ref.err
timeout 600 python bench.py > gpurun_out/r4_bench_default.json 2> gpurun_out/r4_bench_default.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r4_ncu_bench.log 2>&1
TOOLS="memcheck synccheck" bash tools/sanitize.sh > gpurun_out/r4_sanitize.txt 2>&1; tail -6 gpurun_out/r4_sanitize.txt
ls gpurun_out | wc -l; du -sm gpurun_out
