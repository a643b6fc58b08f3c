#!/bin/bash
# round-2 run 22 (1 GPU): the final library after folding more goals — full GPU test suite, smoke, the "more goals" sweep, memcheck
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r22_pytest.log 2>&1; tail -2 gpurun_out/r22_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r22_smoke.log 2>&1; tail -1 gpurun_out/r22_smoke.log
timeout 600 python tools/sweep.py --full-size-only --sections enc,rec --goals 'xor4;ec(4,2);ec(5,2);ec(6,2);ec(10,2);ec(4,3);ec(6,3);ec(8,3);ec(4,4);ec(6,4);ec(10,4);ec(12,4);ec(8,6);ec(4,5);ec(21,4);ec(16,8);ec(31,3)' --rec 'ec(3,2):1;ec(5,3):1,3;ec(5,3):2,3,4;ec(8,4):0,2,5,7;ec(8,6):2;ec(4,5):0,3;ec(21,4):0,20' --bytes $((4<<30)) --out gpurun_out/r22_sweep_more.md > /dev/null 2> gpurun_out/r22_sweep_more.err; grep "^| ec\|^| xor" gpurun_out/r22_sweep_more.md | head -17 | cut -c1-110
TOOLS="memcheck" bash tools/sanitize.sh > gpurun_out/r22_sanitize.log 2>&1; tail -3 gpurun_out/r22_sanitize.log
