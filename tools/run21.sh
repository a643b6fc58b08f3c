#!/bin/bash
# round-2 run 21 (1 GPU): constant-folded instantiations for ec(8,3), ec(6,4), ec(4,4) — parity tests, then the encode sweep of those goals
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chunks.py -m gpu -x -q -k "encode or every_goal or fuzz or geometry" > gpurun_out/r21_pytest_enc.log 2>&1; tail -2 gpurun_out/r21_pytest_enc.log
timeout 300 python tools/sweep.py --full-size-only --sections enc --goals 'ec(8,3);ec(6,4);ec(4,4);ec(6,3);ec(8,4)' --bytes $((4<<30)) --out gpurun_out/r21_enc.md > /dev/null 2> gpurun_out/r21_enc.err
grep -h "^| ec(" gpurun_out/r21_enc.md | cut -c1-110
