#!/bin/bash
# round-2 run 15 (1 GPU): conversion kernel with the destination stripe walk unrolled for three data parts (xor3 / ec(3,2) destinations)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_replication.py -m gpu -x -q > gpurun_out/r15_pytest_repl.log 2>&1; tail -2 gpurun_out/r15_pytest_repl.log
timeout 300 python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --out gpurun_out/r15_conv.md > /dev/null 2> gpurun_out/r15_conv.err
grep -h "lost" gpurun_out/r15_conv.md | cut -c1-140
