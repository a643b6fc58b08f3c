#!/bin/bash
# round-2 run 13 (1 GPU): conversion kernel with table-driven GF addressing and slow polling in the rebuild warps
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_replication.py -m gpu -x -q > gpurun_out/r13_pytest_repl.log 2>&1; tail -2 gpurun_out/r13_pytest_repl.log
timeout 300 python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --out gpurun_out/r13_conv_fused.md > /dev/null 2> gpurun_out/r13_conv_fused.err
grep -h "lost" gpurun_out/r13_conv_fused.md | cut -c1-140
