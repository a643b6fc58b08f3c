#!/bin/bash
# round-2 run 9 (1 GPU): one-pass slice conversion (convert_kernel.cuh) — tests, then timing against the two-pass route
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_replication.py -m gpu -x -q -k "one_pass" > gpurun_out/r9_pytest_conv.log 2>&1; tail -3 gpurun_out/r9_pytest_conv.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r9_pytest.log 2>&1; tail -2 gpurun_out/r9_pytest.log
timeout 300 python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --out gpurun_out/r9_conv_fused.md > /dev/null 2> gpurun_out/r9_conv_fused.err
LZGPU_CONVERT_FUSED=0 timeout 300 python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --out gpurun_out/r9_conv_two.md > /dev/null 2> gpurun_out/r9_conv_two.err
grep -h "lost" gpurun_out/r9_conv_fused.md | cut -c1-140
grep -h "lost" gpurun_out/r9_conv_two.md | cut -c1-140
