#!/bin/bash
# round-2 run 24 (1 GPU): A/B — generic-coefficient (Cauchy) encoder on ONE 16-warp CTA per SM (-DLZ_TGEN=512, 128 registers) against the production two 9-warp CTAs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
CG='ec(8,6);ec(4,5);ec(21,4);ec(16,8);ec(12,5);ec(32,4);ec(31,3)'
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_tgen512.so timeout 600 python -m pytest tests/test_gpu_chunks.py -m gpu -x -q -k "many_parity or every_goal or fuzz" > gpurun_out/r24_pytest_tgen512.log 2>&1; tail -2 gpurun_out/r24_pytest_tgen512.log
timeout 300 python tools/sweep.py --full-size-only --sections enc --goals "$CG" --bytes $((4<<30)) --out gpurun_out/r24_cauchy_prod.md > /dev/null 2> gpurun_out/r24_cauchy_prod.err
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_tgen512.so timeout 300 python tools/sweep.py --full-size-only --sections enc --goals "$CG" --bytes $((4<<30)) --out gpurun_out/r24_cauchy_tgen512.md > /dev/null 2> gpurun_out/r24_cauchy_tgen512.err
grep -h "^| ec(" gpurun_out/r24_cauchy_prod.md | cut -c1-110
grep -h "^| ec(" gpurun_out/r24_cauchy_tgen512.md | cut -c1-110
