#!/bin/bash
# round-2 run 27 (1 GPU): the bit-sliced geometry for THREE parity rows as well (LZGPU_BITSLICE bit 1) — parity tests of both bit-sliced
# routes, A/B of the three-row goals against the packed-byte route (two 8-warp CTAs), memcheck
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_bitslice.py -m gpu -x -q > gpurun_out/r27_pytest_bs.log 2>&1; tail -2 gpurun_out/r27_pytest_bs.log
G3='ec(5,3);ec(6,3);ec(8,3);ec(4,3);ec(9,3);ec(12,3);ec(31,3);ec(8,4)'
for v in 0 3; do
  LZGPU_BITSLICE=$v timeout 200 python tools/sweep.py --full-size-only --sections enc --goals "$G3" --bytes $((4<<30)) --out gpurun_out/r27_m3_bs$v.md > /dev/null 2> gpurun_out/r27_m3_bs$v.err
  grep -h "^| ec(" gpurun_out/r27_m3_bs$v.md | cut -c1-100
done
TOOLS=memcheck SAN_TIMEOUT=200 bash tools/sanitize_bitslice.sh > gpurun_out/r27_sanitize.log 2>&1; tail -3 gpurun_out/r27_sanitize.log
