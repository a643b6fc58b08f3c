#!/bin/bash
# round-2 run 26 (1 GPU): the bit-sliced four-parity-row encoder (csrc/bitslice.cuh; default on in this library) — parity tests of the
# new route, A/B against the packed-byte route (LZGPU_BITSLICE=0) and against the builds with the right shifts on the FMA pipe / with the GF warps in the first four warp slots, memcheck
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_bitslice.py -m gpu -x -q > gpurun_out/r26_pytest_bs.log 2>&1; tail -2 gpurun_out/r26_pytest_bs.log
G4='ec(8,4);ec(6,4);ec(4,4);ec(10,4);ec(12,4);ec(7,4);ec(16,4)'
for v in 0 1; do
  LZGPU_BITSLICE=$v timeout 200 python tools/sweep.py --full-size-only --sections enc --goals "$G4" --bytes $((4<<30)) --out gpurun_out/r26_m4_bs$v.md > /dev/null 2> gpurun_out/r26_m4_bs$v.err
  grep -h "^| ec(" gpurun_out/r26_m4_bs$v.md | cut -c1-100
done
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_shrfma.so LZGPU_BITSLICE=1 timeout 200 python tools/sweep.py --full-size-only --sections enc --goals "$G4" --bytes $((4<<30)) --out gpurun_out/r26_m4_shrfma.md > /dev/null 2> gpurun_out/r26_m4_shrfma.err
grep -h "^| ec(" gpurun_out/r26_m4_shrfma.md | cut -c1-100
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_gffirst.so LZGPU_BITSLICE=1 timeout 200 python tools/sweep.py --full-size-only --sections enc --goals "$G4" --bytes $((4<<30)) --out gpurun_out/r26_m4_gffirst.md > /dev/null 2> gpurun_out/r26_m4_gffirst.err
grep -h "^| ec(" gpurun_out/r26_m4_gffirst.md | cut -c1-100
TOOLS=memcheck SAN_TIMEOUT=200 bash tools/sanitize_bitslice.sh > gpurun_out/r26_sanitize.log 2>&1; tail -3 gpurun_out/r26_sanitize.log
