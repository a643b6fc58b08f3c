#!/bin/bash
# round-2 run 28 (1 GPU): run-time stage count of the bit-sliced kernels (as many stages as fit: narrow stripes get a deeper ring) —
# parity tests, then A/B: 4 stages (run 27's geometry) / up to 8 in 200 KB / up to 8 in 224 KB
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bitslice.py -m gpu -x -q > gpurun_out/r28_pytest_bs.log 2>&1; tail -1 gpurun_out/r28_pytest_bs.log
LZGPU_BS_SMEM_KB=224 timeout 300 python -m pytest tests/test_gpu_bitslice.py -m gpu -x -q -k "every_vandermonde or flat_units or ragged" > gpurun_out/r28_pytest_bs224.log 2>&1; tail -1 gpurun_out/r28_pytest_bs224.log
GG='ec(5,3);ec(6,3);ec(4,3);ec(8,3);ec(9,3);ec(8,4);ec(6,4);ec(4,4);ec(10,4);ec(12,4);ec(7,4)'
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 200 python tools/sweep.py --full-size-only --sections enc --goals "$GG" --bytes $((4<<30)) --out gpurun_out/r28_$label.md > /dev/null 2> gpurun_out/r28_$label.err
  echo "== $label"; grep -h "^| ec(" gpurun_out/r28_$label.md | cut -c1-100
}
run st4 LZGPU_BS_STAGES=4
run st8_200 LZGPU_BS_STAGES=8
run st8_224 LZGPU_BS_STAGES=8 LZGPU_BS_SMEM_KB=224
run st6_224 LZGPU_BS_STAGES=6 LZGPU_BS_SMEM_KB=224
