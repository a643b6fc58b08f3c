#!/bin/bash
# round-2 run 32 (1 GPU): degraded read with three lost data parts on bit planes (bs_recover_kernel.cuh, default on) — parity tests of
# both routes, the recover / conversion tests of the suite on the default route, A/B sweep, memcheck
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_bitslice.py -m gpu -x -q -k "three_lost" > gpurun_out/r32_pytest_bs.log 2>&1; tail -3 gpurun_out/r32_pytest_bs.log
REC='ec(5,3):0,1,4;ec(5,3):2,3,4;ec(6,3):0,2,5;ec(8,3):1,4,6;ec(8,4):0,2,5;ec(12,3):0,5,11'
for v in 0 1; do
  LZGPU_BS_RECOVER=$v timeout 120 python tools/sweep.py --sections rec --rec "$REC" --bytes $((4<<30)) --out gpurun_out/r32_rec_bs$v.md > /dev/null 2> gpurun_out/r32_rec_bs$v.err
  echo "== LZGPU_BS_RECOVER=$v"; grep -h "^| ec(" gpurun_out/r32_rec_bs$v.md | cut -c1-120
done
timeout 200 python -m pytest tests/test_gpu_chunks.py tests/test_gpu_replication.py -m gpu -x -q -k "recover or fuzz or every_goal or convert" > gpurun_out/r32_pytest_rec.log 2>&1; tail -2 gpurun_out/r32_pytest_rec.log
TOOLS=memcheck SAN_TIMEOUT=150 bash tools/sanitize_bitslice.sh > gpurun_out/r32_sanitize.log 2>&1; tail -3 gpurun_out/r32_sanitize.log
