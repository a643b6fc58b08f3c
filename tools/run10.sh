#!/bin/bash
# round-2 run 10 (1 GPU): conversion kernel with the rebuild one step ahead; Cauchy encode A/B (fused kernel vs gf_dot + CRC passes); Cauchy degraded read with the levelled multiply
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r10_pytest.log 2>&1; tail -2 gpurun_out/r10_pytest.log
timeout 300 python tools/sweep.py --sections rec,conv --rec 'ec(8,2):1,4' --out gpurun_out/r10_conv_fused.md > /dev/null 2> gpurun_out/r10_conv_fused.err
grep -h "lost" gpurun_out/r10_conv_fused.md | cut -c1-140
CG='ec(8,6);ec(4,5);ec(21,4);ec(16,8);ec(12,5);ec(32,4)'
timeout 300 python tools/sweep.py --full-size-only --sections enc --goals "$CG" --bytes $((4<<30)) --out gpurun_out/r10_cauchy_fused.md > /dev/null 2> gpurun_out/r10_cauchy_fused.err
LZGPU_CAUCHY_FUSED=0 timeout 300 python tools/sweep.py --full-size-only --sections enc --goals "$CG" --bytes $((4<<30)) --out gpurun_out/r10_cauchy_dot.md > /dev/null 2> gpurun_out/r10_cauchy_dot.err
grep -h "^| ec(" gpurun_out/r10_cauchy_fused.md | cut -c1-110
grep -h "^| ec(" gpurun_out/r10_cauchy_dot.md | cut -c1-110
REC='ec(8,6):2;ec(4,5):0,3;ec(21,4):0,20;ec(12,5):3,7'
LZGPU_DIRECT_WIDE=1 timeout 300 python tools/sweep.py --sections rec --rec "$REC" --bytes $((4<<30)) --out gpurun_out/r10_rec_direct.md > /dev/null 2> gpurun_out/r10_rec_direct.err
grep -h "^| ec(" gpurun_out/r10_rec_direct.md | cut -c1-110
