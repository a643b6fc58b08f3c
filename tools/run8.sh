#!/bin/bash
# round-2 run 8 (1 GPU): DIRECT (Cauchy) form of the fused degraded read — tests, then A/B of the item widths against the generic route
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chunks.py -m gpu -x -q -k "cauchy" > gpurun_out/r8_pytest_cauchy.log 2>&1; tail -3 gpurun_out/r8_pytest_cauchy.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r8_pytest.log 2>&1; tail -2 gpurun_out/r8_pytest.log
REC='ec(8,6):1,4,7;ec(8,6):2;ec(4,5):0,3;ec(21,4):0,20;ec(16,8):2,9,15;ec(32,4):5,6,30,31;ec(12,5):3,7'
for w in -2 0 1; do
  LZGPU_DIRECT_WIDE=$w timeout 300 python tools/sweep.py --sections rec --rec "$REC" --bytes $((4<<30)) --out gpurun_out/r8_rec_w$w.md > /dev/null 2> gpurun_out/r8_rec_w$w.err
done
grep -h "ec(" gpurun_out/r8_rec_w-2.md | cut -c1-110
grep -h "ec(" gpurun_out/r8_rec_w0.md | cut -c1-110
grep -h "ec(" gpurun_out/r8_rec_w1.md | cut -c1-110
