#!/bin/bash
# round-2 run 23 (1 GPU): compile-time k = 4, 6 degraded reads; conversion route check before the status slot — full GPU tests, A/B sweep, memcheck
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r23_pytest.log 2>&1; tail -2 gpurun_out/r23_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r23_smoke.log 2>&1; tail -1 gpurun_out/r23_smoke.log
REC='ec(4,2):0,3;ec(6,2):1,4;ec(6,3):0,2,5;ec(6,3):1,4'
for v in 0 1; do
  LZGPU_RECOVER_K3=$v timeout 300 python tools/sweep.py --sections rec --rec "$REC" --out gpurun_out/r23_rec_k_$v.md > /dev/null 2> gpurun_out/r23_rec_k_$v.err
  grep -h "^| ec(" gpurun_out/r23_rec_k_$v.md | cut -c1-120
done
TOOLS="memcheck" bash tools/sanitize.sh > gpurun_out/r23_sanitize.log 2>&1; tail -2 gpurun_out/r23_sanitize.log
