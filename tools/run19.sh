#!/bin/bash
# round-2 run 19 (1 GPU): compile-time k = 5 instantiations of the degraded read on the 16-warp geometry (ec(5,3)) against the runtime-k ones
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chunks.py tests/test_gpu_replication.py -m gpu -x -q -k "recover or fuzz or every_goal or convert or roundtrip" > gpurun_out/r19_pytest_rec.log 2>&1; tail -2 gpurun_out/r19_pytest_rec.log
REC='ec(5,3):1,3;ec(5,3):0,1,4;ec(5,3):2,3,4'
for v in 0 1; do
  LZGPU_RECOVER_K3=$v timeout 300 python tools/sweep.py --sections rec --rec "$REC" --out gpurun_out/r19_rec_k5_$v.md > /dev/null 2> gpurun_out/r19_rec_k5_$v.err
  grep -h "^| ec(\|^| xor" gpurun_out/r19_rec_k5_$v.md | cut -c1-120
done
