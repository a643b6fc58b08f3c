#!/bin/bash
# round-2 run 33 (1 GPU, the last GPU minutes of the round): GF warp count of the bit-sliced kernels as a function of G (up to 8 GF
# warps: narrow stripes get larger units) — the chunk / bit-slice / conversion suites on the default route, then A/B against 4 GF warps
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_bitslice.py tests/test_gpu_chunks.py tests/test_gpu_replication.py -m gpu -x -q > gpurun_out/r33_pytest.log 2>&1; tail -2 gpurun_out/r33_pytest.log
GG='ec(5,3);ec(6,3);ec(4,3);ec(4,4);ec(6,4)'
for w in 8 4; do
  LZGPU_BITSLICE=7 LZGPU_BS_GFW=$w timeout 60 python tools/sweep.py --full-size-only --sections enc --goals "$GG" --bytes $((4<<30)) --out gpurun_out/r33_enc_gfw$w.md > /dev/null 2> gpurun_out/r33_enc_gfw$w.err
  echo "== enc, GF warps <= $w"; grep -h "^| ec(" gpurun_out/r33_enc_gfw$w.md | cut -c1-100
done
REC='ec(5,3):0,1,4;ec(6,3):0,2,5;ec(8,3):1,4,6'
for w in 8 4; do
  LZGPU_BS_GFW=$w timeout 60 python tools/sweep.py --sections rec --rec "$REC" --bytes $((4<<30)) --out gpurun_out/r33_rec_gfw$w.md > /dev/null 2> gpurun_out/r33_rec_gfw$w.err
  echo "== rec, GF warps <= $w"; grep -h "^| ec(" gpurun_out/r33_rec_gfw$w.md | cut -c1-120
done
