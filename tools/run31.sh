#!/bin/bash
# round-2 run 31 (1 GPU): cache operator of the parity / image stores — st.global.L1::no_allocate (production) against st.global.cs
# (-DLZ_STG_CS) and plain st.global (-DLZ_STG_PLAIN) on the goals with the largest write share
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
GG='ec(3,2);xor2;ec(4,2);xor3;ec(5,3);ec(8,2);ec(8,4)'
for v in "" _stcs _stplain ""; do
  LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu$v.so timeout 200 python tools/sweep.py --full-size-only --sections enc --goals "$GG" --bytes $((4<<30)) --out gpurun_out/r31_st$v.md > /dev/null 2> gpurun_out/r31_st$v.err
  echo "== liblzgpu$v"; grep -h "^| ec(\|^| xor" gpurun_out/r31_st$v.md | cut -c1-100
done
