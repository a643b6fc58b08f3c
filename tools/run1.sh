#!/bin/bash
# round-2 run 1: warp-slot map, baseline sweep, t256 variant, ncu captures of the kernels VERDICT asks about
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r1_gpu.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; tail -3 gpurun_out/r1_pytest.log
./tools/micro/warpmap > gpurun_out/r1_warpmap.txt 2>&1
python tools/sweep.py --full-size-only --sections enc,rec --rec 'ec(8,2):1,4;ec(3,2):0,2;ec(5,3):0,1,4;xor3:1' --out gpurun_out/r1_sweep_base.md > /dev/null 2> gpurun_out/r1_sweep_base.err
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_t256.so python tools/sweep.py --full-size-only --sections enc --out gpurun_out/r1_sweep_t256.md > /dev/null 2> gpurun_out/r1_sweep_t256.err
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_big34.so python tools/sweep.py --full-size-only --sections enc --goals 'ec(5,3);ec(8,4)' --out gpurun_out/r1_sweep_big34.md > /dev/null 2> gpurun_out/r1_sweep_big34.err
for V in t256 big34; do LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_$V.so timeout 300 python -m pytest tests/test_gpu_chunks.py -m gpu -x -q -k "golden or batch_vs_oracle or flat_units or every_bench_goal" > gpurun_out/r1_pytest_$V.log 2>&1; tail -2 gpurun_out/r1_pytest_$V.log; done
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:fused_stream -s 2 -c 1 -o gpurun_out/r1_prof_ec84 python tools/sweep.py --full-size-only --sections enc --goals 'ec(8,4)' --steps 1 --warmup 2 --out gpurun_out/r1_tmp.md > gpurun_out/r1_ncu_ec84.log 2>&1
$NCU -k regex:fused_stream -s 2 -c 1 -o gpurun_out/r1_prof_ec53 python tools/sweep.py --full-size-only --sections enc --goals 'ec(5,3)' --steps 1 --warmup 2 --out gpurun_out/r1_tmp.md > gpurun_out/r1_ncu_ec53.log 2>&1
$NCU -k regex:fused_recover -s 2 -c 1 -o gpurun_out/r1_prof_rec53 python tools/sweep.py --sections rec --rec 'ec(5,3):0,1,4' --rec-variants plain --steps 1 --warmup 2 --out gpurun_out/r1_tmp.md > gpurun_out/r1_ncu_rec53.log 2>&1
# per-sub-partition instance values for the headline kernel
ncu --clock-control none --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,smsp__cycles_active.sum,smsp__warps_active.sum --print-metric-instances values -k regex:fused_stream -s 2 -c 1 python tools/sweep.py --full-size-only --sections enc --goals 'ec(8,2)' --steps 1 --warmup 2 --out gpurun_out/r1_tmp.md > gpurun_out/r1_smsp_ec82.log 2>&1
ncu --clock-control none --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,smsp__cycles_active.sum --print-metric-instances values -k regex:fused_stream -s 2 -c 1 python tools/sweep.py --full-size-only --sections enc --goals 'ec(8,4)' --steps 1 --warmup 2 --out gpurun_out/r1_tmp.md > gpurun_out/r1_smsp_ec84.log 2>&1
timeout 600 python bench.py --steps 4 --cpu-chunks 32 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; tail -c 600 gpurun_out/r1_bench.err
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_t256.so timeout 300 python bench.py --steps 6 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/r1_bench_t256.json 2> gpurun_out/r1_bench_t256.err
python tools/sweep.py --sections conv --out gpurun_out/r1_sweep_conv.md > /dev/null 2> gpurun_out/r1_sweep_conv.err
ls -la gpurun_out | head -60
