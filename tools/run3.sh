#!/bin/bash
# round-2 run 3: decisions of run 2 applied (eight-byte items for four parity rows, per-launch generic item width, 16-warp recover
# geometry, deferred verification in the timing loops); variants: 16-warp CTA for three parity rows, item cap, 16-byte items for M = 4
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=$PWD/lizardfs_b200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1; tail -3 gpurun_out/r3_pytest.log
python tools/sweep.py --full-size-only --sections enc,scrub,rec,conv --goals 'xor2;xor3;ec(3,2);ec(5,3);ec(8,2);ec(8,4);ec(4,2);ec(6,2);ec(6,3);ec(8,3);ec(4,4);ec(6,4);ec(21,4);ec(8,6);ec(4,5);ec(31,4);ec(31,3);ec(16,8)' --rec 'ec(8,2):1,4;ec(8,2):0;ec(3,2):0,2;ec(5,3):0,1,4;ec(5,3):1,3;xor3:1;ec(8,4):0,2,5,7;ec(8,4):1,6' --out gpurun_out/r3_sweep_prod.md > /dev/null 2> gpurun_out/r3_sweep_prod.err
LZGPU_LIB=$L/liblzgpu_big3.so python tools/sweep.py --full-size-only --sections enc --goals 'ec(5,3);ec(6,3);ec(8,3);ec(4,4);ec(6,4);ec(8,4)' --out gpurun_out/r3_sweep_big3.md > /dev/null 2> gpurun_out/r3_sweep_big3.err
LZGPU_LIB=$L/liblzgpu_cap4.so python tools/sweep.py --full-size-only --sections enc --goals 'ec(4,4);ec(6,4);ec(8,4)' --out gpurun_out/r3_sweep_cap4.md > /dev/null 2> gpurun_out/r3_sweep_cap4.err
LZGPU_LIB=$L/liblzgpu_w44.so python tools/sweep.py --full-size-only --sections enc --goals 'ec(4,4);ec(6,4);ec(8,4)' --out gpurun_out/r3_sweep_w44.md > /dev/null 2> gpurun_out/r3_sweep_w44.err
for V in big3 cap4; do LZGPU_LIB=$L/liblzgpu_$V.so timeout 300 python -m pytest tests/test_gpu_chunks.py -m gpu -x -q -k "golden or batch_vs_oracle or flat_units or every_bench_goal or every_goal" > gpurun_out/r3_pytest_$V.log 2>&1; tail -2 gpurun_out/r3_pytest_$V.log; done
NCU="ncu --set full --clock-control none"
$NCU -k regex:fused_stream -s 2 -c 1 -o gpurun_out/r3_prof_ec84 python tools/sweep.py --full-size-only --sections enc --goals 'ec(8,4)' --steps 1 --warmup 2 --out gpurun_out/r3_tmp.md > gpurun_out/r3_ncu_ec84.log 2>&1
$NCU -k regex:fused_recover -s 2 -c 1 -o gpurun_out/r3_prof_rec53 python tools/sweep.py --sections rec --rec 'ec(5,3):0,1,4' --rec-variants plain --steps 1 --warmup 2 --out gpurun_out/r3_tmp.md > gpurun_out/r3_ncu_rec53.log 2>&1
$NCU -k regex:fused_recover -s 2 -c 1 -o gpurun_out/r3_prof_rec82 python tools/sweep.py --sections rec --rec 'ec(8,2):1,4' --rec-variants full --steps 1 --warmup 2 --out gpurun_out/r3_tmp.md > gpurun_out/r3_ncu_rec82.log 2>&1
for r in ec84 rec53 rec82; do python tools/ncu_summary.py gpurun_out/r3_prof_$r.ncu-rep "$r" > gpurun_out/r3_prof_$r.md 2>/dev/null; done
rm -f gpurun_out/r3_prof_*.ncu-rep
timeout 600 python bench.py --steps 20 --cpu-chunks 64 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; tail -c 600 gpurun_out/r3_bench.err
ls -la gpurun_out | grep r3_ | wc -l
