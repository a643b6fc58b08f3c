#!/bin/bash
# round-2 run 2: new default geometry (8-warp CTAs for M <= 2, one 16-warp CTA for M = 4, narrow generic items), item-width variants,
# big recover geometry, many-parity goals on the TMA path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=$PWD/lizardfs_b200
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; tail -3 gpurun_out/r2_pytest.log
python tools/sweep.py --full-size-only --sections enc,scrub,rec,conv --goals 'xor2;xor3;ec(3,2);ec(5,3);ec(8,2);ec(8,4);ec(4,2);ec(6,2);ec(6,3);ec(21,4);ec(8,6);ec(4,5);ec(31,4);ec(31,3);ec(16,8)' --rec 'ec(8,2):1,4;ec(8,2):0;ec(3,2):0,2;ec(5,3):0,1,4;ec(5,3):1,3;xor3:1;ec(8,4):0,2,5,7' --out gpurun_out/r2_sweep_prod.md > /dev/null 2> gpurun_out/r2_sweep_prod.err
LZGPU_RECOVER_GEO=2 python tools/sweep.py --sections rec --rec 'ec(8,2):1,4;ec(8,2):0;ec(3,2):0,2;ec(5,3):0,1,4;ec(5,3):1,3;xor3:1;ec(8,4):0,2,5,7' --out gpurun_out/r2_sweep_geo2.md > /dev/null 2> gpurun_out/r2_sweep_geo2.err
LZGPU_LIB=$L/liblzgpu_w42.so python tools/sweep.py --full-size-only --sections enc --goals 'ec(8,4);ec(6,4);ec(4,4)' --out gpurun_out/r2_sweep_w42.md > /dev/null 2> gpurun_out/r2_sweep_w42.err
LZGPU_LIB=$L/liblzgpu_w32.so python tools/sweep.py --full-size-only --sections enc --goals 'ec(5,3);ec(6,3);ec(8,3)' --out gpurun_out/r2_sweep_w32.md > /dev/null 2> gpurun_out/r2_sweep_w32.err
python tools/sweep.py --full-size-only --sections enc --goals 'ec(6,4);ec(4,4);ec(8,3)' --out gpurun_out/r2_sweep_prod2.md > /dev/null 2> gpurun_out/r2_sweep_prod2.err
LZGPU_LIB=$L/liblzgpu_noaux.so python tools/sweep.py --full-size-only --sections enc,scrub --goals 'xor2;xor3;ec(3,2);ec(5,3);ec(8,2);ec(8,4)' --out gpurun_out/r2_sweep_noaux.md > /dev/null 2> gpurun_out/r2_sweep_noaux.err
LZGPU_LIB=$L/liblzgpu_gen4.so python tools/sweep.py --full-size-only --sections enc --goals 'ec(21,4);ec(8,6);ec(4,5);ec(16,8)' --out gpurun_out/r2_sweep_gen4.md > /dev/null 2> gpurun_out/r2_sweep_gen4.err
LZGPU_RECOVER_GEO=2 timeout 600 python -m pytest tests/test_gpu_chunks.py tests/test_gpu_replication.py -m gpu -x -q -k "recover or convert or roundtrip or every_goal" > gpurun_out/r2_pytest_geo2.log 2>&1; tail -2 gpurun_out/r2_pytest_geo2.log
for V in w42 w32 gen4 noaux; do LZGPU_LIB=$L/liblzgpu_$V.so timeout 300 python -m pytest tests/test_gpu_chunks.py -m gpu -x -q -k "golden or batch_vs_oracle or flat_units or every_bench_goal or every_goal" > gpurun_out/r2_pytest_$V.log 2>&1; tail -2 gpurun_out/r2_pytest_$V.log; done
NCU="ncu --set full --clock-control none"
$NCU -k regex:fused_stream -s 2 -c 1 -o gpurun_out/r2_prof_ec82 python tools/sweep.py --full-size-only --sections enc --goals 'ec(8,2)' --steps 1 --warmup 2 --out gpurun_out/r2_tmp.md > gpurun_out/r2_ncu_ec82.log 2>&1
$NCU -k regex:fused_stream -s 2 -c 1 -o gpurun_out/r2_prof_ec84 python tools/sweep.py --full-size-only --sections enc --goals 'ec(8,4)' --steps 1 --warmup 2 --out gpurun_out/r2_tmp.md > gpurun_out/r2_ncu_ec84.log 2>&1
LZGPU_RECOVER_GEO=2 $NCU -k regex:fused_recover -s 2 -c 1 -o gpurun_out/r2_prof_rec53 python tools/sweep.py --sections rec --rec 'ec(5,3):0,1,4' --rec-variants plain --steps 1 --warmup 2 --out gpurun_out/r2_tmp.md > gpurun_out/r2_ncu_rec53.log 2>&1
for r in ec82 ec84 rec53; do python tools/ncu_summary.py gpurun_out/r2_prof_$r.ncu-rep "$r" > gpurun_out/r2_prof_$r.md 2>/dev/null; ncu -i gpurun_out/r2_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2_prof_$r.csv 2>/dev/null; done
du -sm gpurun_out; if [ $(du -sm gpurun_out | cut -f1) -gt 40 ]; then rm -f gpurun_out/r2_prof_ec84.ncu-rep gpurun_out/r2_prof_rec53.ncu-rep; fi
timeout 600 python bench.py --steps 20 --cpu-chunks 64 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 600 gpurun_out/r2_bench.err
timeout 300 ./tests/cpp/build/test_link_substitution Benchmark > gpurun_out/r2_linksub_bench.txt 2>&1
ls -la gpurun_out | grep r2_ | head -60
