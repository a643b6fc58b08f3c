#!/bin/bash
# round-2 run 5 (2 GPUs): the pool on two real devices, the torchrun arm at N = 2, the single-process pool arm at N = 2,
# chained-x4 experiment
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r5_gpus.txt
timeout 600 python -m pytest tests/test_pool.py -m gpu -x -q > gpurun_out/r5_pytest_pool.log 2>&1; tail -2 gpurun_out/r5_pytest_pool.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 > gpurun_out/r5_bench_n2.json 2> gpurun_out/r5_bench_n2.err; tail -c 300 gpurun_out/r5_bench_n2.err
timeout 600 python bench.py --pool-gpus 2 --steps 8 > gpurun_out/r5_pool_n2.json 2> gpurun_out/r5_pool_n2.err
timeout 600 python bench.py --pool-gpus 1 --steps 8 > gpurun_out/r5_pool_n1.json 2> gpurun_out/r5_pool_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r5_bench_ref_n2.json 2> gpurun_out/r5_bench_ref_n2.err
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_x4c.so python tools/sweep.py --full-size-only --sections enc --goals 'ec(5,3);ec(6,3);ec(8,4);ec(4,4)' --out gpurun_out/r5_sweep_x4c.md > /dev/null 2> gpurun_out/r5_sweep_x4c.err
python tools/sweep.py --full-size-only --sections enc --goals 'ec(5,3);ec(6,3);ec(8,4);ec(4,4)' --out gpurun_out/r5_sweep_prod.md > /dev/null 2> gpurun_out/r5_sweep_prod.err
ls gpurun_out | wc -l
