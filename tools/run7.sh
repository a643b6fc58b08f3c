#!/bin/bash
# round-2 run 7 (8 GPUs): torchrun arm and single-process pool arm at N = 8
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r7_ngpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --cpu-chunks 128 > gpurun_out/r7_bench_n8.json 2> gpurun_out/r7_bench_n8.err; tail -c 300 gpurun_out/r7_bench_n8.err
timeout 600 python bench.py --pool-gpus 8 --steps 6 > gpurun_out/r7_pool_n8.json 2> gpurun_out/r7_pool_n8.err; cut -c1-250 gpurun_out/r7_pool_n8.json
timeout 600 python bench.py --pool-gpus 4 --steps 6 > gpurun_out/r7_pool_n4.json 2> gpurun_out/r7_pool_n4.err
timeout 600 python -m pytest tests/test_pool.py -m gpu -x -q > gpurun_out/r7_pytest_pool.log 2>&1; tail -1 gpurun_out/r7_pytest_pool.log
