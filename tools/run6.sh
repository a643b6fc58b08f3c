#!/bin/bash
# round-2 run 6 (2 GPUs): chained x4 adopted -> full test suite again; pool arm with NUMA-placed caller buffers; 2 x 10-warp variant
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r6_pytest.log 2>&1; tail -2 gpurun_out/r6_pytest.log
timeout 600 python bench.py --pool-gpus 2 --steps 8 > gpurun_out/r6_pool_n2.json 2> gpurun_out/r6_pool_n2.err; cat gpurun_out/r6_pool_n2.json | cut -c1-200
timeout 600 python bench.py --pool-gpus 2 --steps 8 --e2e-chunks 128 > gpurun_out/r6_pool_n2_128.json 2> gpurun_out/r6_pool_n2_128.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --no-extra --no-cpu-baseline > gpurun_out/r6_bench_n2.json 2> gpurun_out/r6_bench_n2.err
python tools/sweep.py --full-size-only --sections enc --goals 'ec(5,3);ec(6,3);ec(8,3);ec(8,4);ec(6,4);ec(4,4);ec(8,2)' --out gpurun_out/r6_sweep_prod.md > /dev/null 2> gpurun_out/r6_sweep_prod.err
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_t320.so python tools/sweep.py --full-size-only --sections enc --goals 'ec(8,4);ec(6,4);ec(4,4)' --out gpurun_out/r6_sweep_t320.md > /dev/null 2> gpurun_out/r6_sweep_t320.err
LZGPU_LIB=$PWD/lizardfs_b200/liblzgpu_t320.so timeout 300 python -m pytest tests/test_gpu_chunks.py -m gpu -x -q -k "golden or batch_vs_oracle or flat_units or every_bench_goal or every_goal" > gpurun_out/r6_pytest_t320.log 2>&1; tail -1 gpurun_out/r6_pytest_t320.log
ls gpurun_out | wc -l
