#!/bin/bash
# round-2 run 17 (2 GPUs): the torchrun arm of the final library (bench line with the conversion entry under torch.distributed)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r17_bench_n2.json 2> gpurun_out/r17_bench_n2.err; cut -c1-250 gpurun_out/r17_bench_n2.json; tail -3 gpurun_out/r17_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r17_bench_ref_n2.json 2> gpurun_out/r17_bench_ref_n2.err; cut -c1-200 gpurun_out/r17_bench_ref_n2.json
