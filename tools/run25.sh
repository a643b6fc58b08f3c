#!/bin/bash
# round-2 run 25 (1 GPU): lzgpu_pool_convert_chunks — pool tests, then the whole GPU suite once more on the final library
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pool.py -m gpu -x -q > gpurun_out/r25_pytest_pool.log 2>&1; tail -3 gpurun_out/r25_pytest_pool.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r25_pytest.log 2>&1; tail -2 gpurun_out/r25_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r25_smoke.log 2>&1; tail -1 gpurun_out/r25_smoke.log
