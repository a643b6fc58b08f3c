#!/bin/bash
# round-2 run 29 (1 GPU): the whole GPU suite on the final library of the round (bit-sliced routes on by default: four parity rows,
# three parity rows with k >= 7), smoke
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r29_pytest.log 2>&1; tail -12 gpurun_out/r29_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r29_smoke.log 2>&1; tail -1 gpurun_out/r29_smoke.log
