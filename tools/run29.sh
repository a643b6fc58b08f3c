#!/bin/bash
# round-2 run 29 (1 GPU): the final library of the round (bit-sliced four-parity-row encoder as measured in runs 26-28) — the encode tests of
# the chunk suite on the default route, bench (every timed buffer checked against the reference in the run), pool tests
# (lzgpu_pool_convert_chunks), smoke, ncu --set full of the ec(8,4) kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chunks.py -m gpu -x -q -k "golden or batch_vs_oracle or flat_units or ragged or every_goal or fuzz or many_parity or sweep_chunk_sizes" > gpurun_out/r29_pytest_chunks4.log 2>&1; tail -2 gpurun_out/r29_pytest_chunks4.log
timeout 300 python bench.py > gpurun_out/r29_bench.json 2> gpurun_out/r29_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r29_bench.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "GiB/s frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 1))
    for e in d.get("extra", []):
        if "ec(8,4)" in e["name"] or "ec(5,3) 64" in e["name"]:
            print(" ", e["name"], round(e["frac_of_measured_hbm"], 3), e["parity"][:40])
except Exception as ex:
    print("bench parse failed", ex)
PY
timeout 240 python -m pytest tests/test_pool.py -m gpu -x -q > gpurun_out/r29_pytest_pool.log 2>&1; tail -2 gpurun_out/r29_pytest_pool.log
cat > /tmp/ncu_bs.py <<'PY'
import sys
sys.path.insert(0, '.')
import lizardfs_b200 as L
e = L.Engine(0)
g = L.SliceType("ec(8,4)")
n, nb, B = 32, 1024, 65536
pb = nb // 8
d = e.dev_alloc(n * nb * B); p = e.dev_alloc(n * 4 * pb * B); c = e.dev_alloc(n * (nb + 4 * pb) * 4)
e.fill_chunks_dev(d, n, nb * B, nb * B, 5)
for _ in range(2):
    e.encode_chunks_dev(g, n, nb * B, d, nb * B, p, 4 * pb * B, c, nb + 4 * pb)
e.sync()
PY
timeout 240 ncu --set full --clock-control none --import-source on -k regex:fused_stream_kernel -s 1 -c 1 -o gpurun_out/r29_ec84_bs -f python /tmp/ncu_bs.py > gpurun_out/r29_ncu.log 2>&1; tail -1 gpurun_out/r29_ncu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" > gpurun_out/r29_smoke.log 2>&1; tail -1 gpurun_out/r29_smoke.log
