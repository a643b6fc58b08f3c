#!/bin/bash
# round-2 run 30 (1 GPU): bench of the final library (every timed buffer checked against the reference in the run), ncu --set full of
# the bit-sliced ec(8,4) kernel, the encode sweep of the goals the bit-sliced routes touch
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r30_bench.json 2> gpurun_out/r30_bench.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r30_bench.json").read().strip().splitlines()[-1])
    print("bench", round(d["value"]), "GiB/s frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"], 1))
    for e in d.get("extra", []):
        if "ec(8,4)" in e["name"] or "ec(5,3) 64" in e["name"]:
            print(" ", e["name"], round(e["frac_of_measured_hbm"], 3), e["parity"][:40])
except Exception as ex:
    print("bench parse failed", ex)
PY
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
timeout 240 ncu --set full --clock-control none --import-source on -k regex:fused_stream_kernel -s 1 -c 1 -o gpurun_out/r30_ec84_bs -f python /tmp/ncu_bs.py > gpurun_out/r30_ncu.log 2>&1; tail -1 gpurun_out/r30_ncu.log
GG='xor2;xor3;ec(3,2);ec(5,3);ec(8,2);ec(8,4);ec(6,3);ec(8,3);ec(4,3);ec(9,3);ec(12,3);ec(31,3);ec(4,4);ec(6,4);ec(10,4);ec(12,4);ec(7,4);ec(16,4);ec(20,4)'
timeout 300 python tools/sweep.py --full-size-only --sections enc --goals "$GG" --bytes $((8<<30)) --out gpurun_out/r30_sweep_enc.md > /dev/null 2> gpurun_out/r30_sweep_enc.err; grep -h "^| ec(\|^| xor" gpurun_out/r30_sweep_enc.md | cut -c1-100
