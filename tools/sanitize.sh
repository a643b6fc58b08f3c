# compute-sanitizer passes over a small encode + recover + scrub through the C ABI (run under gpurun).
# Full reports land in gpurun_out/sanitize_<tool>.log.
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
import lizardfs_b200 as L
from tests import _oracle as O
o = O.load_oracle()
eng = L.Engine(0)
for text, nblk in [("ec(8,2)", 24), ("ec(3,2)", 7), ("xor3", 9)]:
    g = L.SliceType(text)
    data = np.stack([O.fill_chunk(o, nblk * 65536, 3, c) for c in range(2)])
    par, crc = eng.encode_chunks(g, data)
    for c in range(2):
        p_ref, c_ref = o.encode_chunk(g.kind, g.k, g.m, data[c])
        assert (par[c] == p_ref).all() and (crc[c] == c_ref).all()
    parts = [np.stack([O.split_parts(data[c], g.k)[0][j] for c in range(2)]) for j in range(g.k)] + [np.ascontiguousarray(par[:, r]) for r in range(g.m)]
    avail = [None if i == 1 else parts[i] for i in range(g.k + g.m)]
    out, img = eng.recover_chunks(g, nblk, avail, chunk_image=True)
    assert (out[1] == parts[1]).all() and (img == data).all()
eng.verify_blocks(data.reshape(-1), eng.crc_blocks(data.reshape(-1)))
print("sanitizer case OK")
PY
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 8 python /tmp/san_case.py > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "SUMMARY|sanitizer case OK|Error|error" gpurun_out/sanitize_$tool.log | head -5
done
