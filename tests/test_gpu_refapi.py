"""GPU parity tests of the reference-shaped entry points (through the C ABI), mirroring
src/common/reed_solomon_unittest.cc:136-199, crc_unittest.cc:27-63, block_xor_unittest.cc:24-35."""
import numpy as np
import pytest

import lizardfs_b200 as L

pytestmark = pytest.mark.gpu
BLOCK = 65536


def rnd(n, seed):
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8)


def test_mycrc32_known_answers():
    L.mycrc32_init()
    kat = {1: 0xE8B7BE43, 2: 0x78A19D7, 4: 0xAD98E545, 8: 0xBF848046, 16: 0xCFD668D5, 32: 0xCAB11777, 64: 0x89B46555}
    for n, want in kat.items():
        assert L.mycrc32(0, np.full(n, ord("a"), dtype=np.uint8)) == want
    assert L.mycrc32(0, np.zeros(BLOCK, dtype=np.uint8)) == L.mycrc32_zeroblock(0, BLOCK) == 0xD7978EEB


def test_mycrc32_combine_like_reference_test():
    data = (np.arange(BLOCK) & 0xff).astype(np.uint8)
    crc = L.mycrc32(0, data)
    length = 2
    while length < BLOCK:
        for off in (-1, 0, 1):
            n = length + off
            assert L.mycrc32_combine(L.mycrc32(0, data[: BLOCK - n]), L.mycrc32(0, data[BLOCK - n:]), n) == crc
        length *= 2


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 15, 16, 17, 127, 128, 129, 1000, 4095, 65535, 65536, 65537, 200001, 1 << 20])
def test_mycrc32_any_length(oracle, n):
    buf = rnd(n, n)
    assert L.mycrc32(0, buf) == oracle.crc32(0, buf)
    assert L.mycrc32(0xDEADBEEF, buf) == oracle.crc32(0xDEADBEEF, buf)
    assert L.mycrc32_zeroexpanded(0, buf, 777) == oracle.crc32(0, np.concatenate([buf, np.zeros(777, dtype=np.uint8)]))


def test_blockxor_alignment_sweep():
    # block_xor_unittest.cc:24-35 sweeps misalignments; here also the values are checked
    base_d, base_s = rnd(BLOCK + 64, 1), rnd(BLOCK + 64, 2)
    for off_d in (0, 1, 3, 16):
        for off_s in (0, 2, 5):
            for size in (1, 15, 16, 17, 1000, BLOCK):
                d = base_d[off_d:off_d + size].copy()
                s = base_s[off_s:off_s + size].copy()
                want = d ^ s
                L.blockXor(d, s)
                assert (d == want).all()


@pytest.mark.parametrize("erase", [(0, 2), (0, 5), (4, 5)])
def test_rs_recovery(oracle, erase):
    k, m, size = 4, 2, 10000
    rs = L.ReedSolomon(k, m)
    data = [rnd(size, 50 + i) for i in range(k)]
    parity = rs.encode(data)
    want = oracle.rs_encode(k, m, data, size)
    assert all((a == b).all() for a, b in zip(parity, want))
    parts = data + parity
    erased = [1 if i in erase else 0 for i in range(k + m)]
    out = rs.recover(parts, erased)
    for i in erase:
        assert (out[i] == parts[i]).all()


def test_rs_recovery_with_zero_data(oracle):
    k, m, size = 8, 2, 65536
    rs = L.ReedSolomon(k, m)
    data = [rnd(size, 7 + i) if i % 3 else None for i in range(k)]
    dense = [d if d is not None else np.zeros(size, dtype=np.uint8) for d in data]
    parity = rs.encode(data, size)
    want = oracle.rs_encode(k, m, dense, size)
    assert all((a == b).all() for a, b in zip(parity, want))
    erased = [0] * (k + m)
    erased[1] = erased[4] = 1
    out = rs.recover(data + parity, erased, data_size=size)
    assert (out[1] == dense[1]).all() and (out[4] == dense[4]).all()
    # only one of the erased parts requested (NULL output = skip, reed_solomon.h:98-102)
    out = rs.recover(data + parity, erased, wanted=[0, 0, 0, 0, 1, 0, 0, 0, 0, 0], data_size=size)
    assert out[1] is None and (out[4] == dense[4]).all()


@pytest.mark.parametrize("k,m", [(2, 1), (3, 2), (5, 3), (8, 4), (4, 5), (21, 4), (32, 32)])
def test_rs_all_shapes_vs_oracle(oracle, k, m):
    size = 4099  # not a multiple of 16
    rng = np.random.default_rng(k * 100 + m)
    rs = L.ReedSolomon(k, m)
    data = [rnd(size, 1000 + i) for i in range(k)]
    parity = rs.encode(data)
    want = oracle.rs_encode(k, m, data, size)
    assert all((a == b).all() for a, b in zip(parity, want))
    parts = data + parity
    for _ in range(3):
        erased = np.zeros(k + m, dtype=np.uint8)
        erased[rng.choice(k + m, size=m, replace=False)] = 1
        got = rs.recover(parts, erased.tolist())
        ref = oracle.rs_recover(k, m, [None if erased[i] else parts[i] for i in range(k + m)], erased.tolist(), erased.tolist(), size)
        for i in range(k + m):
            if erased[i]:
                assert (got[i] == ref[i]).all() and (got[i] == parts[i]).all()


def test_isal_ec_encode_data(oracle):
    rng = np.random.default_rng(77)
    for srcs, dests, ln in [(1, 1, 1), (3, 2, 100), (8, 2, 65536), (10, 5, 4097), (32, 7, 1234)]:
        coeffs = rng.integers(0, 256, size=(dests, srcs), dtype=np.uint8)
        tables = L.ec_init_tables(srcs, dests, coeffs)
        assert (tables == oracle.init_tables(coeffs)).all()
        src = [rnd(ln, 5 + i) for i in range(srcs)]
        got = L.ec_encode_data(ln, tables, src, dests)
        want = oracle.ec_encode_data(tables, src, dests)
        assert all((a == b).all() for a, b in zip(got, want))


def test_crc_disabled_build_mode(oracle):
    """The reference's ENABLE_CRC switch (src/common/crc.cc:28-41): without it mycrc32 / mycrc32_combine return 0xFEDCBA98.
    lzgpu_set_crc_enabled(0) is that build mode: scalar calls return the constant, batched calls emit it for every block and
    accept exactly it as a stored CRC; parity bytes are unaffected."""
    lib = L._lib.load()
    eng = L.Engine(0)
    try:
        lib.lzgpu_set_crc_enabled(0)
        assert lib.lzgpu_crc_enabled() == 0
        data = np.random.default_rng(7).integers(0, 256, size=(2, 16 * 65536), dtype=np.uint8)
        assert L.mycrc32(0, data[0, :1000]) == 0xFEDCBA98 and L.mycrc32_combine(1, 2, 3) == 0xFEDCBA98
        goal = L.SliceType("ec(3,2)")
        parity, crc = eng.encode_chunks(goal, data)
        assert (crc == 0xFEDCBA98).all()
        p_ref, _ = oracle.encode_chunk(1, 3, 2, data[0])
        assert (parity[0] == p_ref).all()
        assert (eng.crc_blocks(data[0].reshape(-1, 65536)) == 0xFEDCBA98).all()
        # degraded read: stored CRCs equal to the constant pass, anything else is a mismatch
        from tests import _oracle as O
        pb = 6
        per = [O.split_parts(data[c], 3)[0] for c in range(2)]
        parts = [np.stack([per[c][j] for c in range(2)]) for j in range(3)] + [np.ascontiguousarray(parity[:, r]) for r in range(2)]
        good = [np.full((2, pb), 0xFEDCBA98, dtype=np.uint32) for _ in range(5)]
        avail = [None, parts[1], parts[2], parts[3], None]
        out, _ = eng.recover_chunks(goal, 16, avail, part_crc=[None, good[1], good[2], good[3], None])
        assert (out[0] == parts[0]).all()
        bad = [g.copy() for g in good]
        bad[2][1, 4] = 0x12345678
        with pytest.raises(L.ChunkCrcError) as e:
            eng.recover_chunks(goal, 16, avail, part_crc=[None, bad[1], bad[2], bad[3], None])
        assert e.value.where == (1, 2, 4)
    finally:
        lib.lzgpu_set_crc_enabled(1)
        eng.close()
    assert L.mycrc32(0, np.frombuffer(b"a", dtype=np.uint8)) == 0xE8B7BE43
