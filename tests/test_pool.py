"""lzgpu_pool (several GPUs behind one process, include/lzgpu.h) and the threading contract of a context: the mount runs ten
write workers in one process (reference src/mount/lizard_client.h:77, writedata.cc:645), so the engine is driven from many
threads at once.  CPU part: the share arithmetic and the loud failure without a device.  GPU part: a pool of two contexts
(the same device listed twice when the box has one GPU, two devices when it has more), ten threads on one context."""
import threading

import numpy as np
import pytest

import lizardfs_b200 as L
from tests import _oracle as O

BLOCK = 65536


def test_pool_share_covers_every_chunk_once():
    for n in (0, 1, 2, 7, 8, 9, 63, 64, 4096):
        for g in (1, 2, 3, 4, 8):
            runs = [L.Pool.share(n, g, i) for i in range(g)]
            assert sum(c for _, c in runs) == n
            nxt = 0
            for first, count in runs:
                if count:
                    assert first == nxt
                    nxt = first + count
            # batch b -> device b: shares are equal except the last non-empty one
            sizes = [c for _, c in runs if c]
            assert all(s == sizes[0] for s in sizes[:-1])


def test_pool_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(L.LzGpuError) as e:
        L.Pool()
    assert e.value.status == L._lib.ERR_NO_DEVICE


def _rnd(shape, seed):
    return np.random.default_rng(seed).integers(0, 256, size=shape, dtype=np.uint8)


def _devices():
    import torch
    return [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]


@pytest.mark.gpu
def test_pool_encode_and_recover_vs_oracle(oracle):
    pool = L.Pool(_devices())
    assert len(pool) == 2
    goal = L.SliceType("ec(8,2)")
    nblocks, n = 24, 7                      # 7 chunks over 2 devices: shares of 4 and 3
    data = _rnd((n, nblocks * BLOCK), 77)
    parity, crc = pool.encode_chunks(goal, data)
    for c in range(n):
        p_ref, c_ref = oracle.encode_chunk(goal.kind, goal.k, goal.m, data[c])
        assert (parity[c] == p_ref).all() and (crc[c] == c_ref).all(), c
    st = pool.stats()
    assert st["chunks_encoded"] == n and st["kernel_launches"] >= 2 and st["batches_timed"] >= 1
    # degraded read through the pool: data parts 1 and 4 lost, stored CRCs verified, chunk image
    pb = 3
    per = [O.split_parts(data[c], 8)[0] for c in range(n)]
    parts = [np.stack([per[c][j] for c in range(n)]) for j in range(8)] + [np.ascontiguousarray(parity[:, r]) for r in range(2)]
    pcrc = []
    for j in range(8):
        cj = np.full((n, pb), 0xD7978EEB, dtype=np.uint32)
        cj[:, :] = crc[:, j:nblocks:8]
        pcrc.append(cj)
    pcrc += [np.ascontiguousarray(crc[:, nblocks + r * pb: nblocks + (r + 1) * pb]) for r in range(2)]
    avail = [None if i in (1, 4) else parts[i] for i in range(10)]
    out, img = pool.recover_chunks(goal, nblocks, avail, part_crc=[None if a is None else pcrc[i] for i, a in enumerate(avail)], chunk_image=True)
    assert (out[1] == parts[1]).all() and (out[4] == parts[4]).all() and (img == data).all()
    # a corrupt block in chunk 5 (second device's share): reported with its index in the whole batch
    bad = [None if a is None else a.copy() for a in avail]
    bad[6][5, 2 * BLOCK + 17] ^= 0x40
    with pytest.raises(L.ChunkCrcError) as e:
        pool.recover_chunks(goal, nblocks, bad, part_crc=[None if a is None else pcrc[i] for i, a in enumerate(avail)])
    assert e.value.where == (5, 6, 2)
    # replication through the pool (lzgpu_pool_convert_chunks): the same degraded slice -> every ec(3,2) part + CRCs, the shares of
    # the two devices stitched back in chunk order; a corrupt block again reported with its index in the whole batch
    g32 = L.SliceType("ec(3,2)")
    acrc = [None if a is None else pcrc[i] for i, a in enumerate(avail)]
    conv, ccrc = pool.convert_chunks(goal, g32, nblocks, avail, [1] * 5, part_crc=acrc)
    p32, c32 = pool.encode_chunks(g32, data)
    per32 = [O.split_parts(data[c], 3)[0] for c in range(n)]
    pb32 = nblocks // 3
    for j in range(3):
        assert (conv[j] == np.stack([per32[c][j] for c in range(n)])).all(), j
        assert (ccrc[j] == c32[:, j:nblocks:3]).all(), j
    for r in range(2):
        assert (conv[3 + r] == p32[:, r]).all() and (ccrc[3 + r] == c32[:, nblocks + r * pb32: nblocks + (r + 1) * pb32]).all(), r
    with pytest.raises(L.ChunkCrcError) as e:
        pool.convert_chunks(goal, g32, nblocks, bad, [1] * 5, part_crc=acrc)
    assert e.value.where == (5, 6, 2)
    blocks = _rnd((37, BLOCK), 5)
    got = pool.crc_blocks(blocks)
    import zlib
    assert got.tolist() == [zlib.crc32(blocks[i].tobytes()) for i in range(37)]
    pool.close()


@pytest.mark.gpu
def test_ten_threads_on_one_context(oracle):
    """Ten threads (the mount's default number of write workers) call the device-pointer and the host-pointer entry points of ONE
    context at the same time, each on its own stream, with different goals; every result must be exactly the oracle's."""
    import torch
    eng = L.Engine(0)
    dev = torch.device("cuda", 0)
    goals = ["ec(8,2)", "ec(3,2)", "xor3", "ec(5,3)", "ec(8,4)", "ec(4,5)", "xor2", "ec(6,2)", "ec(8,2)", "ec(3,2)"]
    errors = []

    def worker(t):
        try:
            goal = L.SliceType(goals[t])
            k, m = goal.k, goal.m
            nblocks = 8 + 3 * t
            n = 3
            pb = -(-nblocks // k)
            data = _rnd((n, nblocks * BLOCK), 1000 + t)
            ref = [oracle.encode_chunk(goal.kind, k, m, data[c]) for c in range(n)]
            stream = torch.cuda.Stream(dev)
            d_data = torch.from_numpy(data).to(dev)
            d_par = torch.empty(n * m * pb * BLOCK, dtype=torch.uint8, device=dev)
            d_crc = torch.empty(n * (nblocks + m * pb), dtype=torch.int32, device=dev)
            torch.cuda.synchronize(dev)
            for rep in range(6):
                if rep % 2 == 0:
                    d_par.zero_(); d_crc.zero_()
                    torch.cuda.synchronize(dev)
                    eng.encode_chunks_dev(goal, n, nblocks * BLOCK, d_data.data_ptr(), nblocks * BLOCK, d_par.data_ptr(), m * pb * BLOCK,
                                          d_crc.data_ptr(), nblocks + m * pb, stream=stream.cuda_stream)
                    stream.synchronize()
                    par = d_par.cpu().numpy().reshape(n, m, pb * BLOCK)
                    crc = d_crc.cpu().numpy().view(np.uint32).reshape(n, -1)
                else:
                    par, crc = eng.encode_chunks(goal, data)
                for c in range(n):
                    assert (par[c] == ref[c][0]).all(), (t, rep, c)
                    assert (crc[c] == ref[c][1]).all(), (t, rep, c)
                # degraded read of the first data part with verification (takes a result slot; generic or fused route by goal)
                per = [O.split_parts(data[c], k)[0] for c in range(n)]
                parts = [np.stack([per[c][j] for c in range(n)]) for j in range(k)] + [np.ascontiguousarray(par[:, r]) for r in range(m)]
                pcrc = []
                for j in range(k):
                    cj = np.full((n, pb), 0xD7978EEB, dtype=np.uint32)
                    sub = crc[:, j:nblocks:k]
                    cj[:, : sub.shape[1]] = sub
                    pcrc.append(cj)
                pcrc += [np.ascontiguousarray(crc[:, nblocks + r * pb: nblocks + (r + 1) * pb]) for r in range(m)]
                avail = [None if i == 0 else parts[i] for i in range(k + m)]
                out, _ = eng.recover_chunks(goal, nblocks, avail, part_crc=[None if a is None else pcrc[i] for i, a in enumerate(avail)])
                assert (out[0] == parts[0]).all(), (t, rep)
        except Exception as exc:  # noqa: BLE001
            errors.append((t, repr(exc)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(10)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    st = eng.stats()
    assert st["chunks_encoded"] == 10 * 6 * 3
    eng.close()
