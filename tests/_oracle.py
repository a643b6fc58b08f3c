"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when present, the real
reference build (oracle/_ref/liblzref.so).  TEST INFRASTRUCTURE ONLY — nothing under
lizardfs_b200/ imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
BLOCK = 65536

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
pp = C.POINTER(C.c_void_p)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def ptr_array(arrs):
    """list of numpy arrays / None -> C array of void* (keeps nothing alive: caller does)."""
    out = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        out[i] = a.ctypes.data if a is not None else None
    return out


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


class Lib:
    """Thin wrapper giving the same method names for the restatement (prefix lzo_) and the
    compiled reference (prefix ref_)."""

    def __init__(self, cdll, prefix):
        self.dll = cdll
        self.prefix = prefix
        self.is_ref = prefix == "ref_"

    def fn(self, name, restype=C.c_int):
        names = {
            "crc32": "lzo_crc32" if not self.is_ref else "ref_mycrc32",
            "crc32_combine": "lzo_crc32_combine" if not self.is_ref else "ref_mycrc32_combine",
            "crc32_zeroblock": "lzo_crc32_zeroblock" if not self.is_ref else "ref_mycrc32_zeroblock",
            "crc32_zeroexpanded": "lzo_crc32_zeroexpanded" if not self.is_ref else "ref_mycrc32_zeroexpanded",
            "crc32_xorblocks": "lzo_crc32_xorblocks" if not self.is_ref else "ref_mycrc32_xorblocks",
        }
        f = getattr(self.dll, names.get(name, self.prefix + name))
        f.restype = restype
        return f

    # --- scalars -------------------------------------------------------------------
    def gf_mul(self, a, b):
        return self.fn("gf_mul", C.c_uint8)(C.c_uint8(a), C.c_uint8(b))

    def gf_inv(self, a):
        return self.fn("gf_inv", C.c_uint8)(C.c_uint8(a))

    def crc32(self, crc, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        return self.fn("crc32", C.c_uint32)(C.c_uint32(crc), _ptr(data), C.c_uint32(data.size))

    def crc32_combine(self, c1, c2, len2):
        return self.fn("crc32_combine", C.c_uint32)(C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(len2))

    def crc32_zeroblock(self, crc, zeros):
        return self.fn("crc32_zeroblock", C.c_uint32)(C.c_uint32(crc), C.c_uint32(zeros))

    def crc32_xorblocks(self, crc, c1, c2, n):
        return self.fn("crc32_xorblocks", C.c_uint32)(C.c_uint32(crc), C.c_uint32(c1), C.c_uint32(c2), C.c_uint32(n))

    # --- matrices ------------------------------------------------------------------
    def gen_rs_matrix(self, rows, k):
        a = np.zeros(rows * k, dtype=np.uint8)
        self.fn("gf_gen_rs_matrix", None)(_ptr(a), rows, k)
        return a.reshape(rows, k)

    def gen_cauchy1_matrix(self, rows, k):
        a = np.zeros(rows * k, dtype=np.uint8)
        self.fn("gf_gen_cauchy1_matrix", None)(_ptr(a), rows, k)
        return a.reshape(rows, k)

    def invert_matrix(self, mat):
        n = mat.shape[0]
        a = np.ascontiguousarray(mat, dtype=np.uint8).copy()
        out = np.zeros((n, n), dtype=np.uint8)
        rc = self.fn("gf_invert_matrix")(_ptr(a), _ptr(out), n)
        return rc, out

    def init_tables(self, coeffs):
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint8)
        rows, k = coeffs.shape
        t = np.zeros(rows * k * 32, dtype=np.uint8)
        self.fn("ec_init_tables", None)(k, rows, _ptr(coeffs), _ptr(t))
        return t

    def ec_encode_data(self, tables, srcs, n_dst):
        ln = srcs[0].size
        dsts = [np.zeros(ln, dtype=np.uint8) for _ in range(n_dst)]
        self.fn("ec_encode_data", None)(ln, len(srcs), n_dst, _ptr(tables), ptr_array(srcs), ptr_array(dsts))
        return dsts

    # --- ReedSolomon ---------------------------------------------------------------
    def rs_encode(self, k, m, data, size):
        par = [np.zeros(size, dtype=np.uint8) for _ in range(m)]
        rc = self.fn("rs_encode")(k, m, ptr_array(data), ptr_array(par), C.c_size_t(size))
        assert rc == 0
        return par

    def rs_recover(self, k, m, parts, erased, want, size):
        out = [np.zeros(size, dtype=np.uint8) if (erased[i] and want[i]) else None for i in range(k + m)]
        er = np.asarray(erased, dtype=np.uint8)
        rc = self.fn("rs_recover")(k, m, ptr_array(parts), _ptr(er), ptr_array(out), C.c_size_t(size))
        assert rc == 0
        return out

    def block_xor(self, dst, src):
        self.fn("block_xor", None)(_ptr(dst), _ptr(src), C.c_size_t(src.size))

    # --- chunk level ---------------------------------------------------------------
    def encode_chunk(self, kind, k, m, chunk):
        """chunk: uint8 array (any length for the restatement; multiple of 64 KiB for ref)."""
        nb = (chunk.size + BLOCK - 1) // BLOCK
        pb = (nb + k - 1) // k
        parity = np.zeros(m * pb * BLOCK, dtype=np.uint8)
        crc = np.zeros(nb + m * pb, dtype=np.uint32)
        if self.is_ref and chunk.size % BLOCK:
            padded = np.zeros(nb * BLOCK, dtype=np.uint8)
            padded[: chunk.size] = chunk
            chunk = padded
        rc = self.fn("encode_chunk")(kind, k, m, _ptr(chunk), C.c_size_t(chunk.size), _ptr(parity), _ptr(crc))
        assert rc == 0, rc
        return parity.reshape(m, pb * BLOCK), crc

    def recover_chunk(self, kind, k, m, parts, part_crc, want, pb):
        n = k + m
        out = [np.zeros(pb * BLOCK, dtype=np.uint8) if (want[i] and parts[i] is None) else None for i in range(n)]
        bad = (C.c_int * 2)(-1, -1)
        w = np.asarray(want, dtype=np.uint8)
        rc = self.fn("recover_chunk")(kind, k, m, ptr_array(parts),
                                      ptr_array(part_crc) if part_crc is not None else None,
                                      _ptr(w), ptr_array(out), pb, bad)
        return rc, out, (bad[0], bad[1])


def write_data_prefix(lib, chunk_id, write_id, block, offset, size, crc):
    """38-byte LIZ_CLTOCS_WRITE_DATA prefix from the oracle restatement or the compiled reference."""
    out = np.zeros(38, dtype=np.uint8)
    f = lib.fn("write_data_prefix", None if not lib.is_ref else C.c_int)
    f(_ptr(out), C.c_uint64(chunk_id), C.c_uint32(write_id), C.c_uint16(block), C.c_uint32(offset), C.c_uint32(size), C.c_uint32(crc))
    return out


def load_oracle():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    return Lib(C.CDLL(path), "lzo_")


def load_ref():
    path = os.path.join(ORACLE_DIR, "_ref", "liblzref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/src/common"):
            build_oracle()
        if not os.path.exists(path):
            return None
    return Lib(C.CDLL(path), "ref_")


def fill_chunk(oracle, nbytes, seed, chunk_index=0):
    a = np.zeros(nbytes, dtype=np.uint8)
    f = oracle.dll.lzo_fill_chunk
    f.restype = None
    f(_ptr(a), C.c_size_t(nbytes), C.c_uint64(seed), C.c_uint64(chunk_index))
    return a


def split_parts(chunk, k):
    """chunk order -> k part-major zero-padded data parts (chunk_writer.cc:505)."""
    nb = (chunk.size + BLOCK - 1) // BLOCK
    pb = (nb + k - 1) // k
    padded = np.zeros(nb * BLOCK, dtype=np.uint8)
    padded[: chunk.size] = chunk
    blocks = padded.reshape(nb, BLOCK)
    parts = [np.zeros((pb, BLOCK), dtype=np.uint8) for _ in range(k)]
    for b in range(nb):
        parts[b % k][b // k] = blocks[b]
    return [p.reshape(-1) for p in parts], pb
