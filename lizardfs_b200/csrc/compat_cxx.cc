// compat_cxx.cc — C++-linkage symbols with the reference's exact names and signatures, so that
// liblzgpu.so can stand in for src/common/{crc,block_xor,galois_field_isal,galois_field_encode}.cc
// at link time (the ISA-L-named functions already have C linkage in engine.cu/host_math.cc, which is
// what <isa-l/erasure_code.h> declares; the local galois_field.h declares them with C++ linkage,
// hence the C++-linkage twins of those five in compat_cxx_gf.cc, which forward to the lzgpu_isal_* aliases defined at the
// end of this file).
#include <cstddef>
#include <cstdint>

#include "lzgpu.h"

// src/common/crc.h:25-36
uint32_t mycrc32(uint32_t crc, const uint8_t *block, uint32_t leng) { return lzgpu_mycrc32(crc, block, leng); }
uint32_t mycrc32_combine(uint32_t crc1, uint32_t crc2, uint32_t leng2) { return lzgpu_mycrc32_combine(crc1, crc2, leng2); }
void mycrc32_init(void) { lzgpu_mycrc32_init(); }
void recompute_crc_if_block_empty(uint8_t *block, uint32_t &crc) { lzgpu_recompute_crc_if_block_empty(block, &crc); }

// src/common/block_xor.h:33
void blockXor(uint8_t *dest, const uint8_t *source, size_t size) { lzgpu_block_xor(dest, source, size); }

// lzgpu_-prefixed aliases of the ISA-L names: the way to reach them from a translation unit that already has C++-linkage
// declarations of the plain names (compat_cxx_gf.cc), or from a host that must not pollute its namespace
extern "C" {
void lzgpu_isal_gf_gen_rs_matrix(unsigned char *a, int m, int k) { gf_gen_rs_matrix(a, m, k); }
void lzgpu_isal_gf_gen_cauchy1_matrix(unsigned char *a, int m, int k) { gf_gen_cauchy1_matrix(a, m, k); }
int lzgpu_isal_gf_invert_matrix(unsigned char *in, unsigned char *out, const int n) { return gf_invert_matrix(in, out, n); }
void lzgpu_isal_ec_init_tables(int k, int rows, unsigned char *a, unsigned char *gftbls) { ec_init_tables(k, rows, a, gftbls); }
void lzgpu_isal_ec_encode_data(int len, int srcs, int dests, unsigned char *v, unsigned char **src, unsigned char **dest) {
	ec_encode_data(len, srcs, dests, v, src, dest);
}
}
