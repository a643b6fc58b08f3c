// compat_cxx.cc — C++-linkage symbols with the reference's exact names and signatures, so that
// liblzgpu.so can stand in for src/common/{crc,block_xor,galois_field_isal,galois_field_encode}.cc
// at link time (the ISA-L-named functions already have C linkage in engine.cu/host_math.cc, which is
// what <isa-l/erasure_code.h> declares; the local galois_field.h declares them with C++ linkage,
// hence the forwarding overloads in namespace-less C++ below under LZGPU_CXX_GF_NAMES).
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
