/*
 * lzoracle.h — CPU ORACLE for the LizardFS erasure-coding + CRC hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * library (lizardfs_b200/liblzgpu.so) never links, loads or calls anything in oracle/.
 *
 * It is a from-scratch plain-C restatement of the reference algorithms (each function
 * cites the reference file:line it follows, paths relative to /root/reference).
 * Parity status: PINNED — validated (tests/test_oracle.py) against
 *   (1) the reference's own known answers (src/common/crc_unittest.cc:27-63),
 *   (2) the known answers printed by the compiled reference, and
 *   (3) the real reference compiled from its sources into oracle/_ref/liblzref.so
 *       (oracle/Makefile, oracle/ref_shim.cc) on random inputs, whenever that library
 *       is present; golden vectors produced by it are committed under tests/golden/.
 */
#ifndef LZORACLE_H
#define LZORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZO_BLOCK_SIZE 65536u      /* MFSBLOCKSIZE, src/protocol/MFSCommunication.h:65 */
#define LZO_BLOCKS_IN_CHUNK 1024u  /* MFSBLOCKSINCHUNK */
#define LZO_MAX_PARTS 64           /* ReedSolomon<32,32>::kMaxPartCount, reed_solomon.h:46 */

/* ---- GF(2^8), polynomial 0x11d (galois_coeff.h:30-71, galois_field_isal.cc:37-51) ---- */
uint8_t lzo_gf_mul(uint8_t a, uint8_t b);
uint8_t lzo_gf_inv(uint8_t a);
const uint8_t *lzo_gf_log_table(void); /* 256 entries, log[1] == 255 as in the reference */
const uint8_t *lzo_gf_exp_table(void); /* 256 entries, exp[0] == 1, exp[255] == 1 */

/* ---- ISA-L shaped primitives (galois_field.h:35-88) ---- */
void lzo_gf_gen_rs_matrix(uint8_t *a, int m, int k);      /* galois_field_isal.cc:53-69 */
void lzo_gf_gen_cauchy1_matrix(uint8_t *a, int m, int k); /* galois_field_isal.cc:71-85 */
int lzo_gf_invert_matrix(uint8_t *in, uint8_t *out, int n); /* galois_field_isal.cc:87-139 */
void lzo_gf_vect_mul_init(uint8_t c, uint8_t *tbl32);     /* galois_field_isal.cc:143-244 */
void lzo_ec_init_tables(int k, int rows, const uint8_t *a, uint8_t *g_tbls); /* :246-255 */
void lzo_ec_encode_data(int len, int srcs, int dests, const uint8_t *v,
                        const uint8_t *const *src, uint8_t *const *dest); /* galois_field_encode.cc:28-47 */

/* ---- ReedSolomon<32,32> semantics (reed_solomon.h:87-155, 163-281) ----
 * in[i] / out[i] are indexed by part (data 0..k-1, parity k..k+m-1); NULL input on an
 * available part means "all zeros"; NULL output on an erased part means "skip".
 * erased[i] != 0 marks part i erased; exactly m parts must be erased (reed_solomon.h:95).
 * Returns 0, or -1 on argument errors (the reference asserts). */
int lzo_rs_generator(int k, int m, uint8_t *matrix /* (k+m)*k */); /* reed_solomon.h:163-178 */
int lzo_rs_encode(int k, int m, const uint8_t *const *data, uint8_t *const *parity, size_t size);
int lzo_rs_recover(int k, int m, const uint8_t *const *in, const uint8_t *erased,
                   uint8_t *const *out, size_t size);
/* The coefficient matrix rs_recover would use: rows = wanted outputs (ascending part id),
 * cols = the k available parts (ascending part id, zero inputs NOT removed).
 * Returns the number of rows, or -1. */
int lzo_rs_recovery_matrix(int k, int m, const uint8_t *erased, const uint8_t *wanted,
                           uint8_t *matrix /* up to m*k */);

/* ---- blockXor (block_xor.cc:47-88) ---- */
void lzo_block_xor(uint8_t *dest, const uint8_t *src, size_t size);

/* ---- CRC-32 (crc.cc:52-60, crc.h:25-29; crcutil generic_crc.h, gf_util.h:92-105) ---- */
uint32_t lzo_crc32(uint32_t crc, const uint8_t *block, uint32_t len);
uint32_t lzo_crc32_combine(uint32_t crc1, uint32_t crc2, uint32_t len2);
uint32_t lzo_crc32_zeroblock(uint32_t crc, uint32_t zeros);   /* crc.h:27 */
uint32_t lzo_crc32_zeroexpanded(uint32_t crc, const uint8_t *block, uint32_t len, uint32_t zeros);
uint32_t lzo_crc32_xorblocks(uint32_t crc, uint32_t crcblock1, uint32_t crcblock2, uint32_t len);
void lzo_recompute_crc_if_block_empty(const uint8_t *block, uint32_t *crc); /* crc.cc:235-243 */

/* ---- slice geometry (slice_traits.h:311-349) ----
 * kind: 0 = xorN (k = N data parts, 1 parity), 1 = ec(k,m).
 * "data index" j is 0-based for both kinds (xor part number = j+1, slice_traits.h:283-288). */
int lzo_part_blocks(int k, int data_index_or_minus1_for_parity, uint32_t blocks_in_chunk);
int lzo_part_length(int k, int data_index_or_minus1_for_parity, int chunk_length);

/* ---- chunk-level restatement of the reference call pattern ----
 * chunk: chunk_len bytes in chunk order (block b -> data part b%k, index b/k,
 *        chunk_writer.cc:501-509).  A trailing partial block is zero-extended to a
 *        full block (as the chunkserver stores it, hddspacemgr.cc:1983-1999).
 * parity: m parts, each pb = ceil(nb/k) blocks, part r at parity + r*pb*65536.
 * crc:    nb data-block CRCs in chunk order, then for r<m the pb CRCs of parity part r.
 *         CRCs are over full 64 KiB (zero-extended) blocks.
 * Follows ChunkWriter::computeParityBlock (chunk_writer.cc:365-401): one call per stripe
 * per parity part; absent blocks of the last stripe are NULL (= zero) inputs.
 * kind 0 (xor): m must be 1. */
int lzo_encode_chunk(int kind, int k, int m, const uint8_t *chunk, size_t chunk_len,
                     uint8_t *parity, uint32_t *crc);

/* Whole-part form ("best-case reference kernel"): de-interleaves into
 * zero-padded part-major buffers and calls rs.encode once over whole parts; same outputs. */
int lzo_encode_chunk_whole(int kind, int k, int m, const uint8_t *chunk, size_t chunk_len,
                           uint8_t *parity, uint32_t *crc);

/* Degraded read: parts[i] (i < k+m; xor: index 0 = parity, 1..k data as in the reference is
 * NOT used here — for both kinds index = data 0..k-1 then parity) point to part-major
 * buffers of pb blocks each, or NULL when unavailable.  part_crc[i] (may be NULL) holds the
 * pb stored CRCs of part i, verified with mycrc32 (read_operation_executor.cc:257-269).
 * Recovers every part in `want` (flags, k+m) that is unavailable into out[i]
 * (ECReadPlan::recoverParts, ec_read_plan.h:113-146: first k available parts are the
 * inputs, surplus ones are marked erased; XorReadPlan::postProcessRead, xor_read_plan.h:77-126).
 * Returns 0, -2 if fewer than k parts are available, -3 on CRC mismatch
 * (first bad (part, block) stored in bad[0], bad[1] when bad != NULL). */
int lzo_recover_chunk(int kind, int k, int m, const uint8_t *const *parts,
                      const uint32_t *const *part_crc, const uint8_t *want,
                      uint8_t *const *out, int pb, int *bad);

/* part-major -> chunk order gather (chunk_read_planner.h:36-70) */
void lzo_parts_to_chunk(int k, const uint8_t *const *data_parts, uint32_t nb, uint8_t *chunk);

/* Replication / slice-type conversion (SliceRecoveryPlanner, src/chunkserver/slice_recovery_planner.h:87-204 + the CRC loop
 * of chunk_replicator.cc:186-192).  kinds: 0 xor, 1 ec, 2 standard (k = 1, m = 0).  Parts indexed data 0..k-1, parity k...
 * out[i] holds pbd = ceil(nb/dk) blocks (nb for a standard destination), out_crc[i] as many CRCs.  Returns as lzo_recover_chunk. */
int lzo_convert_chunk(int skind, int sk, int sm, const uint8_t *const *parts, const uint32_t *const *part_crc,
                      int dkind, int dk, int dm, const uint8_t *want, uint8_t *const *out, uint32_t *const *out_crc,
                      uint32_t nb, int *bad);

/* hdd_int_test (hddspacemgr.cc:2148-2210) over in-memory chunk files of both on-disk formats (chunk.cc:126-209). */
int lzo_scrub_interleaved(const uint8_t *records, size_t n_blocks, int64_t *first_bad);
size_t lzo_moosefs_header_size(int data_parts);
int lzo_scrub_moosefs(const uint8_t *image, int data_parts, size_t n_blocks, int64_t *first_bad);

/* hdd_write of one block (hddspacemgr.cc:1898-2008): packet CRC check, stored-block check via mycrc32_combine, new CRC. */
int lzo_hdd_write_block(uint8_t *block, uint32_t *stored_crc, uint32_t offset, uint32_t size, uint32_t crc,
                        const uint8_t *buffer, uint8_t *new_block);

/* LIZ_CLTOCS_WRITE_DATA packet prefix (src/protocol/cltocs.h:116-137, packet.h:130-136, MFSCommunication.h:630):
 * header type:u32 = 1212, length:u32 = 30 + size; then version:u32 = 0, chunkId:u64, writeId:u32, block:u16,
 * offset:u32, size:u32, crc:u32 — all big-endian, 38 bytes; `size` data bytes follow on the wire. */
#define LZO_WRITE_PREFIX_SIZE 38
void lzo_write_data_prefix(uint8_t *out38, uint64_t chunk_id, uint32_t write_id, uint16_t block, uint32_t offset,
                           uint32_t size, uint32_t crc);

/* deterministic synthetic data shared by oracle, tests and the GPU generator:
 * splitmix64 counter stream, 8-byte word w of chunk c = mix(seed + ((c<<23) + w + 1) * 0x9E3779B97F4A7C15), little-endian. */
void lzo_fill_chunk(uint8_t *dst, size_t len, uint64_t seed, uint64_t chunk_index);

#ifdef __cplusplus
}
#endif
#endif
