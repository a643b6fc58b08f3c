/*
 * ref_shim.cc — C API over the UNMODIFIED reference implementation (test infrastructure).
 *
 * Compiled by oracle/Makefile together with the reference's own translation units, which
 * stay where they lie under /root/reference (nothing is copied into this repository):
 *   src/common/{galois_field_isal,galois_field_encode,crc,block_xor}.cc
 *   external/crcutil-1.0/code/*.cc
 * The result, oracle/_ref/liblzref.so, is the "real reference" arm used to pin
 * oracle/lzoracle.c, to generate tests/golden/, and as the CPU baseline (kind "reference").
 */
#include "common/platform.h"

#include <stdexcept>
#include <functional>
#include <numeric>
#include <cstring>
#include <vector>

#include "common/block_xor.h"
#include "common/crc.h"
#include "common/galois_coeff.h"
#include "common/reed_solomon.h"
#include "protocol/cltocs.h"

// not declared in galois_field.h but defined in galois_field_isal.cc:37,46
uint8_t gf_mul(uint8_t a, uint8_t b);
uint8_t gf_inv(uint8_t a);

typedef ReedSolomon<32, 32> RS;

static const size_t kBlock = MFSBLOCKSIZE;

extern "C" {

uint8_t ref_gf_mul(uint8_t a, uint8_t b) { return gf_mul(a, b); }
uint8_t ref_gf_inv(uint8_t a) { return gf_inv(a); }
void ref_gf_tables(uint8_t *log256, uint8_t *exp256) {
	for (int i = 0; i < 256; ++i) { log256[i] = gf_log_table[i]; exp256[i] = gf_exp_table[i]; }
}
void ref_gf_gen_rs_matrix(uint8_t *a, int m, int k) { gf_gen_rs_matrix(a, m, k); }
void ref_gf_gen_cauchy1_matrix(uint8_t *a, int m, int k) { gf_gen_cauchy1_matrix(a, m, k); }
int ref_gf_invert_matrix(uint8_t *in, uint8_t *out, int n) { return gf_invert_matrix(in, out, n); }
void ref_ec_init_tables(int k, int rows, uint8_t *a, uint8_t *tbls) { ec_init_tables(k, rows, a, tbls); }
void ref_ec_encode_data(int len, int srcs, int dests, uint8_t *v, uint8_t **src, uint8_t **dest) {
	ec_encode_data(len, srcs, dests, v, src, dest);
}

int ref_rs_encode(int k, int m, const uint8_t *const *data, uint8_t *const *parity, size_t size) {
	RS rs(k, m);
	RS::ConstFragmentMap in{{0}};
	RS::FragmentMap out{{0}};
	for (int i = 0; i < k; ++i) in[i] = data[i];
	for (int i = 0; i < m; ++i) out[i] = parity[i];
	rs.encode(in, out, size);
	return 0;
}

int ref_rs_recover(int k, int m, const uint8_t *const *in_parts, const uint8_t *erased_flags,
                   uint8_t *const *out_parts, size_t size) {
	RS rs(k, m);
	RS::ConstFragmentMap in{{0}};
	RS::FragmentMap out{{0}};
	RS::ErasedMap erased;
	for (int i = 0; i < k + m; ++i) {
		in[i] = in_parts[i];
		out[i] = out_parts[i];
		if (erased_flags[i]) erased.set(i);
	}
	rs.recover(in, erased, out, size);
	return 0;
}

void ref_block_xor(uint8_t *dest, const uint8_t *src, size_t size) { blockXor(dest, src, size); }

uint32_t ref_mycrc32(uint32_t crc, const uint8_t *block, uint32_t len) { return mycrc32(crc, block, len); }
uint32_t ref_mycrc32_combine(uint32_t c1, uint32_t c2, uint32_t len2) { return mycrc32_combine(c1, c2, len2); }
uint32_t ref_mycrc32_zeroblock(uint32_t crc, uint32_t zeros) { return mycrc32_zeroblock(crc, zeros); }
uint32_t ref_mycrc32_zeroexpanded(uint32_t crc, const uint8_t *b, uint32_t len, uint32_t zeros) {
	return mycrc32_zeroexpanded(crc, b, len, zeros);
}
uint32_t ref_mycrc32_xorblocks(uint32_t crc, uint32_t c1, uint32_t c2, uint32_t len) {
	return mycrc32_xorblocks(crc, c1, c2, len);
}
void ref_recompute_crc_if_block_empty(uint8_t *block, uint32_t *crc) {
	uint32_t c = *crc;
	recompute_crc_if_block_empty(block, c);
	*crc = c;
}

/* cltocs::writeData::serializePrefix (src/protocol/cltocs.h:118-123): the bytes WriteExecutor::addDataPacket puts
 * in front of every block it sends (src/common/write_executor.cc:99-103).  Returns the prefix length (38). */
int ref_write_data_prefix(uint8_t *out, uint64_t chunk_id, uint32_t write_id, uint16_t block, uint32_t offset,
                          uint32_t size, uint32_t crc) {
	std::vector<uint8_t> buf;
	cltocs::writeData::serializePrefix(buf, chunk_id, write_id, block, offset, size, crc);
	std::memcpy(out, buf.data(), buf.size());
	return static_cast<int>(buf.size());
}

/* The reference CALL PATTERN of the write path for one chunk held in chunk order:
 * ChunkWriter::computeParityBlock (src/mount/chunk_writer.cc:365-401) once per stripe per parity
 * part — a fresh ReedSolomon object, erased = all parity parts, one output — and
 * mycrc32(0, block, 65536) on every data and parity block (src/common/write_executor.cc:97).
 * chunk_len must be a multiple of 65536 here (callers pad). kind 0 = xor, 1 = ec. */
int ref_encode_chunk(int kind, int k, int m, const uint8_t *chunk, size_t chunk_len,
                     uint8_t *parity, uint32_t *crc) {
	if (chunk_len % kBlock) return -1;
	uint32_t nb = chunk_len / kBlock;
	uint32_t pb = (nb + k - 1) / k;
	for (uint32_t s = 0; s < pb; ++s) {
		for (int r = 0; r < m; ++r) {
			uint8_t *dst = parity + ((size_t)r * pb + s) * kBlock;
			if (kind == 0) {
				std::memcpy(dst, chunk + (size_t)s * k * kBlock, kBlock);
				for (int i = 1; i < k; ++i) {
					uint32_t b = s * k + i;
					if (b < nb) blockXor(dst, chunk + (size_t)b * kBlock, kBlock);
				}
			} else {
				RS rs(k, m);
				RS::ErasedMap erased;
				RS::ConstFragmentMap in{{0}};
				RS::FragmentMap out{{0}};
				for (int i = 0; i < m; ++i) erased.set(k + i);
				for (int i = 0; i < k; ++i) {
					uint32_t b = s * k + i;
					in[i] = b < nb ? chunk + (size_t)b * kBlock : nullptr;
				}
				out[k + r] = dst;
				rs.recover(in, erased, out, kBlock);
			}
		}
	}
	for (uint32_t b = 0; b < nb; ++b) crc[b] = mycrc32(0, chunk + (size_t)b * kBlock, kBlock);
	for (uint32_t i = 0; i < (uint32_t)m * pb; ++i) crc[nb + i] = mycrc32(0, parity + (size_t)i * kBlock, kBlock);
	return 0;
}

/* "Best-case reference kernel": one rs.encode over whole part-major parts
 * (reed_solomon.h:134-155) + mycrc32 per block.  `parts` = k part-major data parts of pb blocks
 * (zero padded) prepared by the caller, so that only the arithmetic is timed. */
int ref_encode_parts(int kind, int k, int m, const uint8_t *const *parts, uint32_t pb,
                     uint8_t *parity, uint32_t *crc_parts /* (k+m)*pb */) {
	size_t bytes = (size_t)pb * kBlock;
	if (kind == 0) {
		std::memcpy(parity, parts[0], bytes);
		for (int i = 1; i < k; ++i) blockXor(parity, parts[i], bytes);
	} else {
		RS rs(k, m);
		RS::ConstFragmentMap in{{0}};
		RS::FragmentMap out{{0}};
		for (int i = 0; i < k; ++i) in[i] = parts[i];
		for (int r = 0; r < m; ++r) out[r] = parity + (size_t)r * bytes;
		rs.encode(in, out, bytes);
	}
	for (int i = 0; i < k; ++i)
		for (uint32_t b = 0; b < pb; ++b) crc_parts[(size_t)i * pb + b] = mycrc32(0, parts[i] + (size_t)b * kBlock, kBlock);
	for (uint32_t i = 0; i < (uint32_t)m * pb; ++i) crc_parts[(size_t)k * pb + i] = mycrc32(0, parity + (size_t)i * kBlock, kBlock);
	return 0;
}

/* Degraded read of one chunk, ECReadPlan::recoverParts semantics (src/common/ec_read_plan.h:113-146):
 * the first k available parts are the inputs, the others are marked erased, every wanted
 * unavailable part is produced by ONE rs.recover over whole parts; inputs are CRC-verified per
 * block first (src/common/read_operation_executor.cc:257-269).  xor: xor_read_plan.h:77-126. */
int ref_recover_chunk(int kind, int k, int m, const uint8_t *const *parts,
                      const uint32_t *const *part_crc, const uint8_t *want,
                      uint8_t *const *out, int pb, int *bad) {
	size_t bytes = (size_t)pb * kBlock;
	int avail = 0;
	for (int i = 0; i < k + m; ++i) avail += parts[i] != nullptr;
	if (avail < k) return -2;
	if (part_crc) {
		for (int i = 0; i < k + m; ++i) {
			if (!parts[i] || !part_crc[i]) continue;
			for (int b = 0; b < pb; ++b)
				if (mycrc32(0, parts[i] + (size_t)b * kBlock, kBlock) != part_crc[i][b]) {
					if (bad) { bad[0] = i; bad[1] = b; }
					return -3;
				}
		}
	}
	if (kind == 0) {
		for (int w = 0; w < k + 1; ++w) {
			if (!want[w] || parts[w] || !out[w]) continue;
			bool first = true;
			for (int i = 0; i < k + 1; ++i) {
				if (i == w || !parts[i]) continue;
				if (first) { std::memcpy(out[w], parts[i], bytes); first = false; }
				else blockXor(out[w], parts[i], bytes);
			}
		}
		return 0;
	}
	RS rs(k, m);
	RS::ConstFragmentMap in{{0}};
	RS::FragmentMap res{{0}};
	RS::ErasedMap erased;
	int used = 0, n_out = 0;
	for (int i = 0; i < k + m; ++i) {
		if (!parts[i] || used >= k) erased.set(i);
		else { in[i] = parts[i]; ++used; }
	}
	for (int i = 0; i < k + m; ++i)
		if (want[i] && !parts[i] && out[i]) { res[i] = out[i]; ++n_out; }
	if (n_out) rs.recover(in, erased, res, bytes);
	return 0;
}


// hdd_write (src/chunkserver/hddspacemgr.cc:1898-2008) cannot be compiled without the chunkserver (Chunk objects, file I/O), so its
// body is transcribed here with the SAME sequence of calls into the reference's own crc.cc — mycrc32, mycrc32_combine,
// mycrc32_zeroblock, recompute_crc_if_block_empty (the interleaved-format reader applies it, :1779) — and the block file replaced by a
// 64 KiB memory buffer.  Every CRC value is therefore computed by reference code; only the I/O is stood in for.  Return codes as
// lzo_hdd_write_block: 0 ok, -1 wrong size / offset (:1911-1916), -3 packet CRC (:1917-1919), -4 stored block fails (:1962-1971).
int ref_hdd_write_block(uint8_t *block /* nullptr: blocknum >= chunk->blocks */, uint32_t *stored_crc, uint32_t offset, uint32_t size,
                        uint32_t crc, const uint8_t *buffer, uint8_t *new_block) {
	uint32_t precrc, postcrc, combinedcrc, chcrc;
	if (size > MFSBLOCKSIZE) return -1;
	if ((offset >= MFSBLOCKSIZE) || (offset + size > MFSBLOCKSIZE)) return -1;
	if (crc != mycrc32(0, buffer, size)) return -3;
	if (offset == 0 && size == MFSBLOCKSIZE) {
		std::memcpy(block ? block : new_block, buffer, MFSBLOCKSIZE);
		*stored_crc = crc;
		return 0;
	}
	uint8_t *data_in_buffer;
	if (block) {
		uint32_t have = *stored_crc;
		recompute_crc_if_block_empty(block, have);
		data_in_buffer = block;
		precrc = mycrc32(0, data_in_buffer, offset);
		chcrc = mycrc32(0, data_in_buffer + offset, size);
		postcrc = mycrc32(0, data_in_buffer + offset + size, MFSBLOCKSIZE - (offset + size));
		if (offset == 0) {
			combinedcrc = mycrc32_combine(chcrc, postcrc, MFSBLOCKSIZE - (offset + size));
		} else {
			combinedcrc = mycrc32_combine(precrc, chcrc, size);
			if ((offset + size) < MFSBLOCKSIZE) combinedcrc = mycrc32_combine(combinedcrc, postcrc, MFSBLOCKSIZE - (offset + size));
		}
		if (have != combinedcrc) return -4;
	} else {
		data_in_buffer = new_block;
		std::memset(data_in_buffer, 0, MFSBLOCKSIZE);   // ftruncate: the block comes into being as zeros
		precrc = mycrc32_zeroblock(0, offset);
		postcrc = mycrc32_zeroblock(0, MFSBLOCKSIZE - (offset + size));
	}
	if (offset == 0) {
		combinedcrc = mycrc32_combine(crc, postcrc, MFSBLOCKSIZE - (offset + size));
	} else {
		combinedcrc = mycrc32_combine(precrc, crc, size);
		if ((offset + size) < MFSBLOCKSIZE) combinedcrc = mycrc32_combine(combinedcrc, postcrc, MFSBLOCKSIZE - (offset + size));
	}
	std::memcpy(data_in_buffer + offset, buffer, size);
	*stored_crc = combinedcrc;
	return 0;
}

}  // extern "C"
