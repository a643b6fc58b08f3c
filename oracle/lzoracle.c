/*
 * lzoracle.c — CPU ORACLE (test infrastructure only; see lzoracle.h for the rules).
 *
 * Plain-C restatement of the reference's arithmetic for the erasure-coding + CRC hot
 * path.  Written for clarity, not speed: scalar loops, byte tables.  Every function
 * cites the reference code it restates (paths relative to /root/reference).
 */
#include "lzoracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * GF(2^8) tables.  galois_coeff.h:30-32 (mul2 with 0x1d), :40-42/:56-58 (log/exp by
 * repeated doubling starting at n=1, pow=2), :68-71 (tables: log[0]=0, exp[0]=1).
 * Consequence restated here: exp[i] = 2^i for i = 0..255 (exp[255] = 1) and
 * log[x] = the i in 1..255 with 2^i = x, i.e. log[1] = 255, not 0.
 * ---------------------------------------------------------------------------------- */
static uint8_t g_log[256], g_exp[256];
static int g_gf_ready;

static uint8_t mul2(uint8_t x) { return (uint8_t)((x << 1) ^ ((x & 0x80) ? 0x1d : 0)); }

static void gf_setup(void) {
	if (g_gf_ready) return;
	uint8_t p = 1;
	g_exp[0] = 1;
	g_log[0] = 0;
	for (int i = 1; i <= 255; ++i) {
		p = mul2(p);
		g_exp[i] = p;      /* 2^i */
		g_log[p] = (uint8_t)i; /* last write for p==1 is i==255 */
	}
	g_gf_ready = 1;
}

const uint8_t *lzo_gf_log_table(void) { gf_setup(); return g_log; }
const uint8_t *lzo_gf_exp_table(void) { gf_setup(); return g_exp; }

/* galois_field_isal.cc:37-44 */
uint8_t lzo_gf_mul(uint8_t a, uint8_t b) {
	gf_setup();
	if (a == 0 || b == 0) return 0;
	int s = g_log[a] + g_log[b];
	return g_exp[s > 254 ? s - 255 : s];
}

/* galois_field_isal.cc:46-51 */
uint8_t lzo_gf_inv(uint8_t a) {
	gf_setup();
	if (a == 0) return 0;
	return g_exp[255 - g_log[a]];
}

/* galois_field_isal.cc:53-69: identity on top, then row i>=k is the geometric
 * progression 1, g, g^2, ... with g = 2^(i-k). */
void lzo_gf_gen_rs_matrix(uint8_t *a, int m, int k) {
	memset(a, 0, (size_t)k * m);
	for (int i = 0; i < k; ++i) a[k * i + i] = 1;
	uint8_t gen = 1;
	for (int i = k; i < m; ++i) {
		uint8_t p = 1;
		for (int j = 0; j < k; ++j) {
			a[k * i + j] = p;
			p = lzo_gf_mul(p, gen);
		}
		gen = lzo_gf_mul(gen, 2);
	}
}

/* galois_field_isal.cc:71-85: identity on top, then 1/(i xor j). */
void lzo_gf_gen_cauchy1_matrix(uint8_t *a, int m, int k) {
	memset(a, 0, (size_t)k * m);
	for (int i = 0; i < k; ++i) a[k * i + i] = 1;
	for (int i = k; i < m; ++i)
		for (int j = 0; j < k; ++j) a[k * i + j] = lzo_gf_inv((uint8_t)(i ^ j));
}

/* galois_field_isal.cc:87-139: Gauss-Jordan; pivot search only below the diagonal,
 * row swap, scale pivot row by 1/pivot, eliminate the column from every other row.
 * Destroys `in`.  Returns -1 when no pivot can be found. */
int lzo_gf_invert_matrix(uint8_t *in, uint8_t *out, int n) {
	memset(out, 0, (size_t)n * n);
	for (int i = 0; i < n; ++i) out[i * n + i] = 1;
	for (int i = 0; i < n; ++i) {
		if (in[i * n + i] == 0) {
			int j = i + 1;
			while (j < n && in[j * n + i] == 0) ++j;
			if (j == n) return -1;
			for (int c = 0; c < n; ++c) {
				uint8_t t = in[i * n + c]; in[i * n + c] = in[j * n + c]; in[j * n + c] = t;
				t = out[i * n + c]; out[i * n + c] = out[j * n + c]; out[j * n + c] = t;
			}
		}
		uint8_t inv = lzo_gf_inv(in[i * n + i]);
		for (int c = 0; c < n; ++c) {
			in[i * n + c] = lzo_gf_mul(in[i * n + c], inv);
			out[i * n + c] = lzo_gf_mul(out[i * n + c], inv);
		}
		for (int r = 0; r < n; ++r) {
			if (r == i) continue;
			uint8_t f = in[r * n + i];
			for (int c = 0; c < n; ++c) {
				out[r * n + c] ^= lzo_gf_mul(f, out[i * n + c]);
				in[r * n + c] ^= lzo_gf_mul(f, in[i * n + c]);
			}
		}
	}
	return 0;
}

/* galois_field_isal.cc:143-244: 32-byte table for coefficient c:
 * tbl[n] = c*n for the low nibble n, tbl[16+n] = c*(n<<4) for the high nibble. */
void lzo_gf_vect_mul_init(uint8_t c, uint8_t *tbl) {
	for (int n = 0; n < 16; ++n) {
		tbl[n] = lzo_gf_mul(c, (uint8_t)n);
		tbl[16 + n] = lzo_gf_mul(c, (uint8_t)(n << 4));
	}
}

/* galois_field_isal.cc:246-255: rows*k coefficients, row-major, 32 B each. */
void lzo_ec_init_tables(int k, int rows, const uint8_t *a, uint8_t *g_tbls) {
	for (int i = 0; i < rows * k; ++i) lzo_gf_vect_mul_init(a[i], g_tbls + 32 * i);
}

/* galois_field_encode.cc:28-47 (the portable kernel; the SIMD variants compute the same
 * bytes): dest[l][i] = XOR_j lo[l][j][src[j][i] & 15] ^ hi[l][j][src[j][i] >> 4]. */
void lzo_ec_encode_data(int len, int srcs, int dests, const uint8_t *v,
                        const uint8_t *const *src, uint8_t *const *dest) {
	for (int l = 0; l < dests; ++l) {
		const uint8_t *row = v + (size_t)l * srcs * 32;
		for (int i = 0; i < len; ++i) {
			uint8_t s = 0;
			for (int j = 0; j < srcs; ++j) {
				uint8_t a = src[j][i];
				s ^= row[32 * j + (a & 15)] ^ row[32 * j + 16 + (a >> 4)];
			}
			dest[l][i] = s;
		}
	}
}

/* ------------------------------------------------------------------------------------
 * ReedSolomon<32,32>
 * ---------------------------------------------------------------------------------- */

/* reed_solomon.h:163-178: Cauchy when m >= 5 or (m == 4 and k > 20), else Vandermonde. */
int lzo_rs_generator(int k, int m, uint8_t *matrix) {
	if (k < 1 || k > 32 || m < 1 || m > 32) return -1;
	if (m >= 5 || (m == 4 && k > 20)) lzo_gf_gen_cauchy1_matrix(matrix, k + m, k);
	else lzo_gf_gen_rs_matrix(matrix, k + m, k);
	return 0;
}

/* Plain GF matrix product C(r x c) = A(r x n) * B(n x c); reed_solomon.h:343-358 computes
 * the same thing through ec_encode_data. */
static void gf_matmul(uint8_t *C, const uint8_t *A, const uint8_t *B, int r, int n, int c) {
	for (int i = 0; i < r; ++i)
		for (int j = 0; j < c; ++j) {
			uint8_t s = 0;
			for (int t = 0; t < n; ++t) s ^= lzo_gf_mul(A[i * n + t], B[t * c + j]);
			C[i * c + j] = s;
		}
}

/* reed_solomon.h:189-217 and :229-281 without the zero-column removal: returns the
 * coefficient rows for the wanted parts over the k available parts. */
int lzo_rs_recovery_matrix(int k, int m, const uint8_t *erased, const uint8_t *wanted,
                           uint8_t *matrix) {
	uint8_t G[64 * 32], avail[32 * 32], inv[32 * 32], want_rows[32 * 32];
	if (lzo_rs_generator(k, m, G) != 0) return -1;
	int n_erased = 0, n_avail_data = 0, n_want = 0, want_parity = 0;
	for (int i = 0; i < k + m; ++i) {
		n_erased += erased[i] != 0;
		if (!erased[i] && i < k) ++n_avail_data;
		if (erased[i] && wanted[i]) { ++n_want; want_parity += i >= k; }
	}
	if (n_erased != m) return -1; /* reed_solomon.h:95 */
	if (n_want == 0) return 0;
	if (n_avail_data == k) {
		/* reed_solomon.h:113-114, :189-217: all data present -> generator rows */
		int r = 0;
		for (int i = 0; i < k + m; ++i)
			if (erased[i] && wanted[i]) memcpy(matrix + (r++) * k, G + i * k, (size_t)k);
		return n_want;
	}
	/* reed_solomon.h:242-251: rows of available parts, inverted (failure ignored there) */
	int r = 0;
	for (int i = 0; i < k + m; ++i)
		if (!erased[i]) memcpy(avail + (r++) * k, G + i * k, (size_t)k);
	(void)lzo_gf_invert_matrix(avail, inv, k);
	if (want_parity == 0) {
		/* reed_solomon.h:262-264: rows of the inverse for the wanted data parts */
		r = 0;
		for (int i = 0; i < k; ++i)
			if (erased[i] && wanted[i]) memcpy(matrix + (r++) * k, inv + i * k, (size_t)k);
	} else {
		/* reed_solomon.h:253-260: generator rows of wanted parts times the inverse */
		r = 0;
		for (int i = 0; i < k + m; ++i)
			if (erased[i] && wanted[i]) memcpy(want_rows + (r++) * k, G + i * k, (size_t)k);
		gf_matmul(matrix, want_rows, inv, n_want, k, k);
	}
	return n_want;
}

/* reed_solomon.h:87-121 */
int lzo_rs_recover(int k, int m, const uint8_t *const *in, const uint8_t *erased,
                   uint8_t *const *out, size_t size) {
	uint8_t wanted[LZO_MAX_PARTS] = {0}, matrix[32 * 32], reduced[32 * 32];
	uint8_t *tables;
	const uint8_t *srcs[LZO_MAX_PARTS];
	uint8_t *dsts[LZO_MAX_PARTS];
	int n_src = 0, n_dst = 0;
	if (k < 1 || k > 32 || m < 1 || m > 32) return -1;
	for (int i = 0; i < k + m; ++i)
		if (erased[i] && out[i]) { wanted[i] = 1; dsts[n_dst++] = out[i]; }
	int rows = lzo_rs_recovery_matrix(k, m, erased, wanted, matrix);
	if (rows < 0) return -1;
	if (rows == 0) return 0;
	/* reed_solomon.h:104-110, :202-209, :266-276: NULL inputs are zero and their
	 * columns are removed before the tables are built */
	int col = 0;
	uint8_t keep[32];
	for (int i = 0; i < k + m; ++i) {
		if (erased[i]) continue;
		keep[col] = in[i] != NULL;
		if (in[i]) srcs[n_src++] = in[i];
		++col;
	}
	if (n_src == 0) { /* the reference asserts non_zero_input.count() > 0 */
		for (int d = 0; d < n_dst; ++d) memset(dsts[d], 0, size);
		return 0;
	}
	for (int r = 0; r < rows; ++r) {
		int c2 = 0;
		for (int c = 0; c < k; ++c)
			if (keep[c]) reduced[r * n_src + c2++] = matrix[r * k + c];
	}
	tables = (uint8_t *)malloc((size_t)rows * n_src * 32);
	if (!tables) return -1;
	lzo_ec_init_tables(n_src, rows, reduced, tables);
	lzo_ec_encode_data((int)size, n_src, rows, tables, srcs, dsts);
	free(tables);
	return 0;
}

/* reed_solomon.h:134-155: encode == recover with every parity part erased and wanted. */
int lzo_rs_encode(int k, int m, const uint8_t *const *data, uint8_t *const *parity, size_t size) {
	const uint8_t *in[LZO_MAX_PARTS] = {0};
	uint8_t *out[LZO_MAX_PARTS] = {0};
	uint8_t erased[LZO_MAX_PARTS] = {0};
	if (k < 1 || k > 32 || m < 1 || m > 32) return -1;
	for (int i = 0; i < k; ++i) in[i] = data[i];
	for (int r = 0; r < m; ++r) {
		if (!parity[r]) return -1; /* reed_solomon.h:148 */
		erased[k + r] = 1;
		out[k + r] = parity[r];
	}
	return lzo_rs_recover(k, m, in, erased, out, size);
}

/* block_xor.cc:47-88: dest ^= source (the alignment handling there only affects speed) */
void lzo_block_xor(uint8_t *dest, const uint8_t *src, size_t size) {
	for (size_t i = 0; i < size; ++i) dest[i] ^= src[i];
}

/* ------------------------------------------------------------------------------------
 * CRC-32.  crc.cc:52-56: crcutil GenericCrc(CRC_POLY = 0xEDB88320, degree 32,
 * canonical = true).CrcDefault(block, len, crc) — the zlib CRC: reflected polynomial,
 * the running value is complemented on entry and exit (MFSCommunication.h:81;
 * the legacy table code crc.cc:71-151 computes the same function).
 * ---------------------------------------------------------------------------------- */
#define CRC_POLY_REFLECTED 0xEDB88320u

static uint32_t g_crc_tab[256];
static int g_crc_ready;

static void crc_setup(void) {
	if (g_crc_ready) return;
	for (uint32_t i = 0; i < 256; ++i) {
		uint32_t c = i;
		for (int b = 0; b < 8; ++b) c = (c >> 1) ^ ((c & 1) ? CRC_POLY_REFLECTED : 0);
		g_crc_tab[i] = c;
	}
	g_crc_ready = 1;
}

uint32_t lzo_crc32(uint32_t crc, const uint8_t *block, uint32_t len) {
	crc_setup();
	uint32_t c = ~crc;
	for (uint32_t i = 0; i < len; ++i) c = g_crc_tab[(c ^ block[i]) & 0xff] ^ (c >> 8);
	return ~c;
}

/* Reflected-domain polynomial helpers: bit 31 is the coefficient of x^0. */
static uint32_t crc_mulmod(uint32_t a, uint32_t b) {
	uint32_t r = 0;
	for (int i = 0; i < 32; ++i) {
		if (a & 0x80000000u) r ^= b;          /* a's x^i term times b*x^i */
		a <<= 1;
		b = (b >> 1) ^ ((b & 1) ? CRC_POLY_REFLECTED : 0); /* b *= x */
	}
	return r;
}

static uint32_t crc_xpow8n(uint32_t nbytes) { /* x^(8*nbytes) mod P */
	uint32_t result = 0x80000000u;            /* 1 */
	uint32_t base = 0x00800000u;              /* x^8 */
	while (nbytes) {
		if (nbytes & 1) result = crc_mulmod(result, base);
		base = crc_mulmod(base, base);
		nbytes >>= 1;
	}
	return result;
}

/* crc.cc:58-60 -> crcutil GfUtil::Concatenate (gf_util.h:92-105):
 * CRC(A||B) = CRC(A) * x^(8|B|) + CRC(B); the complement-on-entry/exit terms cancel. */
uint32_t lzo_crc32_combine(uint32_t crc1, uint32_t crc2, uint32_t len2) {
	return crc_mulmod(crc1, crc_xpow8n(len2)) ^ crc2;
}

/* crc.h:27-29 macros restated as functions */
uint32_t lzo_crc32_zeroblock(uint32_t crc, uint32_t zeros) {
	return lzo_crc32_combine(crc ^ 0xFFFFFFFFu, 0xFFFFFFFFu, zeros);
}
uint32_t lzo_crc32_zeroexpanded(uint32_t crc, const uint8_t *block, uint32_t len, uint32_t zeros) {
	return lzo_crc32_zeroblock(lzo_crc32(crc, block, len), zeros);
}
uint32_t lzo_crc32_xorblocks(uint32_t crc, uint32_t crcblock1, uint32_t crcblock2, uint32_t len) {
	return crcblock1 ^ crcblock2 ^ lzo_crc32_zeroblock(crc, len);
}

/* crc.cc:235-243 */
void lzo_recompute_crc_if_block_empty(const uint8_t *block, uint32_t *crc) {
	if (*crc != 0) return;
	for (uint32_t i = 0; i < LZO_BLOCK_SIZE; ++i)
		if (block[i]) return;
	*crc = lzo_crc32_zeroblock(0, LZO_BLOCK_SIZE);
}

/* ------------------------------------------------------------------------------------
 * Geometry.  slice_traits.h:311-316 and :332-349.  j < 0 selects a parity part
 * (data_part_index = 0 there).
 * ---------------------------------------------------------------------------------- */
int lzo_part_blocks(int k, int j, uint32_t blocks_in_chunk) {
	int idx = j < 0 ? 0 : j;
	return (int)((blocks_in_chunk + (uint32_t)(k - idx - 1)) / (uint32_t)k);
}

int lzo_part_length(int k, int j, int chunk_length) {
	if (k == 1) return chunk_length;
	int idx = j < 0 ? 0 : j;
	int full_stripe = chunk_length / (k * (int)LZO_BLOCK_SIZE);
	int base_len = full_stripe * (int)LZO_BLOCK_SIZE;
	int rest = chunk_length - base_len * k;
	int part_rest = rest - idx * (int)LZO_BLOCK_SIZE;
	if (part_rest < 0) part_rest = 0;
	if (part_rest > (int)LZO_BLOCK_SIZE) part_rest = (int)LZO_BLOCK_SIZE;
	return base_len + part_rest;
}

/* ------------------------------------------------------------------------------------
 * Chunk-level call patterns.
 * ---------------------------------------------------------------------------------- */

/* Copy of the chunk zero-extended to whole blocks (what the chunkserver ends up storing
 * for a short trailing block, hddspacemgr.cc:1983-1999). */
static uint8_t *padded_chunk(const uint8_t *chunk, size_t chunk_len, uint32_t *nb_out) {
	uint32_t nb = (uint32_t)((chunk_len + LZO_BLOCK_SIZE - 1) / LZO_BLOCK_SIZE);
	uint8_t *p = (uint8_t *)calloc((size_t)nb ? nb : 1, LZO_BLOCK_SIZE);
	if (p) memcpy(p, chunk, chunk_len);
	*nb_out = nb;
	return p;
}

/* ChunkWriter::computeParityBlock, chunk_writer.cc:365-401 */
static void parity_block(int kind, int k, int m, int r, uint8_t *parity_blk,
                         const uint8_t *const *stripe /* k entries, NULL = absent */, size_t size) {
	if (kind == 0) {
		/* :373-381: memcpy the first block, blockXor the other present ones */
		memcpy(parity_blk, stripe[0], size);
		for (int i = 1; i < k; ++i)
			if (stripe[i]) lzo_block_xor(parity_blk, stripe[i], size);
		return;
	}
	/* :386-400: erased = every parity part, only this parity part has an output */
	const uint8_t *in[LZO_MAX_PARTS] = {0};
	uint8_t *out[LZO_MAX_PARTS] = {0};
	uint8_t erased[LZO_MAX_PARTS] = {0};
	for (int i = 0; i < m; ++i) erased[k + i] = 1;
	for (int i = 0; i < k; ++i) in[i] = stripe[i];
	out[k + r] = parity_blk;
	lzo_rs_recover(k, m, in, erased, out, size);
}

int lzo_encode_chunk(int kind, int k, int m, const uint8_t *chunk, size_t chunk_len,
                     uint8_t *parity, uint32_t *crc) {
	if (kind == 0 && m != 1) return -1;
	if (k < 1 || k > 32 || m < 1 || m > 32) return -1;
	if (chunk_len == 0 || chunk_len > (size_t)LZO_BLOCK_SIZE * LZO_BLOCKS_IN_CHUNK) return -1;
	uint32_t nb;
	uint8_t *data = padded_chunk(chunk, chunk_len, &nb);
	if (!data) return -1;
	uint32_t pb = (uint32_t)lzo_part_blocks(k, -1, nb);
	for (uint32_t s = 0; s < pb; ++s) {
		const uint8_t *stripe[32];
		for (int i = 0; i < k; ++i) {
			uint32_t b = s * (uint32_t)k + (uint32_t)i; /* chunk_writer.cc:505: part = b % k */
			stripe[i] = b < nb ? data + (size_t)b * LZO_BLOCK_SIZE : NULL;
		}
		for (int r = 0; r < m; ++r) {
			uint8_t *dst = parity + ((size_t)r * pb + s) * LZO_BLOCK_SIZE;
			parity_block(kind, k, m, r, dst, stripe, LZO_BLOCK_SIZE);
		}
	}
	/* write_executor.cc:97: mycrc32(0, data, size) on every outgoing block */
	for (uint32_t b = 0; b < nb; ++b)
		crc[b] = lzo_crc32(0, data + (size_t)b * LZO_BLOCK_SIZE, LZO_BLOCK_SIZE);
	for (int r = 0; r < m; ++r)
		for (uint32_t s = 0; s < pb; ++s)
			crc[nb + (uint32_t)r * pb + s] =
			    lzo_crc32(0, parity + ((size_t)r * pb + s) * LZO_BLOCK_SIZE, LZO_BLOCK_SIZE);
	free(data);
	return 0;
}

int lzo_encode_chunk_whole(int kind, int k, int m, const uint8_t *chunk, size_t chunk_len,
                           uint8_t *parity, uint32_t *crc) {
	if (kind == 0 && m != 1) return -1;
	if (k < 1 || k > 32 || m < 1 || m > 32) return -1;
	if (chunk_len == 0 || chunk_len > (size_t)LZO_BLOCK_SIZE * LZO_BLOCKS_IN_CHUNK) return -1;
	uint32_t nb;
	uint8_t *data = padded_chunk(chunk, chunk_len, &nb);
	if (!data) return -1;
	uint32_t pb = (uint32_t)lzo_part_blocks(k, -1, nb);
	size_t part_bytes = (size_t)pb * LZO_BLOCK_SIZE;
	uint8_t *parts = (uint8_t *)calloc((size_t)k, part_bytes);
	if (!parts) { free(data); return -1; }
	const uint8_t *in[32];
	uint8_t *out[32];
	for (uint32_t b = 0; b < nb; ++b)
		memcpy(parts + (size_t)(b % (uint32_t)k) * part_bytes + (size_t)(b / (uint32_t)k) * LZO_BLOCK_SIZE,
		       data + (size_t)b * LZO_BLOCK_SIZE, LZO_BLOCK_SIZE);
	for (int i = 0; i < k; ++i) in[i] = parts + (size_t)i * part_bytes;
	for (int r = 0; r < m; ++r) out[r] = parity + (size_t)r * part_bytes;
	if (kind == 0) {
		memcpy(out[0], in[0], part_bytes);
		for (int i = 1; i < k; ++i) lzo_block_xor(out[0], in[i], part_bytes);
	} else {
		lzo_rs_encode(k, m, in, out, part_bytes);
	}
	for (uint32_t b = 0; b < nb; ++b)
		crc[b] = lzo_crc32(0, data + (size_t)b * LZO_BLOCK_SIZE, LZO_BLOCK_SIZE);
	for (uint32_t i = 0; i < (uint32_t)m * pb; ++i)
		crc[nb + i] = lzo_crc32(0, parity + (size_t)i * LZO_BLOCK_SIZE, LZO_BLOCK_SIZE);
	free(parts);
	free(data);
	return 0;
}

int lzo_recover_chunk(int kind, int k, int m, const uint8_t *const *parts,
                      const uint32_t *const *part_crc, const uint8_t *want,
                      uint8_t *const *out, int pb, int *bad) {
	if (kind == 0 && m != 1) return -1;
	if (k < 1 || k > 32 || m < 1 || m > 32 || pb < 1) return -1;
	size_t part_bytes = (size_t)pb * LZO_BLOCK_SIZE;
	int n_avail = 0;
	for (int i = 0; i < k + m; ++i) n_avail += parts[i] != NULL;
	if (n_avail < k) return -2;
	/* read_operation_executor.cc:257-269: every received block is CRC-checked */
	if (part_crc) {
		for (int i = 0; i < k + m; ++i) {
			if (!parts[i] || !part_crc[i]) continue;
			for (int b = 0; b < pb; ++b) {
				if (lzo_crc32(0, parts[i] + (size_t)b * LZO_BLOCK_SIZE, LZO_BLOCK_SIZE) != part_crc[i][b]) {
					if (bad) { bad[0] = i; bad[1] = b; }
					return -3;
				}
			}
		}
	}
	if (kind == 0) {
		/* xor_read_plan.h:77-126: at most one part can be missing; it is the XOR of the rest */
		for (int w = 0; w < k + 1; ++w) {
			if (!want[w] || parts[w] || !out[w]) continue;
			int first = 1;
			for (int i = 0; i < k + 1; ++i) {
				if (i == w || !parts[i]) continue;
				if (first) { memcpy(out[w], parts[i], part_bytes); first = 0; }
				else lzo_block_xor(out[w], parts[i], part_bytes);
			}
		}
		return 0;
	}
	/* ec_read_plan.h:126-133: the first k available parts are used, the rest are erased */
	uint8_t erased[LZO_MAX_PARTS] = {0};
	const uint8_t *in[LZO_MAX_PARTS] = {0};
	uint8_t *dst[LZO_MAX_PARTS] = {0};
	int used = 0, n_out = 0;
	for (int i = 0; i < k + m; ++i) {
		if (!parts[i] || used >= k) erased[i] = 1;
		else { in[i] = parts[i]; ++used; }
	}
	/* ec_read_plan.h:139-143: only requested parts that were not read get an output */
	for (int i = 0; i < k + m; ++i)
		if (want[i] && !parts[i] && out[i]) { dst[i] = out[i]; ++n_out; }
	if (n_out == 0) return 0;
	return lzo_rs_recover(k, m, in, erased, dst, part_bytes);
}

/* chunk_read_planner.h:41-58: block b of the chunk comes from part b % k, index b / k */
void lzo_parts_to_chunk(int k, const uint8_t *const *data_parts, uint32_t nb, uint8_t *chunk) {
	for (uint32_t b = 0; b < nb; ++b)
		memcpy(chunk + (size_t)b * LZO_BLOCK_SIZE,
		       data_parts[b % (uint32_t)k] + (size_t)(b / (uint32_t)k) * LZO_BLOCK_SIZE, LZO_BLOCK_SIZE);
}

/* ------------------------------------------------------------------------------------
 * Replication: rebuild parts of slice type (dkind,dk,dm) from parts of slice type (skind,sk,sm).
 * Restates SliceRecoveryPlanner::prepare/buildPlan (src/chunkserver/slice_recovery_planner.h:87-204):
 *   same slice type            -> kReadDataPart: read the part, or rebuild it inside the slice
 *   data part of another type  -> kRecoverDataPart: chunk data, then BlockConverter (:41-57)
 *   parity part                -> kRecoverParityPart: chunk data, then XorReadPlan::RecoverParity
 *                                 (xor_read_plan.h:39-62) / ECReadPlan::RecoverParity (ec_read_plan.h:38-76)
 * and the per-block mycrc32 of ChunkReplicator::replicate (chunk_replicator.cc:186-192).
 * kind 2 = standard slice (k = 1, m = 0; part 0 is the chunk).  Output parts are pbd = ceil(nb/dk) blocks,
 * short data parts zero-padded.
 * ---------------------------------------------------------------------------------- */
int lzo_convert_chunk(int skind, int sk, int sm, const uint8_t *const *parts, const uint32_t *const *part_crc,
                      int dkind, int dk, int dm, const uint8_t *want, uint8_t *const *out, uint32_t *const *out_crc,
                      uint32_t nb, int *bad) {
	const size_t B = LZO_BLOCK_SIZE;
	const uint32_t pbs = (nb + (uint32_t)sk - 1) / (uint32_t)sk, pbd = (nb + (uint32_t)dk - 1) / (uint32_t)dk;
	int rc = 0;
	if (skind == dkind && sk == dk && sm == dm && skind != 2) {
		rc = lzo_recover_chunk(skind, sk, sm, parts, part_crc, want, out, (int)pbs, bad);
		if (rc) return rc;
		for (int i = 0; i < dk + dm; ++i)
			if (want[i] && parts[i] && out[i]) memcpy(out[i], parts[i], pbs * B);
	} else {
		/* chunk data in chunk order (ChunkReadPlanner + its BlockConverter, chunk_read_planner.h:36-70) */
		uint8_t *chunk = (uint8_t *)calloc((size_t)pbd * (size_t)dk + 1, B); /* zero tail = absent blocks of the last stripe */
		uint8_t *tmp[LZO_MAX_PARTS] = {0};
		if (!chunk) return -1;
		if (skind == 2) {
			if (!parts[0]) { free(chunk); return -2; }
			if (part_crc && part_crc[0])
				for (uint32_t b = 0; b < nb; ++b)
					if (lzo_crc32(0, parts[0] + b * B, LZO_BLOCK_SIZE) != part_crc[0][b]) {
						if (bad) { bad[0] = 0; bad[1] = (int)b; }
						free(chunk);
						return -3;
					}
			memcpy(chunk, parts[0], nb * B);
		} else {
			uint8_t all_data[LZO_MAX_PARTS] = {0};
			const uint8_t *data_parts[LZO_MAX_PARTS] = {0};
			for (int j = 0; j < sk; ++j) {
				all_data[j] = 1;
				if (!parts[j]) tmp[j] = (uint8_t *)calloc(pbs, B);
			}
			rc = lzo_recover_chunk(skind, sk, sm, parts, part_crc, all_data, tmp, (int)pbs, bad);
			if (rc == 0) {
				for (int j = 0; j < sk; ++j) data_parts[j] = parts[j] ? parts[j] : tmp[j];
				lzo_parts_to_chunk(sk, data_parts, nb, chunk);
			}
			for (int j = 0; j < sk; ++j) free(tmp[j]);
			if (rc) { free(chunk); return rc; }
		}
		if (dkind == 2) {
			if (want[0] && out[0]) memcpy(out[0], chunk, nb * B);
		} else {
			for (int j = 0; j < dk; ++j) {
				if (!want[j] || !out[j]) continue;
				/* SliceRecoveryPlanner::BlockConverter: dst block s <- chunk block s*dk + j */
				const uint32_t mine = (uint32_t)lzo_part_blocks(dk, j, nb);
				memset(out[j], 0, pbd * B);
				for (uint32_t s2 = 0; s2 < mine; ++s2) memcpy(out[j] + s2 * B, chunk + ((size_t)s2 * dk + j) * B, B);
			}
			for (int r = 0; r < dm; ++r) {
				uint8_t *dst = out[dk + r];
				if (!want[dk + r] || !dst) continue;
				for (uint32_t s2 = 0; s2 < pbd; ++s2) {
					const uint8_t *src = chunk + (size_t)s2 * dk * B;
					if (dkind == 0) { /* XorReadPlan::RecoverParity: memcpy the first block, blockXor the rest */
						memcpy(dst + s2 * B, src, B);
						for (int i = 1; i < dk; ++i) lzo_block_xor(dst + s2 * B, src + i * B, B);
					} else { /* ECReadPlan::RecoverParity: rs.recover with every parity part erased, one output */
						const uint8_t *in[LZO_MAX_PARTS] = {0};
						uint8_t erased[LZO_MAX_PARTS] = {0};
						uint8_t *res[LZO_MAX_PARTS] = {0};
						for (int i = 0; i < dk; ++i) in[i] = src + i * B;
						for (int i = 0; i < dm; ++i) erased[dk + i] = 1;
						res[dk + r] = dst + s2 * B;
						lzo_rs_recover(dk, dm, in, erased, res, B);
					}
				}
			}
		}
		free(chunk);
	}
	if (out_crc)
		for (int i = 0; i < dk + dm; ++i)
			if (want[i] && out[i] && out_crc[i])
				for (uint32_t b = 0; b < (dkind == 2 ? nb : pbd); ++b) out_crc[i][b] = lzo_crc32(0, out[i] + b * B, LZO_BLOCK_SIZE);
	return 0;
}

/* ------------------------------------------------------------------------------------
 * Chunkserver scrub, hdd_int_test (src/chunkserver/hddspacemgr.cc:2148-2210) over an in-memory chunk file.
 * Interleaved format (chunk.h:40, chunk.cc:195-209): records of 4-byte big-endian CRC + 64 KiB; the reader applies
 * the sparse-block rule (hddspacemgr.cc:1766-1780).  MooseFS format (chunk.cc:126-190): 1 KiB signature, CRC table,
 * data from the header size on; no sparse rule (hddspacemgr.cc:1748-1764).  Returns 0 or -3 with *first_bad.
 * ---------------------------------------------------------------------------------- */
static uint32_t get_be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

int lzo_scrub_interleaved(const uint8_t *records, size_t n_blocks, int64_t *first_bad) {
	for (size_t b = 0; b < n_blocks; ++b) {
		const uint8_t *rec = records + b * (4 + (size_t)LZO_BLOCK_SIZE);
		uint32_t stored = get_be32(rec);
		lzo_recompute_crc_if_block_empty(rec + 4, &stored);
		if (stored != lzo_crc32(0, rec + 4, LZO_BLOCK_SIZE)) {
			if (first_bad) *first_bad = (int64_t)b;
			return -3;
		}
	}
	return 0;
}

size_t lzo_moosefs_header_size(int data_parts) {
	size_t max_blocks = (LZO_BLOCKS_IN_CHUNK + (size_t)data_parts - 1) / (size_t)data_parts; /* chunk.cc:74-77 */
	size_t required = 1024 + 4 * max_blocks;                                                   /* chunk.cc:171,175 */
	return data_parts == 1 ? required : (required + 4095) / 4096 * 4096;                       /* chunk.cc:176-180 */
}

int lzo_scrub_moosefs(const uint8_t *image, int data_parts, size_t n_blocks, int64_t *first_bad) {
	const uint8_t *crc_table = image + 1024; /* getCrcOffset, chunk.cc:183-185 */
	const uint8_t *data = image + lzo_moosefs_header_size(data_parts);
	for (size_t b = 0; b < n_blocks; ++b)
		if (get_be32(crc_table + 4 * b) != lzo_crc32(0, data + b * (size_t)LZO_BLOCK_SIZE, LZO_BLOCK_SIZE)) {
			if (first_bad) *first_bad = (int64_t)b;
			return -3;
		}
	return 0;
}

/* ------------------------------------------------------------------------------------
 * hdd_write of one block (src/chunkserver/hddspacemgr.cc:1898-2008), the CRC arithmetic only.
 *   block      the 64 KiB block as stored (modified in place), or NULL when the block does not exist yet
 *              (blocknum >= chunk->blocks: it is created as zeros, :1976-1993)
 *   stored_crc in: CRC stored for the block (ignored when block == NULL); out: the CRC to store
 *   offset,size,crc,buffer  the write request (LIZ_CLTOCS_WRITE_DATA fields)
 *   new_block  when block == NULL, receives the created block (64 KiB)
 * Returns 0, -3 when the packet CRC is wrong (LIZARDFS_ERROR_CRC, :1916-1918), -4 when the stored block fails its CRC
 * check (:1962-1971), -1 on bad offset/size (:1910-1915).
 * ---------------------------------------------------------------------------------- */
int lzo_hdd_write_block(uint8_t *block, uint32_t *stored_crc, uint32_t offset, uint32_t size, uint32_t crc,
                        const uint8_t *buffer, uint8_t *new_block) {
	const uint32_t B = LZO_BLOCK_SIZE;
	uint32_t precrc, postcrc, combined;
	if (size > B || offset >= B || offset + size > B) return -1;
	if (crc != lzo_crc32(0, buffer, size)) return -3;
	if (offset == 0 && size == B) {
		uint8_t *dst = block ? block : new_block;
		memcpy(dst, buffer, B);
		*stored_crc = crc;
		return 0;
	}
	if (block) {
		uint32_t have = *stored_crc;
		lzo_recompute_crc_if_block_empty(block, &have); /* hdd_int_read_block_and_crc, interleaved format, :1779 */
		precrc = lzo_crc32(0, block, offset);
		uint32_t chcrc = lzo_crc32(0, block + offset, size);
		postcrc = lzo_crc32(0, block + offset + size, B - (offset + size));
		if (offset == 0) combined = lzo_crc32_combine(chcrc, postcrc, B - (offset + size));
		else {
			combined = lzo_crc32_combine(precrc, chcrc, size);
			if (offset + size < B) combined = lzo_crc32_combine(combined, postcrc, B - (offset + size));
		}
		if (have != combined) return -4;
	} else {
		block = new_block;
		memset(block, 0, B);
		precrc = lzo_crc32_zeroblock(0, offset);
		postcrc = lzo_crc32_zeroblock(0, B - (offset + size));
	}
	if (offset == 0) combined = lzo_crc32_combine(crc, postcrc, B - (offset + size));
	else {
		combined = lzo_crc32_combine(precrc, crc, size);
		if (offset + size < B) combined = lzo_crc32_combine(combined, postcrc, B - (offset + size));
	}
	memcpy(block + offset, buffer, size);
	*stored_crc = combined;
	return 0;
}

/* cltocs.h:116-137: serializePacketPrefix(destination, size, LIZ_CLTOCS_WRITE_DATA, 0, chunkId, writeId, block,
 * offset, size, crc) — PacketHeader(type, length = serialized size of version + fields (30) + size), big-endian. */
static uint8_t *put_be(uint8_t *p, uint64_t v, int bytes) {
	for (int i = bytes - 1; i >= 0; --i) *p++ = (uint8_t)(v >> (8 * i));
	return p;
}

void lzo_write_data_prefix(uint8_t *out, uint64_t chunk_id, uint32_t write_id, uint16_t block, uint32_t offset,
                           uint32_t size, uint32_t crc) {
	uint8_t *p = out;
	p = put_be(p, 1212u, 4);           /* LIZ_CLTOCS_WRITE_DATA = 1000 + 212 */
	p = put_be(p, 30u + size, 4);      /* 4 + 8 + 4 + 2 + 4 + 4 + 4 = kPrefixSize, plus the payload */
	p = put_be(p, 0u, 4);              /* packet version */
	p = put_be(p, chunk_id, 8);
	p = put_be(p, write_id, 4);
	p = put_be(p, block, 2);
	p = put_be(p, offset, 4);
	p = put_be(p, size, 4);
	p = put_be(p, crc, 4);
}

/* splitmix64 counter stream (not from the reference: SURVEY.md §8d asks for a generator
 * reproducible on CPU and GPU; seed 1, chunk 0 reproduces the survey's known answers). */
static uint64_t splitmix(uint64_t z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

void lzo_fill_chunk(uint8_t *dst, size_t len, uint64_t seed, uint64_t chunk_index) {
	const uint64_t golden = 0x9E3779B97F4A7C15ull;
	uint64_t base = seed + (chunk_index << 23) * golden;
	for (size_t off = 0; off < len; off += 8) {
		uint64_t w = splitmix(base + (uint64_t)(off / 8 + 1) * golden);
		size_t n = len - off < 8 ? len - off : 8;
		for (size_t i = 0; i < n; ++i) dst[off + i] = (uint8_t)(w >> (8 * i));
	}
}
