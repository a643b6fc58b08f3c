// bitslice.cuh — GF(2^8) on BIT PLANES: the arithmetic of the three- and four-parity-row encoders (fused_kernel.cuh, items of W = 8 words).
//
// A group is 8 packed words (32 bytes) of ONE part.  bs_transpose turns it into 8 planes: plane i holds bit i of each of the 32
// bytes (plane i, byte lane B, bit w  <-  word w, byte lane B, bit i) with three rounds of masked exchanges between register pairs
// (2 shifts + 2 selects per pair and round: 48 instructions per group, and the same 48 take planes back to bytes — the map is an
// involution).  On planes, multiplying all 32 bytes by x is a renaming of registers plus three XORs (x^8 = x^4+x^3+x^2+1,
// reference galois_coeff.h:30-32), so a Horner step of generator row r (acc <- acc * 2^r ^ d, reference generator
// galois_field_isal.cc:53-69) costs 8 / 9 / 11 three-input XORs per GROUP for r = 1 / 2 / 3 — against 5 / 10 / 11 instructions per
// WORD on packed bytes (device_math.cuh gf_x2_add ..).  Per data word of an ec(8,4) encode: 6 (transpose in) + 4.5 (four rows) +
// 2.25 (three parity groups back to bytes, row 0 stays in bytes) = 12.75 instructions instead of 27.
//
// Everything here is plain C++ on uint32_t: the host build of the same functions is what tests/test_host_math.py checks
// against the table-free field arithmetic of host_math.cc (lzgpu_debug_bitslice_rows).
#pragma once
#include <cstdint>

#ifndef LZ_HD
#ifdef __CUDACC__
#define LZ_HD __host__ __device__
#else
#define LZ_HD
#endif
#endif

// the loops below must be unrolled in device code (register arrays, constant-bank masks); host compilers do not know the pragma
#ifdef __CUDA_ARCH__
#define LZ_UNROLL _Pragma("unroll")
#else
#define LZ_UNROLL
#endif

namespace lzd {

// logical right shift; the device build may take it from the FMA pipe (IMAD.HI by 2^(32-s)) instead of the ALU pipe's funnel shift
template <int S>
LZ_HD inline uint32_t bs_shr(uint32_t a) {
#if defined(__CUDA_ARCH__) && defined(LZ_BS_SHR_FMA)
	return __umulhi(a, 1u << (32 - S));
#else
	return a >> S;
#endif
}

// (x & MASK) | (y & ~MASK).  On the device ONE LOP3 with the mask as its immediate (written as PTX: from the C expression nvcc
// makes two, one per constant — 264 of the 755 LOP3 of an ec(8,4) item)
template <uint32_t MASK>
LZ_HD inline uint32_t bs_select(uint32_t x, uint32_t y) {
#ifdef __CUDA_ARCH__
	uint32_t d;
	asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(x), "r"(y), "n"(MASK));
	return d;
#else
	return (x & MASK) | (y & ~MASK);
#endif
}

// one masked exchange: the bits of `hi_reg` whose in-byte index has bit t clear trade places with the bits of `lo_reg` whose index
// has it set (S = 2^t, MASK = the positions with bit t clear)
template <int S, uint32_t MASK>
LZ_HD inline void bs_exchange(uint32_t &lo_reg, uint32_t &hi_reg) {
	const uint32_t a = lo_reg, b = hi_reg;
	lo_reg = bs_select<MASK>(a, b << S);
	hi_reg = bs_select<MASK>(bs_shr<S>(a), b);
}

// bytes <-> planes (an involution): x[w] bit (8 B + i)  <->  x[i] bit (8 B + w)
LZ_HD inline void bs_transpose(uint32_t (&x)[8]) {
	bs_exchange<1, 0x55555555u>(x[0], x[1]);
	bs_exchange<1, 0x55555555u>(x[2], x[3]);
	bs_exchange<1, 0x55555555u>(x[4], x[5]);
	bs_exchange<1, 0x55555555u>(x[6], x[7]);
	bs_exchange<2, 0x33333333u>(x[0], x[2]);
	bs_exchange<2, 0x33333333u>(x[1], x[3]);
	bs_exchange<2, 0x33333333u>(x[4], x[6]);
	bs_exchange<2, 0x33333333u>(x[5], x[7]);
	bs_exchange<4, 0x0F0F0F0Fu>(x[0], x[4]);
	bs_exchange<4, 0x0F0F0F0Fu>(x[1], x[5]);
	bs_exchange<4, 0x0F0F0F0Fu>(x[2], x[6]);
	bs_exchange<4, 0x0F0F0F0Fu>(x[3], x[7]);
}

// Horner steps on planes: a <- a * 2^R ^ d  (a, d: 8 planes each).  Written out so that every new plane is ONE three-input XOR
// (plus the shared sums of the planes that wrap around):
//   a * 2 = [a7, a0, a1^a7, a2^a7, a3^a7, a4, a5, a6]
//   a * 4 = [a6, a7, a0^a6, a1^a6^a7, a2^a6^a7, a3^a7, a4, a5]
//   a * 8 = [a5, a6, a5^a7, a0^a5^a6, a1^a5^a6^a7, a2^a6^a7, a3^a7, a4]
template <int R>
LZ_HD inline void bs_horner(uint32_t (&a)[8], const uint32_t (&d)[8]) {
	static_assert(R >= 1 && R <= 3, "generator rows 1..3");
	const uint32_t a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5], a6 = a[6], a7 = a[7];
	if (R == 1) {
		a[0] = a7 ^ d[0];
		a[1] = a0 ^ d[1];
		a[2] = a1 ^ a7 ^ d[2];
		a[3] = a2 ^ a7 ^ d[3];
		a[4] = a3 ^ a7 ^ d[4];
		a[5] = a4 ^ d[5];
		a[6] = a5 ^ d[6];
		a[7] = a6 ^ d[7];
	} else if (R == 2) {
		const uint32_t t67 = a6 ^ a7;
		a[0] = a6 ^ d[0];
		a[1] = a7 ^ d[1];
		a[2] = a0 ^ a6 ^ d[2];
		a[3] = a1 ^ t67 ^ d[3];
		a[4] = a2 ^ t67 ^ d[4];
		a[5] = a3 ^ a7 ^ d[5];
		a[6] = a4 ^ d[6];
		a[7] = a5 ^ d[7];
	} else {
		const uint32_t t67 = a6 ^ a7, t56 = a5 ^ a6, t567 = t67 ^ a5;
		a[0] = a5 ^ d[0];
		a[1] = a6 ^ d[1];
		a[2] = a5 ^ a7 ^ d[2];
		a[3] = a0 ^ t56 ^ d[3];
		a[4] = a1 ^ t567 ^ d[4];
		a[5] = a2 ^ t67 ^ d[5];
		a[6] = a3 ^ a7 ^ d[6];
		a[7] = a4 ^ d[7];
	}
}

// a <- a * 2^R on planes, nothing added (a column the degraded read does not have)
template <int R>
LZ_HD inline void bs_mulpow(uint32_t (&a)[8]) {
	static_assert(R >= 1 && R <= 3, "2, 4 or 8");
	const uint32_t a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4], a5 = a[5], a6 = a[6], a7 = a[7];
	if (R == 1) {
		a[0] = a7; a[1] = a0; a[2] = a1 ^ a7; a[3] = a2 ^ a7; a[4] = a3 ^ a7; a[5] = a4; a[6] = a5; a[7] = a6;
	} else if (R == 2) {
		const uint32_t t67 = a6 ^ a7;
		a[0] = a6; a[1] = a7; a[2] = a0 ^ a6; a[3] = a1 ^ t67; a[4] = a2 ^ t67; a[5] = a3 ^ a7; a[6] = a4; a[7] = a5;
	} else {
		const uint32_t t67 = a6 ^ a7, t56 = a5 ^ a6;
		a[0] = a5; a[1] = a6; a[2] = a5 ^ a7; a[3] = a0 ^ t56; a[4] = a1 ^ t67 ^ a5; a[5] = a2 ^ t67; a[6] = a3 ^ a7; a[7] = a4;
	}
}

// Multiplication by an arbitrary constant c on planes: the 8 x 8 bit matrix of x -> c x as 64 words that are all-ones or zero,
// mask[8 i + b] = bit i of c * 2^b (bs_mask_set, host).  out_i (^)= XOR_b mask[8 i + b] & in_b — on the device each term is ONE LOP3
// whose mask operand comes straight from the constant bank (the masks live in the kernel parameters): 64 instructions per
// product of 32 bytes, against ~20 per 4 bytes for the bit-plane multiply on packed words (device_math.cuh gf_mac).
LZ_HD inline void bs_mask_set(uint32_t (&mask)[64], uint32_t c) {
	uint32_t v = c & 0xffu;
LZ_UNROLL
	for (int b = 0; b < 8; ++b) {
LZ_UNROLL
		for (int i = 0; i < 8; ++i) mask[8 * i + b] = ((v >> i) & 1u) ? 0xffffffffu : 0u;
		v = ((v << 1) ^ ((v & 0x80u) ? 0x1du : 0u)) & 0xffu;
	}
}
template <bool ACCUMULATE>
LZ_HD inline void bs_mul_mask(uint32_t (&out)[8], const uint32_t (&in)[8], const uint32_t (&mask)[64]) {
LZ_UNROLL
	for (int i = 0; i < 8; ++i) {
		uint32_t o = ACCUMULATE ? out[i] : 0u;
LZ_UNROLL
		for (int b = 0; b < 8; ++b) o ^= mask[8 * i + b] & in[b];
		out[i] = o;
	}
}

// Three unknowns a < b < c of a Vandermonde code with parity rows 1, 2^j, 4^j (the elimination of fused_recover_kernel, E = 3), on
// planes: s0, s1, s2 are the syndromes, ta = A s0 and tb = A^2 s0 (A = 2^a) come from the caller (doublings or two more masked
// multiplies); masks: alpha, beta, gamma, delta of the kernel comment.  Results: da, db, dc (planes).
LZ_HD inline void bs_solve3(const uint32_t (&s0)[8], const uint32_t (&s1)[8], const uint32_t (&s2)[8], const uint32_t (&ta)[8], const uint32_t (&tb)[8],
                            const uint32_t (&m_alpha)[64], const uint32_t (&m_beta)[64], const uint32_t (&m_gamma)[64], const uint32_t (&m_delta)[64],
                            uint32_t (&da)[8], uint32_t (&db)[8], uint32_t (&dc)[8]) {
	uint32_t t1[8], t2[8];
LZ_UNROLL
	for (int i = 0; i < 8; ++i) { t1[i] = ta[i] ^ s1[i]; t2[i] = tb[i] ^ s2[i]; }
	bs_mul_mask<false>(dc, t2, m_alpha);
	bs_mul_mask<true>(dc, t1, m_beta);
	bs_mul_mask<false>(db, t1, m_gamma);
	bs_mul_mask<true>(db, dc, m_delta);
LZ_UNROLL
	for (int i = 0; i < 8; ++i) da[i] = s0[i] ^ db[i] ^ dc[i];
}

// The GF role of one item of the M-row encoder (M = 3, 4), as the kernel runs it: column j = k-1 .. 0 of the stripe arrives as 8
// packed words, row 0 accumulates on bytes, rows 1..M-1 on planes; bs_rows_finish turns the plane accumulators back into bytes.
template <int M>
struct BsRows {
	static_assert(M == 3 || M == 4, "three or four Vandermonde rows");
	uint32_t p0[8];          // row 0 (XOR), bytes
	uint32_t p[M - 1][8];    // rows 1..M-1, planes until bs_rows_finish
};
template <int M>
LZ_HD inline void bs_rows_clear(BsRows<M> &s) {
LZ_UNROLL
	for (int i = 0; i < 8; ++i) {
		s.p0[i] = 0;
LZ_UNROLL
		for (int r = 0; r < M - 1; ++r) s.p[r][i] = 0;
	}
}
template <int M>
LZ_HD inline void bs_rows_add_column(BsRows<M> &s, uint32_t (&v)[8]) {
LZ_UNROLL
	for (int i = 0; i < 8; ++i) s.p0[i] ^= v[i];
	bs_transpose(v);
	bs_horner<1>(s.p[0], v);
	bs_horner<2>(s.p[1], v);
	if constexpr (M == 4) bs_horner<3>(s.p[2], v);
}
template <int M>
LZ_HD inline void bs_rows_finish(BsRows<M> &s) {
LZ_UNROLL
	for (int r = 0; r < M - 1; ++r) bs_transpose(s.p[r]);
}
using BsRows4 = BsRows<4>;

}  // namespace lzd
