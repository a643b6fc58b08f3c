// lzgpu_reed_solomon.hpp — drop-in C++ mirror of the reference's ReedSolomon<MAXK, MAXM>
// (src/common/reed_solomon.h:41-373) on top of the C ABI (lzgpu.h).  Same public types and call
// signatures, so src/mount/chunk_writer.cc:386-400, src/common/ec_read_plan.h:40-63,115-145 and
// src/unittests/plan_tester.cc:281-300 compile unchanged when this header shadows common/reed_solomon.h.
// All arithmetic on fragment bytes happens on the GPU; the object itself is stateless (the reference's
// 34 KiB table cache is unnecessary: coefficient rows are recomputed per call on the host, O(k^3) bytes).
#pragma once
#include <array>
#include <bitset>
#include <cassert>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "lzgpu.h"

template <int MAXK, int MAXM>
class ReedSolomon {
public:
	static const int kMaxDataCount = MAXK;
	static const int kMaxParityCount = MAXM;
	static const int kMaxPartCount = MAXK + MAXM;

	typedef std::bitset<kMaxPartCount> ErasedMap;
	typedef std::array<uint8_t *, kMaxPartCount> FragmentMap;
	typedef std::array<const uint8_t *, kMaxPartCount> ConstFragmentMap;

	ReedSolomon() : rs_k_(), rs_m_() {}
	ReedSolomon(int k, int m) : rs_k_(k), rs_m_(m) {
		assert(k >= 1 && k <= kMaxDataCount);
		assert(m >= 1 && m <= kMaxParityCount);
	}

	// reed_solomon.h:87-121
	void recover(const ConstFragmentMap &input_fragments, const ErasedMap &erased, FragmentMap &output_fragments,
	             std::size_t data_size) {
		assert((int)erased.count() == rs_m_);
		uint8_t flags[LZGPU_MAX_PARTS] = {0};
		const uint8_t *in[LZGPU_MAX_PARTS] = {nullptr};
		uint8_t *out[LZGPU_MAX_PARTS] = {nullptr};
		for (int i = 0; i < rs_k_ + rs_m_; ++i) {
			flags[i] = erased[i] ? 1 : 0;
			in[i] = input_fragments[i];
			out[i] = output_fragments[i];
		}
		check(lzgpu_rs_recover(rs_k_, rs_m_, in, flags, out, data_size), "ReedSolomon::recover");
	}

	// reed_solomon.h:134-155
	void encode(const ConstFragmentMap &data_fragments, FragmentMap &parity_fragments, std::size_t data_size) {
		const uint8_t *in[LZGPU_MAX_DATA] = {nullptr};
		uint8_t *out[LZGPU_MAX_PARITY] = {nullptr};
		for (int i = 0; i < rs_k_; ++i) in[i] = data_fragments[i];
		for (int i = 0; i < rs_m_; ++i) {
			assert(parity_fragments[i]);
			out[i] = parity_fragments[i];
		}
		check(lzgpu_rs_encode(rs_k_, rs_m_, in, out, data_size), "ReedSolomon::encode");
	}

private:
	// the reference has void signatures and asserts; a GPU failure must not go unnoticed
	static void check(int rc, const char *what) {
		if (rc != LZGPU_OK) {
			std::fprintf(stderr, "%s failed on the GPU engine (status %d): %s\n", what, rc, lzgpu_last_error());
			std::abort();
		}
	}
	int rs_k_, rs_m_;
};
