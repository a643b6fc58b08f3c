// oracle_backend.cc — TEST-ONLY stand-in for the few batched entry points of lzgpu.h that the C++ host-side headers call
// (include/lzgpu_stripe_batcher.hpp, include/lzgpu_read_plan.hpp), implemented with the CPU oracle (oracle/lzoracle.h).
//
// Purpose: the `-m "not gpu"` suite exercises the HOST LOGIC of those headers — stripe-slot bookkeeping, write ids, packet
// prefixes, the reference<->API part numbering, the read-buffer layout — on a machine without a GPU, by linking the same
// test sources (test_stripe_batcher.cc, test_read_plan.cc) against this file instead of liblzgpu.so.  The goal / geometry
// helpers come from the real csrc/host_math.cc (pure host code).  This is never part of the product: liblzgpu.so has no CPU
// path and fails loudly without a device; the GPU runs of the same tests link the real library.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "lzgpu.h"
#include "../../oracle/lzoracle.h"

extern "C" {

static int g_dummy_ctx;

lzgpu_ctx *lzgpu_default_ctx(void) { return reinterpret_cast<lzgpu_ctx *>(&g_dummy_ctx); }
const char *lzgpu_last_error(void) { return "oracle backend (tests only)"; }

int lzgpu_host_alloc(lzgpu_ctx *, size_t bytes, void **h_ptr) {
	*h_ptr = std::malloc(bytes ? bytes : 1);
	return *h_ptr ? LZGPU_OK : LZGPU_ERR_NOMEM;
}
int lzgpu_host_free(lzgpu_ctx *, void *h_ptr) {
	std::free(h_ptr);
	return LZGPU_OK;
}

int lzgpu_test_fail_next_encode = 0;  // failure injection for the host-logic tests

int lzgpu_encode_chunks(lzgpu_ctx *, const lzgpu_goal *goal, uint32_t n_chunks, uint32_t chunk_len, const uint8_t *data, size_t chunk_stride,
                        uint8_t *parity, size_t parity_stride, uint32_t *crc, size_t crc_stride) {
	if (lzgpu_test_fail_next_encode) {
		lzgpu_test_fail_next_encode = 0;
		return LZGPU_ERR_CUDA;
	}
	for (uint32_t c = 0; c < n_chunks; ++c)
		if (lzo_encode_chunk(goal->kind, goal->k, goal->m, data + c * chunk_stride, chunk_len, parity + c * parity_stride, crc + c * crc_stride))
			return LZGPU_ERR_ARG;
	return LZGPU_OK;
}

int lzgpu_rs_recover(int k, int m, const uint8_t *const *in, const uint8_t *erased, uint8_t *const *out, size_t size) {
	return lzo_rs_recover(k, m, in, erased, out, size) ? LZGPU_ERR_ARG : LZGPU_OK;
}

int lzgpu_recover_chunks(lzgpu_ctx *, const lzgpu_goal *goal, uint32_t n_chunks, uint32_t nb, const uint8_t *const *parts, size_t part_stride,
                         const uint32_t *const *part_crc, const uint8_t *want, uint8_t *const *out, uint8_t *chunk_out, size_t chunk_out_stride,
                         int64_t *bad) {
	const int n = goal->k + goal->m;
	const int pb = static_cast<int>((nb + goal->k - 1) / goal->k);
	for (uint32_t c = 0; c < n_chunks; ++c) {
		const uint8_t *p[LZO_MAX_PARTS] = {nullptr};
		const uint32_t *pc[LZO_MAX_PARTS] = {nullptr};
		uint8_t *o[LZO_MAX_PARTS] = {nullptr};
		for (int i = 0; i < n; ++i) {
			if (parts[i]) p[i] = parts[i] + c * part_stride;
			if (part_crc && part_crc[i]) pc[i] = part_crc[i] + static_cast<size_t>(c) * pb;
			if (out && out[i]) o[i] = out[i] + c * part_stride;
		}
		int where[2] = {-1, -1};
		const int rc = lzo_recover_chunk(goal->kind, goal->k, goal->m, p, part_crc ? pc : nullptr, want, o, pb, where);
		if (rc == -3) {
			if (bad) { bad[0] = c; bad[1] = where[0]; bad[2] = where[1]; }
			return LZGPU_ERR_CRC;
		}
		if (rc == -2) return LZGPU_ERR_TOO_FEW_PARTS;
		if (rc) return LZGPU_ERR_ARG;
		if (chunk_out) {
			const uint8_t *dp[LZO_MAX_PARTS];
			for (int j = 0; j < goal->k; ++j) dp[j] = p[j] ? p[j] : o[j];
			lzo_parts_to_chunk(goal->k, dp, nb, chunk_out + c * chunk_out_stride);
		}
	}
	return LZGPU_OK;
}

}  // extern "C"
