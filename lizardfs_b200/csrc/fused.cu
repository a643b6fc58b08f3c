// fused.cu — placeholder hooks (the TMA-streamed fused kernels land here)
#include "engine_internal.h"
int lz_fused_init(lzgpu_ctx *) { return LZGPU_OK; }
void lz_fused_destroy(lzgpu_ctx *) {}
int lz_fused_encode(lzgpu_ctx *, const lzgpu_goal *, uint32_t, uint32_t, const void *, size_t, void *, size_t, void *, size_t, cudaStream_t) {
	return LZGPU_NOT_HANDLED;
}
int lz_fused_crc(lzgpu_ctx *, const void *, unsigned long long, unsigned long long, unsigned long long, void *, unsigned long long, cudaStream_t) {
	return LZGPU_NOT_HANDLED;
}
