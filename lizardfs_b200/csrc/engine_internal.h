// engine_internal.h — private declarations shared by engine.cu and fused.cu
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <mutex>

#include "lzgpu.h"

#define LZGPU_NOT_HANDLED 1  // internal: the fused path does not specialise this shape, use the generic kernels

void lz_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                         \
	do {                                                                                       \
		cudaError_t e__ = (expr);                                                              \
		if (e__ != cudaSuccess) {                                                              \
			cudaGetLastError();                                                                \
			lz_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
			             cudaGetErrorString(e__));                                             \
			return LZGPU_ERR_CUDA;                                                             \
		}                                                                                      \
	} while (0)

constexpr size_t kHostTileBytes = size_t(128) << 20;  // staging tile of the host-pointer entry points
constexpr int kHostSlots = 3;                          // tiles in flight: H2D(t+1) | kernel(t) | D2H(t-1)
constexpr int kCoefSlots = 8;
enum ScratchSlot {
	kScratchIn0 = 0, kScratchIn1, kScratchIn2, kScratchPar0, kScratchPar1, kScratchPar2, kScratchCrc0, kScratchCrc1, kScratchCrc2, kScratchTmpCrc,
	kScratchCoef0, kScratchCoefLast = kScratchCoef0 + kCoefSlots - 1,
	kScratchFused0, kScratchFused1,
	kScratchTmpPart0, kScratchTmpPartLast = kScratchTmpPart0 + LZGPU_MAX_PARTS - 1,
	kScratchConvImage, kScratchConvPar, kScratchConvCrc,
	kScratchCount
};

struct ScratchBuf {
	void *ptr = nullptr;
	size_t size = 0;
};

struct FusedState;  // fused.cu

struct lzgpu_ctx {
	int device = 0;
	int sm_count = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t slot_stream[kHostSlots] = {nullptr, nullptr, nullptr};
	uint32_t *d_crc_tables = nullptr;
	unsigned long long *d_first_bad = nullptr, *h_first_bad = nullptr;
	ScratchBuf scratch[kScratchCount];
	unsigned coef_rr = 0;
	FusedState *fused = nullptr;
	lzgpu_stats stats{};
	std::mutex mu;
};

int lz_scratch(lzgpu_ctx *ctx, int slot, size_t bytes, void **out);

// generic GF dot product descriptor (see kernels_generic.cuh DotArgs)
struct DotDesc {
	const uint8_t *src[32];
	uint8_t *const *dst;
	unsigned n_src, n_dst;
	unsigned long long total_units;
	unsigned long long src_chunk_stride, src_block_stride, dst_chunk_stride, dst_block_stride;
	unsigned units_per_block, blocks_per_chunk, valid_k, valid_nb;
};
int lz_gf_dot(lzgpu_ctx *ctx, const DotDesc &d, const uint8_t *coef, cudaStream_t st);
int lz_crc_blocks(lzgpu_ctx *ctx, const void *base, unsigned long long n_blocks, unsigned long long blocks_per_chunk,
                  unsigned long long chunk_stride, unsigned long long block_stride, uint32_t len, void *out,
                  unsigned long long out_chunk_stride, cudaStream_t st);

// fused TMA-streamed kernels (fused.cu).  Return LZGPU_NOT_HANDLED when the shape is not specialised.
int lz_fused_init(lzgpu_ctx *ctx);
void lz_fused_destroy(lzgpu_ctx *ctx);
int lz_fused_encode(lzgpu_ctx *ctx, const lzgpu_goal *goal, uint32_t n_chunks, uint32_t nb, const void *d_data, size_t chunk_stride,
                    void *d_parity, size_t parity_stride, void *d_crc, size_t crc_stride, cudaStream_t st);
// Fused degraded read (verify + rebuild erased data parts + chunk-order image).  On success *handled = true and, when
// any part was verified, first-bad information sits in ctx->d_first_bad[0] encoded as (chunk*64 + part)*1024 + block.
int lz_fused_recover(lzgpu_ctx *ctx, const lzgpu_goal *goal, uint32_t n_chunks, uint32_t nb, const void *const *d_parts, size_t part_stride,
                     const void *const *d_part_crc, const uint8_t *want, void *const *d_out, void *d_chunk_out, size_t chunk_out_stride,
                     cudaStream_t st, bool *verifying);
// CRC of 64 KiB blocks: block (c, b) at base + c*chunk_stride + b*65536, out[c*out_chunk_stride + b]
int lz_fused_crc(lzgpu_ctx *ctx, const void *base, unsigned long long n_blocks, unsigned long long blocks_per_chunk,
                 unsigned long long chunk_stride, void *out, unsigned long long out_chunk_stride, cudaStream_t st);
