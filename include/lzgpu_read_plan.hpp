// lzgpu_read_plan.hpp — GPU-backed mirror of the reference's read-plan post-processing (SURVEY.md §8 b, f1).
//
// The hook point of a degraded read is ReadPlan::postProcessData(uint8_t *buffer, const PartsContainer &available)
// (src/common/read_plan.h:141-160): the executor has filled the read buffer according to the plan's read_operations, then
//   1. SliceReadPlan::postProcessRead zero-fills the tail of short requested parts (src/common/slice_read_plan.h:94-105),
//   2. XorReadPlan / ECReadPlan::postProcessRead rebuild the requested parts that were not read
//      (src/common/xor_read_plan.h:77-126, src/common/ec_read_plan.h:88-146), IN PLACE in the buffer,
//   3. the post-process functors run back to front; for a chunk read that is ChunkReadPlanner::BlockConverter
//      (src/common/chunk_read_planner.h:36-70): part-major -> chunk order.
// lzgpu::SliceReadPlan below has the same public fields as the reference's SliceReadPlan (+ the BlockConverter's
// parameters as plain members instead of an opaque functor) and the same two entry points with the same buffer contract,
// so a maintainer can copy the fields of a plan built by SliceReadPlanner / ChunkReadPlanner and call it instead.
// Step 2 — all the arithmetic — is ONE lzgpu_recover_chunks call; steps 1 and 3 are the reference's memset / memcpy loops.
// tests/cpp/test_read_plan.cc runs it on plans made by the reference's own planners and compares with the reference's
// own post-processing, byte for byte.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "lzgpu.h"

namespace lzgpu {

struct ReadOperation {  // ReadPlan::ReadOperation, read_plan.h:50-62
	int request_offset;
	int request_size;
	int buffer_offset;
	int wave;
};

struct RequestedPartInfo {  // SliceReadPlan::RequestedPartInfo, slice_read_plan.h:35-38
	int part;  // slice part number, the REFERENCE's numbering (xor: 0 = parity, 1..N data)
	int size;
};

class ChunkCrcException : public std::runtime_error {  // what the mount throws on a bad block (read_operation_executor.cc:262-264)
public:
	ChunkCrcException(const std::string &what, int bad_part, int bad_block) : std::runtime_error(what), part(bad_part), block(bad_block) {}
	int part, block;
};

struct SliceReadPlan {
	int slice_type = 0;  // Goal::Slice::Type value
	std::vector<std::pair<int, ReadOperation>> read_operations;  // (slice part, operation), read_plan.h:165
	std::vector<RequestedPartInfo> requested_parts;              // slice_read_plan.h:110
	int buffer_part_size = 0;                                    // slice_read_plan.h:111
	int read_buffer_size = 0;                                    // read_plan.h:160

	// ChunkReadPlanner::BlockConverter (chunk_read_planner.h:36-70), set when the plan reads chunk blocks
	bool has_block_converter = false;
	int chunk_first_block = 0, chunk_block_count = 0, part_first_block = 0, part_block_count = 0, first_required_part = 0,
	    data_part_count = 0;

	int readOffset() const { return has_block_converter ? chunk_block_count * static_cast<int>(LZGPU_BLOCK_SIZE) : 0; }  // read_plan.h:73-80
	int fullBufferSize() const { return read_buffer_size + readOffset(); }

	// {Xor,EC}ReadPlan::postProcessRead: `buffer` is the READ buffer; `available_parts` holds the slice part numbers that were
	// read.  Optional stored CRCs (per read operation, `crc_of_part[slice part]` = CRCs of the blocks of that read, or nullptr)
	// are verified in the same GPU pass (the per-block check of read_operation_executor.cc:257-269).
	int postProcessRead(lzgpu_ctx *ctx, uint8_t *buffer, const std::vector<int> &available_parts,
	                    const uint32_t *const *crc_of_part = nullptr) const {
		lzgpu_goal goal;
		if (lzgpu_goal_from_slice_type(slice_type, &goal) != LZGPU_OK) throw std::invalid_argument("SliceReadPlan: not an xor/ec slice type");
		const int B = static_cast<int>(LZGPU_BLOCK_SIZE), k = goal.k, n = goal.k + goal.m;
		if (buffer_part_size <= 0 || buffer_part_size % B) throw std::invalid_argument("SliceReadPlan: buffer_part_size must be whole blocks");
		// 1. slice_read_plan.h:94-105
		int part_offset = 0;
		for (const RequestedPartInfo &info : requested_parts) {
			std::memset(buffer + part_offset + info.size, 0, buffer_part_size - info.size);
			part_offset += buffer_part_size;
		}
		const int result = static_cast<int>(requested_parts.size()) * buffer_part_size;
		// 2. which requested parts were not read?
		bool available[LZGPU_MAX_PARTS] = {false};
		for (int p : available_parts) available[api_part(goal, p)] = true;
		uint8_t want[LZGPU_MAX_PARTS] = {0};
		uint8_t *out[LZGPU_MAX_PARTS] = {nullptr};
		bool any_missing = false;
		for (size_t i = 0; i < requested_parts.size(); ++i) {
			const int a = api_part(goal, requested_parts[i].part);
			if (available[a]) continue;
			want[a] = 1;
			out[a] = buffer + i * static_cast<size_t>(buffer_part_size);
			any_missing = true;
		}
		if (!any_missing && !crc_of_part) return result;
		const uint8_t *parts[LZGPU_MAX_PARTS] = {nullptr};
		const uint32_t *crcs[LZGPU_MAX_PARTS] = {nullptr};
		for (const auto &op : read_operations) {
			const int a = api_part(goal, op.first);
			if (!available[a]) continue;
			parts[a] = buffer + op.second.buffer_offset;  // ec_read_plan.h:135-137
			if (crc_of_part) crcs[a] = crc_of_part[op.first];
		}
		const uint32_t pb = static_cast<uint32_t>(buffer_part_size / B);
		int64_t bad[3] = {-1, -1, -1};
		const int rc = lzgpu_recover_chunks(ctx, &goal, 1, pb * k, parts, buffer_part_size, crc_of_part ? crcs : nullptr, want, out, nullptr, 0, bad);
		if (rc == LZGPU_ERR_CRC)
			throw ChunkCrcException(lzgpu_last_error(), lzgpu_ref_part_index(&goal, static_cast<int>(bad[1])), static_cast<int>(bad[2]));
		if (rc != LZGPU_OK) throw std::runtime_error(std::string("SliceReadPlan::postProcessRead: ") + lzgpu_last_error());
		(void)n;
		return result;
	}

	// ReadPlan::postProcessData (read_plan.h:141-160): `buffer` is the FULL buffer (post-process area first, read buffer after)
	int postProcessData(lzgpu_ctx *ctx, uint8_t *buffer, const std::vector<int> &available_parts,
	                    const uint32_t *const *crc_of_part = nullptr) const {
		uint8_t *read_buffer = buffer + readOffset();
		int size = postProcessRead(ctx, read_buffer, available_parts, crc_of_part);
		if (!has_block_converter) return size;
		// 3. BlockConverter: chunk block chunk_first_block + i lives in requested part (b % k - first_required_part) at block b / k
		const size_t B = LZGPU_BLOCK_SIZE;
		uint8_t *dst = buffer;
		for (int i = 0; i < chunk_block_count; ++i) {
			const int block = (chunk_first_block + i) / data_part_count - part_first_block;
			int part = (chunk_first_block + i) % data_part_count - first_required_part;
			if (part < 0) part += data_part_count;
			std::memcpy(dst, read_buffer + (static_cast<size_t>(part) * part_block_count + block) * B, B);
			dst += B;
		}
		return chunk_block_count * static_cast<int>(B);
	}

private:
	// reference slice part number -> this library's part index (data 0..k-1, parity k..): xor keeps its parity in part 0
	static int api_part(const lzgpu_goal &g, int ref_part) {
		if (g.kind == LZGPU_KIND_XOR) return ref_part == 0 ? g.k : ref_part - 1;
		return ref_part;
	}
};

}  // namespace lzgpu
