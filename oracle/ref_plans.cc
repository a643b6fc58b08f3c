/*
 * ref_plans.cc — runs the UNMODIFIED reference read/recovery PLANNERS in memory (test infrastructure).
 *
 * ref_shim.cc covers the arithmetic layer; this file covers the layer above it, the code that decides what the
 * degraded-read and replication paths compute and in which layout:
 *   ChunkReadPlanner            src/common/chunk_read_planner.h:31-200   (mount read of chunk blocks)
 *   SliceRecoveryPlanner        src/chunkserver/slice_recovery_planner.h:37-213 (replication: rebuild one part, possibly
 *                                                                         from another slice type)
 *   ReadPlan::postProcessData   src/common/read_plan.h:141-160            (XorReadPlan / ECReadPlan post-processing and
 *                                                                         the BlockConverter / RecoverParity functors)
 * The network executor is replaced by the same in-memory copy loop the reference's own test helper uses
 * (src/unittests/plan_tester.cc:155-184, 64-110): every planned read operation of every wave is served from the part
 * buffers the caller supplies, then the plan's own postProcessData runs.  Nothing here re-implements plan logic.
 *
 * Part numbering at this boundary is the REFERENCE's (xor: 0 = parity, 1..N data; ec: 0..k-1 data, k.. parity;
 * standard: part 0); slice types are Goal::Slice::Type values (standard 0, xorN 2+(N-2), ec 10+32(k-2)+(m-1)).
 */
#include "common/platform.h"

#include <stdexcept>
#include <functional>
#include <numeric>
#include <cstring>
#include <memory>
#include <vector>

#include "chunkserver/slice_recovery_planner.h"
#include "common/chunk_read_planner.h"
#include "common/crc.h"
#include "common/slice_traits.h"

namespace {

typedef ChunkReadPlanner::PartsContainer PartsContainer;

struct PartSource {
	int slice_type;
	int slice_part;
	const uint8_t *data;  // part-major bytes of this part, `bytes` long (shorter reads are zero-filled like a short part)
	size_t bytes;
};

// serve the plan's read operations from memory, wave by wave, then post-process (plan_tester.cc:155-184)
int run_plan(std::unique_ptr<ReadPlan> plan, const std::vector<PartSource> &sources, std::vector<uint8_t> &buffer) {
	buffer.assign(plan->fullBufferSize(), 0);
#ifndef NDEBUG
	plan->buffer_start = buffer.data();
	plan->buffer_read = buffer.data() + plan->readOffset();
	plan->buffer_end = buffer.data() + plan->fullBufferSize();
#endif
	PartsContainer available;
	for (int wave = 0; wave < 10; ++wave) {
		for (const auto &op : plan->read_operations) {
			if (op.second.wave != wave) continue;
			const PartSource *src = nullptr;
			for (const auto &s : sources)
				if (s.slice_type == (int)op.first.getSliceType() && s.slice_part == op.first.getSlicePart()) src = &s;
			if (!src) continue;  // unreachable part: the next wave reads the spare ones
			uint8_t *dst = buffer.data() + plan->readOffset() + op.second.buffer_offset;
			const size_t off = op.second.request_offset, size = op.second.request_size;
			if (off < src->bytes) std::memcpy(dst, src->data + off, std::min(size, src->bytes - off));
			available.push_back(op.first);
		}
		if (plan->isReadingFinished(available)) break;
	}
	if (!plan->isReadingFinished(available)) return -1;
	const int size = plan->postProcessData(buffer.data(), available);
	buffer.resize(size);
	return size;
}

PartsContainer parts_of(const std::vector<PartSource> &sources) {
	PartsContainer r;
	for (const auto &s : sources) r.push_back(ChunkPartType(Goal::Slice::Type(s.slice_type), s.slice_part));
	return r;
}

std::vector<PartSource> gather(int n, const int *types, const int *parts, const uint8_t *const *data, const size_t *bytes) {
	std::vector<PartSource> v;
	for (int i = 0; i < n; ++i) v.push_back(PartSource{types[i], parts[i], data[i], bytes[i]});
	return v;
}

}  // namespace

extern "C" {

/* Mount-side degraded read: blocks [first_block, first_block + block_count) of the chunk in chunk order from the available
 * parts, through ChunkReadPlanner::buildPlan and the plan's own post-processing.  Returns bytes written, -1 when the
 * reference says reading is impossible. */
long ref_plan_read_chunk(int n_avail, const int *types, const int *parts, const uint8_t *const *data, const size_t *bytes,
                         int first_block, int block_count, uint8_t *out) {
	std::vector<PartSource> sources = gather(n_avail, types, parts, data, bytes);
	ChunkReadPlanner planner;
	planner.prepare(first_block, block_count, parts_of(sources));
	if (!planner.isReadingPossible()) return -1;
	std::vector<uint8_t> buffer;
	const int size = run_plan(planner.buildPlan(), sources, buffer);
	if (size < 0) return -1;
	std::memcpy(out, buffer.data(), size);
	return size;
}

/* Chunkserver replication: blocks [first_block, +block_count) of part (dst_type, dst_part) from the available parts, via
 * SliceRecoveryPlanner (read the part / convert chunk data / recompute parity), followed by the replicator's per-block CRC
 * loop (src/chunkserver/chunk_replicator.cc:186-192).  Returns bytes written, -1 when impossible. */
long ref_plan_recover_part(int n_avail, const int *types, const int *parts, const uint8_t *const *data, const size_t *bytes,
                           int dst_type, int dst_part, int first_block, int block_count, uint8_t *out, uint32_t *out_crc) {
	std::vector<PartSource> sources = gather(n_avail, types, parts, data, bytes);
	SliceRecoveryPlanner planner;
	planner.prepare(ChunkPartType(Goal::Slice::Type(dst_type), dst_part), first_block, block_count, parts_of(sources));
	if (!planner.isReadingPossible()) return -1;
	std::vector<uint8_t> buffer;
	const int size = run_plan(planner.buildPlan(), sources, buffer);
	if (size < 0) return -1;
	std::memcpy(out, buffer.data(), size);
	if (out_crc)
		for (int i = 0; i < block_count; ++i) out_crc[i] = mycrc32(0, buffer.data() + (size_t)i * MFSBLOCKSIZE, MFSBLOCKSIZE);
	return size;
}

}  // extern "C"

/* ------------------------------------------------------------------------------------------------
 * The same chunk read, stopped between the network part and the post-processing: exports the plan the reference built
 * (the public fields of SliceReadPlan + the parameters ChunkReadPlanner::buildPlan gives its BlockConverter), the full
 * buffer as the executor leaves it, and the reference's own post-processed result.  Used to test
 * include/lzgpu_read_plan.hpp (the GPU-backed mirror of ReadPlan::postProcessData) on plans made by the reference.
 * desc layout (ints): slice_type, buffer_part_size, read_buffer_size, read_offset, n_requested, {part, size}*,
 *   n_ops, {slice_part, request_offset, request_size, buffer_offset, wave}*, has_converter, chunk_first_block,
 *   chunk_block_count, part_first_block, part_block_count, first_required_part, data_part_count, n_available, {slice_part}*
 * Returns the full buffer size, -1 when reading is impossible, -2 when a capacity is too small.
 * ---------------------------------------------------------------------------------------------- */
namespace {
struct PlannerProbe : ChunkReadPlanner {  // read access to what buildPlan() hands to the BlockConverter
	using ChunkReadPlanner::chunk_block_count_;
	using ChunkReadPlanner::chunk_first_block_;
	using ChunkReadPlanner::part_block_count_;
	using ChunkReadPlanner::part_first_block_;
	using ChunkReadPlanner::read_from_type_;
	using ChunkReadPlanner::read_parts_;
};
}  // namespace

extern "C" long ref_plan_chunk_read_staged(int n_avail, const int *types, const int *parts, const uint8_t *const *data, const size_t *bytes,
                                           int first_block, int block_count, int *desc, int desc_cap, uint8_t *staged, uint8_t *expected,
                                           size_t buffer_cap) {
	std::vector<PartSource> sources = gather(n_avail, types, parts, data, bytes);
	PlannerProbe planner;
	planner.prepare(first_block, block_count, parts_of(sources));
	if (!planner.isReadingPossible()) return -1;
	std::unique_ptr<ReadPlan> plan = planner.buildPlan();
	SliceReadPlan *sp = dynamic_cast<SliceReadPlan *>(plan.get());
	if (!sp) return -1;
	const size_t full = plan->fullBufferSize();
	if (full > buffer_cap) return -2;
	std::vector<int> d;
	d.push_back((int)sp->slice_type);
	d.push_back(sp->buffer_part_size);
	d.push_back(plan->read_buffer_size);
	d.push_back(plan->readOffset());
	d.push_back((int)sp->requested_parts.size());
	for (const auto &r : sp->requested_parts) { d.push_back(r.part); d.push_back(r.size); }
	d.push_back((int)plan->read_operations.size());
	for (const auto &op : plan->read_operations) {
		d.push_back(op.first.getSlicePart()); d.push_back(op.second.request_offset); d.push_back(op.second.request_size);
		d.push_back(op.second.buffer_offset); d.push_back(op.second.wave);
	}
	const bool conv = !plan->postprocess_operations.empty();
	d.push_back(conv ? 1 : 0);
	d.push_back(planner.chunk_first_block_); d.push_back(planner.chunk_block_count_);
	d.push_back(planner.part_first_block_); d.push_back(planner.part_block_count_);
	d.push_back(slice_traits::isXor(planner.read_from_type_) ? planner.read_parts_[0] - 1 : planner.read_parts_[0]);
	d.push_back(slice_traits::requiredPartsToRecover(planner.read_from_type_));

	// the executor part (plan_tester.cc:155-184), stopped before postProcessData
	std::vector<uint8_t> buffer(full, 0);
	PartsContainer available;
	for (int wave = 0; wave < 10; ++wave) {
		for (const auto &op : plan->read_operations) {
			if (op.second.wave != wave) continue;
			const PartSource *src = nullptr;
			for (const auto &s : sources)
				if (s.slice_type == (int)op.first.getSliceType() && s.slice_part == op.first.getSlicePart()) src = &s;
			if (!src) continue;
			uint8_t *dst = buffer.data() + plan->readOffset() + op.second.buffer_offset;
			const size_t off = op.second.request_offset, size = op.second.request_size;
			if (off < src->bytes) std::memcpy(dst, src->data + off, std::min(size, src->bytes - off));
			available.push_back(op.first);
		}
		if (plan->isReadingFinished(available)) break;
	}
	if (!plan->isReadingFinished(available)) return -1;
	d.push_back((int)available.size());
	for (const auto &a : available) d.push_back(a.getSlicePart());
	if ((int)d.size() > desc_cap) return -2;
	std::memcpy(desc, d.data(), d.size() * sizeof(int));
	std::memcpy(staged, buffer.data(), full);
#ifndef NDEBUG
	plan->buffer_start = buffer.data();
	plan->buffer_read = buffer.data() + plan->readOffset();
	plan->buffer_end = buffer.data() + full;
#endif
	const int size = plan->postProcessData(buffer.data(), available);
	std::memcpy(expected, buffer.data(), size);
	return (long)full;
}
