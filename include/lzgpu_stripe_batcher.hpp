// lzgpu_stripe_batcher.hpp — the mount's write path, batched.
//
// The reference's ChunkWriter (src/mount/chunk_writer.cc) turns every stripe of the write journal into one
// "operation": startOperation (:475-547) completes the stripe (fillStripe, :437-466: blocks that were not written
// are READ from the chunkservers), computes each parity block with computeParityBlock (:365-401, one ReedSolomon
// object and one pass over the k data blocks per parity block) and hands every block to
// WriteExecutor::addDataPacket (src/common/write_executor.cc:91-107), which CRCs it (mycrc32, :97) and serialises
// the LIZ_CLTOCS_WRITE_DATA prefix (src/protocol/cltocs.h:116-137).
//
// StripeBatcher keeps the same unit — the stripe — and the same contract — a stripe is encoded only when all of
// its k blocks are present, blocks that were read back are not sent again — but collects the complete stripes of
// any number of chunks in page-locked memory and encodes them in ONE lzgpu_encode_chunks call (each stripe is a
// k-block mini chunk, the flat-unit path of the fused kernel): parity of every part, CRC of every data and parity
// block, and the 38-byte packet prefixes come back together; the sink receives exactly what addDataPacket would
// have been given, plus the CRC and the finished prefix, so sending is a pointer hand-off.
//
// Sub-block operations (WriteCacheBlock::from / to, src/mount/write_cache_block.cc:64-68; startOperation takes block_from /
// block_to / block_size from the first block of the stripe and gives the parity blocks the same range,
// chunk_writer.cc:479-481,522-529) are batched as well: a stripe whose blocks cover [from, to) is staged as whole 64 KiB blocks
// that are zero outside the range.  GF(2^8) parity is byte-wise, so the whole-block parity is the range's parity inside
// [from, to) and zero outside; the CRC of the `to - from` bytes that travel comes from the whole-block CRC through the
// concatenation identity run backwards (lzgpu_mycrc32_subrange, host scalar).  The sink then receives offset = from,
// size = to - from, data = block + from and a prefix carrying that offset and size — what addDataPacket(writeId, block,
// from, size, data) is given in the reference.  The per-call path (computeParityBlock below) remains for callers that want a
// single stripe encoded at once.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "lzgpu.h"

namespace lzgpu {

// what WriteExecutor::addDataPacket receives (write id, block of the PART, offset 0, size 64 KiB, data), plus the results
struct PartBlock {
	uint64_t chunk_id;
	int part;                  // this API's numbering: data 0..k-1, parity k..k+m-1 (lzgpu_ref_part_index converts)
	uint32_t block;            // block index inside the part = stripe index (blockIndex / data_part_count, chunk_writer.cc:541)
	uint32_t write_id;
	uint32_t offset, size;     // byte range inside the block: 0 / 64 KiB for whole blocks, from / to - from for a sub-block stripe
	const uint8_t *data;       // `size` bytes (the block's bytes from `offset` on), valid until the next flush()/addBlock()
	uint32_t crc;              // mycrc32(0, data, size)
	const uint8_t *prefix;     // LZGPU_WRITE_PREFIX_SIZE bytes, ready to send in front of `data`
};

// ChunkWriter::computeParityBlock (src/mount/chunk_writer.cc:365-401) with the reference's arguments — the per-call path for
// what the batcher does not take (sub-block ranges: `size` < 64 KiB at the same offset in every block).  data_blocks[offset + i]
// is block i of the stripe, nullptr = a block that does not exist (zeros, :377,:396).  xorN: memcpy + blockXor == the all-ones row.
inline void computeParityBlock(const lzgpu_goal &goal, int parity_index, uint8_t *parity_block, const std::vector<uint8_t *> &data_blocks,
                               int offset, int size) {
	const uint8_t *in[LZGPU_MAX_PARTS] = {nullptr};
	uint8_t erased[LZGPU_MAX_PARTS] = {0};
	uint8_t *out[LZGPU_MAX_PARTS] = {nullptr};
	for (int i = 0; i < goal.k; ++i) in[i] = data_blocks[offset + i];
	for (int i = 0; i < goal.m; ++i) erased[goal.k + i] = 1;  // rs.recover with every parity part erased, one output (:386-400)
	out[goal.k + parity_index] = parity_block;
	if (lzgpu_rs_recover(goal.k, goal.m, in, erased, out, static_cast<size_t>(size)) != LZGPU_OK)
		throw std::runtime_error(std::string("computeParityBlock: ") + lzgpu_last_error());
}

class StripeBatcher {
public:
	typedef std::function<void(const PartBlock &)> Sink;

	StripeBatcher(lzgpu_ctx *ctx, const lzgpu_goal &goal, uint32_t max_stripes)
	    : ctx_(ctx), goal_(goal), k_(goal.k), m_(goal.m), capacity_(max_stripes) {
		if (!ctx || !lzgpu_goal_valid(&goal) || max_stripes == 0) throw std::invalid_argument("StripeBatcher: bad arguments");
		const size_t B = LZGPU_BLOCK_SIZE;
		alloc(reinterpret_cast<void **>(&data_), static_cast<size_t>(capacity_) * k_ * B);
		alloc(reinterpret_cast<void **>(&parity_), static_cast<size_t>(capacity_) * m_ * B);
		alloc(reinterpret_cast<void **>(&crc_), static_cast<size_t>(capacity_) * (k_ + m_) * sizeof(uint32_t));
		prefix_.resize(static_cast<size_t>(capacity_) * (k_ + m_) * LZGPU_WRITE_PREFIX_SIZE);
		slots_.reserve(capacity_);
	}
	~StripeBatcher() {
		lzgpu_host_free(ctx_, data_);
		lzgpu_host_free(ctx_, parity_);
		lzgpu_host_free(ctx_, crc_);
	}
	StripeBatcher(const StripeBatcher &) = delete;
	StripeBatcher &operator=(const StripeBatcher &) = delete;

	// ChunkWriter::addOperation for a whole block.  `read_back` marks a block fetched to complete a stripe
	// (WriteCacheBlock::kReadBlock): it takes part in the parity but is not handed to the sink (chunk_writer.cc:503-508).
	// A second write of the same block replaces the first.  Returns false when no stripe slot is free (flush first).
	bool addBlock(uint64_t chunk_id, uint32_t block_index, const uint8_t *data, bool read_back = false) {
		return addBlockRange(chunk_id, block_index, 0, LZGPU_BLOCK_SIZE, data, read_back);
	}

	// ChunkWriter::addOperation for bytes [from, to) of a block (`data` points at byte `from`).  Every block of a stripe must
	// carry the same range (the reference builds an operation from journal positions of one range, Operation::isExpandPossible):
	// a different range for a buffered stripe throws.
	bool addBlockRange(uint64_t chunk_id, uint32_t block_index, uint32_t from, uint32_t to, const uint8_t *data, bool read_back = false) {
		if (block_index >= LZGPU_BLOCKS_IN_CHUNK || !data || from >= to || to > LZGPU_BLOCK_SIZE)
			throw std::invalid_argument("StripeBatcher::addBlock: bad block or range");
		const Key key(chunk_id, block_index / k_);
		auto it = index_.find(key);
		if (it == index_.end()) {
			if (slots_.size() == capacity_) return false;
			Slot s;
			s.chunk_id = chunk_id;
			s.stripe = block_index / k_;
			// blocks past the end of the chunk do not exist: the last stripe of a chunk is complete without them
			// (range_end, chunk_writer.cc:449), and they enter the parity as zeros
			s.expected = std::min<uint32_t>(k_, LZGPU_BLOCKS_IN_CHUNK - s.stripe * k_);
			s.from = from;
			s.to = to;
			std::memset(slot_data(slots_.size()) + static_cast<size_t>(s.expected) * LZGPU_BLOCK_SIZE, 0,
			            static_cast<size_t>(k_ - s.expected) * LZGPU_BLOCK_SIZE);
			it = index_.emplace(key, static_cast<uint32_t>(slots_.size())).first;
			slots_.push_back(s);
		}
		Slot &s = slots_[it->second];
		if (s.from != from || s.to != to) throw std::invalid_argument("StripeBatcher::addBlock: the blocks of a stripe must cover the same byte range");
		const uint32_t j = block_index % k_;
		uint8_t *dst = slot_data(it->second) + static_cast<size_t>(j) * LZGPU_BLOCK_SIZE;
		if (from) std::memset(dst, 0, from);  // staged as a whole block that is zero outside the range
		std::memcpy(dst + from, data, to - from);
		if (to < LZGPU_BLOCK_SIZE) std::memset(dst + to, 0, LZGPU_BLOCK_SIZE - to);
		s.present |= 1ull << j;
		if (read_back) s.read_back |= 1ull << j;
		else s.read_back &= ~(1ull << j);
		return true;
	}

	// (chunk id, chunk block index) of every block still missing from a buffered stripe — what fillStripe would read
	std::vector<std::pair<uint64_t, uint32_t>> missingBlocks() const {
		std::vector<std::pair<uint64_t, uint32_t>> r;
		for (const Slot &s : slots_)
			for (uint32_t j = 0; j < s.expected; ++j)
				if (!(s.present >> j & 1)) r.emplace_back(s.chunk_id, s.stripe * k_ + j);
		return r;
	}

	size_t bufferedStripes() const { return slots_.size(); }

	// Encodes every COMPLETE stripe in one GPU call and hands its blocks to `sink`; incomplete stripes stay buffered.
	// Write ids are allocated consecutively from first_write_id (ChunkWriter::allocateId).  Returns the number of
	// stripes encoded; throws std::runtime_error on an engine failure (there is no CPU fallback).
	size_t flush(uint32_t first_write_id, const Sink &sink) {
		const size_t B = LZGPU_BLOCK_SIZE;
		// complete stripes to the front (stable for the incomplete ones)
		size_t n = 0;
		for (size_t i = 0; i < slots_.size(); ++i) {
			if (!complete(slots_[i])) continue;
			if (i != n) swap_slots(i, n);
			++n;
		}
		// the slots moved: the (chunk, stripe) -> slot map is rebuilt BEFORE anything can throw, so a failed encode leaves the
		// batcher consistent and a retry copies into the right stripe
		rebuild_index();
		if (n == 0) return 0;
		const uint32_t chunk_len = static_cast<uint32_t>(k_ * B);
		int rc = lzgpu_encode_chunks(ctx_, &goal_, static_cast<uint32_t>(n), chunk_len, data_, chunk_len, parity_, m_ * B, crc_, k_ + m_);
		if (rc != LZGPU_OK) throw std::runtime_error(std::string("StripeBatcher::flush: ") + lzgpu_last_error());
		uint32_t write_id = first_write_id;
		for (size_t i = 0; i < n; ++i) {
			const Slot &s = slots_[i];
			const uint32_t *crc = crc_ + i * (k_ + m_);
			for (int part = 0; part < k_ + m_; ++part) {
				if (part < k_ && (part >= static_cast<int>(s.expected) || (s.read_back >> part & 1))) continue;
				PartBlock pb;
				pb.chunk_id = s.chunk_id;
				pb.part = part;
				pb.block = s.stripe;
				pb.write_id = write_id++;
				pb.offset = s.from;
				pb.size = s.to - s.from;
				pb.data = (part < k_ ? slot_data(i) + part * B : parity_ + (i * m_ + (part - k_)) * B) + s.from;
				// layout of lzgpu_encode_chunks: k data CRCs in chunk order, then one per parity part (whole-block CRCs)
				pb.crc = (s.from == 0 && s.to == B) ? crc[part] : lzgpu_mycrc32_subrange(crc[part], s.from, s.to);
				uint8_t *px = prefix_.data() + (i * (k_ + m_) + part) * LZGPU_WRITE_PREFIX_SIZE;
				write_prefix(px, s.chunk_id, pb.write_id, static_cast<uint16_t>(s.stripe), pb.offset, pb.size, pb.crc);
				pb.prefix = px;
				sink(pb);
			}
		}
		// drop the encoded stripes, keep the rest (moved to the front)
		std::vector<Slot> rest(slots_.begin() + n, slots_.end());
		for (size_t i = 0; i < rest.size(); ++i) std::memmove(slot_data(i), slot_data(n + i), static_cast<size_t>(k_) * B);
		slots_.swap(rest);
		rebuild_index();
		return n;
	}

private:
	typedef std::pair<uint64_t, uint32_t> Key;  // (chunk id, stripe)
	struct Slot {
		uint64_t chunk_id = 0;
		uint32_t stripe = 0, expected = 0;
		uint32_t from = 0, to = LZGPU_BLOCK_SIZE;  // byte range every block of the stripe covers
		uint64_t present = 0, read_back = 0;  // bit j = data part j (k <= 32)
	};

	void rebuild_index() {
		index_.clear();
		for (size_t i = 0; i < slots_.size(); ++i) index_.emplace(Key(slots_[i].chunk_id, slots_[i].stripe), static_cast<uint32_t>(i));
	}

	void alloc(void **p, size_t bytes) {
		if (lzgpu_host_alloc(ctx_, bytes, p) != LZGPU_OK) throw std::runtime_error(std::string("StripeBatcher: ") + lzgpu_last_error());
	}
	uint8_t *slot_data(size_t i) { return data_ + i * static_cast<size_t>(k_) * LZGPU_BLOCK_SIZE; }
	bool complete(const Slot &s) const { return s.present == (s.expected == 64 ? ~0ull : (1ull << s.expected) - 1); }
	void swap_slots(size_t a, size_t b) {
		const size_t bytes = static_cast<size_t>(k_) * LZGPU_BLOCK_SIZE;
		scratch_.resize(bytes);
		std::memcpy(scratch_.data(), slot_data(a), bytes);
		std::memcpy(slot_data(a), slot_data(b), bytes);
		std::memcpy(slot_data(b), scratch_.data(), bytes);
		std::swap(slots_[a], slots_[b]);
	}
	// cltocs::writeData::serializePrefix (src/protocol/cltocs.h:116-137): header (type 1212, length 30 + size), version 0,
	// chunkId, writeId, block, offset, size, crc — big-endian
	static void write_prefix(uint8_t *p, uint64_t chunk_id, uint32_t write_id, uint16_t block, uint32_t offset, uint32_t size, uint32_t crc) {
		auto be = [&p](uint64_t v, int bytes) {
			for (int i = bytes - 1; i >= 0; --i) *p++ = static_cast<uint8_t>(v >> (8 * i));
		};
		be(1212, 4); be(30u + size, 4); be(0, 4); be(chunk_id, 8); be(write_id, 4); be(block, 2); be(offset, 4);
		be(size, 4); be(crc, 4);
	}

	lzgpu_ctx *ctx_;
	lzgpu_goal goal_;
	int k_, m_;
	uint32_t capacity_;
	uint8_t *data_ = nullptr, *parity_ = nullptr;
	uint32_t *crc_ = nullptr;
	std::vector<uint8_t> prefix_, scratch_;
	std::vector<Slot> slots_;
	std::map<Key, uint32_t> index_;
};

}  // namespace lzgpu
