// GPU test of lzgpu::StripeBatcher (include/lzgpu_stripe_batcher.hpp), the batched form of the mount write path
// (ChunkWriter::startOperation, src/mount/chunk_writer.cc:475-547 + WriteExecutor::addDataPacket,
// src/common/write_executor.cc:91-107).  Blocks of several chunks arrive in random order; every block the sink receives
// is checked against the CPU oracle (oracle/lzoracle.h — the checker, linked by this test only): parity bytes of the
// stripe, mycrc32 of the block, the serialized packet prefix.  Exit code 0 = all passed.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <set>
#include <vector>

#include "lzgpu_stripe_batcher.hpp"
#include "../../oracle/lzoracle.h"

static int failures = 0;
#define EXPECT(cond)                                                        \
	do {                                                                    \
		if (!(cond)) {                                                      \
			std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
			++failures;                                                     \
		}                                                                   \
	} while (0)

static const size_t B = LZGPU_BLOCK_SIZE;

struct Written {
	int kind, k, m;
	std::map<std::pair<uint64_t, uint32_t>, std::vector<uint8_t>> blocks;  // (chunk, chunk block) -> bytes
};

// what the sink saw, copied (the batcher's buffers are recycled by the next flush)
struct Seen {
	lzgpu::PartBlock pb;
	std::vector<uint8_t> data, prefix;
};

static void check_stripe(const Written &w, uint64_t chunk, uint32_t stripe, const std::map<int, Seen> &got,
                         const std::set<uint32_t> &read_back) {
	// the stripe as a k-block mini chunk, absent tail blocks zero
	std::vector<uint8_t> data(w.k * B, 0), parity(w.m * B);
	std::vector<uint32_t> crc(w.k + w.m);
	int present = 0;
	for (int j = 0; j < w.k; ++j) {
		auto it = w.blocks.find({chunk, stripe * w.k + j});
		if (it == w.blocks.end()) continue;
		std::memcpy(&data[j * B], it->second.data(), B);
		++present;
	}
	EXPECT(lzo_encode_chunk(w.kind, w.k, w.m, data.data(), data.size(), parity.data(), crc.data()) == 0);
	size_t expected_blocks = 0;
	for (int part = 0; part < w.k + w.m; ++part) {
		const bool sent = part >= w.k || (w.blocks.count({chunk, stripe * w.k + part}) && !read_back.count(stripe * w.k + part));
		if (!sent) {
			EXPECT(!got.count(part));
			continue;
		}
		++expected_blocks;
		auto it = got.find(part);
		EXPECT(it != got.end());
		if (it == got.end()) continue;
		const lzgpu::PartBlock &pb = it->second.pb;
		const uint8_t *want = part < w.k ? &data[part * B] : &parity[(part - w.k) * B];
		EXPECT(pb.block == stripe && pb.chunk_id == chunk);
		EXPECT(std::memcmp(it->second.data.data(), want, B) == 0);
		EXPECT(pb.crc == crc[part]);
		EXPECT(pb.crc == lzo_crc32(0, want, LZGPU_BLOCK_SIZE));
		uint8_t prefix[LZO_WRITE_PREFIX_SIZE];
		lzo_write_data_prefix(prefix, chunk, pb.write_id, static_cast<uint16_t>(stripe), 0, LZGPU_BLOCK_SIZE, pb.crc);
		EXPECT(std::memcmp(it->second.prefix.data(), prefix, sizeof(prefix)) == 0);
	}
	EXPECT(got.size() == expected_blocks);
	(void)present;
}

static void run(const char *text, unsigned seed) {
	lzgpu_goal goal;
	EXPECT(lzgpu_goal_parse(text, &goal) == LZGPU_OK);
	std::mt19937_64 rng(seed);
	Written w{goal.kind, goal.k, goal.m, {}};
	lzgpu::StripeBatcher batcher(lzgpu_default_ctx(), goal, 24);

	// three chunks: whole stripes, the last stripe of a full chunk (k may not divide 1024), one stripe left incomplete
	struct Item { uint64_t chunk; uint32_t block; bool read_back; };
	std::vector<Item> items;
	const uint32_t last_stripe = (LZGPU_BLOCKS_IN_CHUNK - 1) / goal.k;
	for (uint32_t s : {0u, 1u, 5u})
		for (int j = 0; j < goal.k; ++j) items.push_back({0x1122334455667788ull, s * goal.k + j, s == 5 && j == 0});
	for (uint32_t b = last_stripe * goal.k; b < LZGPU_BLOCKS_IN_CHUNK; ++b) items.push_back({42, b, false});
	for (int j = 0; j < goal.k; ++j) items.push_back({7, 3u * goal.k + j, false});
	std::shuffle(items.begin(), items.end(), rng);
	// chunk 9, stripe 2: one block short until the second round
	const Item held{9, 2u * goal.k + (goal.k - 1), false};
	for (int j = 0; j + 1 < goal.k; ++j) items.push_back({9, 2u * goal.k + j, false});

	std::set<uint32_t> read_back_blocks_chunk0;
	for (const Item &it : items) {
		std::vector<uint8_t> blk(B);
		for (auto &x : blk) x = static_cast<uint8_t>(rng());
		if (it.chunk == 42 && it.block == LZGPU_BLOCKS_IN_CHUNK - 1) std::fill(blk.begin(), blk.end(), 0);  // a zero block
		EXPECT(batcher.addBlock(it.chunk, it.block, blk.data(), it.read_back));
		w.blocks[{it.chunk, it.block}] = blk;
		if (it.read_back) read_back_blocks_chunk0.insert(it.block);
	}
	// rewriting a block replaces it
	{
		std::vector<uint8_t> blk(B, 0x5a);
		EXPECT(batcher.addBlock(7, 3u * goal.k, blk.data()));
		w.blocks[{7, 3u * goal.k}] = blk;
	}
	auto missing = batcher.missingBlocks();
	EXPECT(missing.size() == 1 && missing[0].first == 9 && missing[0].second == held.block);

	std::map<std::pair<uint64_t, uint32_t>, std::map<int, Seen>> got;
	std::vector<uint32_t> ids;
	auto sink = [&](const lzgpu::PartBlock &pb) {
		std::map<int, Seen> &stripe = got[std::make_pair(pb.chunk_id, pb.block)];
		EXPECT(!stripe.count(pb.part));
		stripe[pb.part] = Seen{pb, std::vector<uint8_t>(pb.data, pb.data + B), std::vector<uint8_t>(pb.prefix, pb.prefix + LZGPU_WRITE_PREFIX_SIZE)};
		ids.push_back(pb.write_id);
	};
	const size_t n1 = batcher.flush(1000, sink);
	EXPECT(n1 == 5);                       // chunk 0: stripes 0,1,5; chunk 42: last stripe; chunk 7: stripe 3
	EXPECT(batcher.bufferedStripes() == 1);
	for (size_t i = 0; i < ids.size(); ++i) EXPECT(ids[i] == 1000 + i);  // consecutive write ids
	for (auto &kv : got) check_stripe(w, kv.first.first, kv.first.second, kv.second, kv.first.first == 0x1122334455667788ull ? read_back_blocks_chunk0 : std::set<uint32_t>());
	EXPECT(got.size() == 5);

	// second round: the held block arrives, the stripe completes
	got.clear();
	ids.clear();
	{
		std::vector<uint8_t> blk(B);
		for (auto &x : blk) x = static_cast<uint8_t>(rng());
		EXPECT(batcher.addBlock(held.chunk, held.block, blk.data()));
		w.blocks[{held.chunk, held.block}] = blk;
	}
	EXPECT(batcher.missingBlocks().empty());
	EXPECT(batcher.flush(5, sink) == 1 && batcher.bufferedStripes() == 0);
	EXPECT(got.size() == 1);
	for (auto &kv : got) check_stripe(w, kv.first.first, kv.first.second, kv.second, {});
	EXPECT(batcher.flush(0, sink) == 0);

	// capacity: the 25th distinct stripe is refused until a flush
	std::vector<uint8_t> blk(B, 1);
	for (uint32_t s = 0; s < 24; ++s) EXPECT(batcher.addBlock(100 + s, 0, blk.data()));
	EXPECT(!batcher.addBlock(999, 0, blk.data()));
	std::printf("%s: ok\n", text);
}

// the per-call mirror of ChunkWriter::computeParityBlock: sub-block sizes, absent (NULL) blocks, every parity part
static void test_compute_parity_block(const char *text) {
	lzgpu_goal goal;
	EXPECT(lzgpu_goal_parse(text, &goal) == LZGPU_OK);
	std::mt19937_64 rng(99);
	for (int size : {65536, 4096, 1, 12345}) {
		std::vector<std::vector<uint8_t>> blocks(2 + goal.k, std::vector<uint8_t>(size));
		std::vector<uint8_t *> ptrs(2 + goal.k, nullptr);
		for (int i = 0; i < goal.k; ++i) {
			for (auto &x : blocks[2 + i]) x = static_cast<uint8_t>(rng());
			ptrs[2 + i] = (i == goal.k - 1 && size != 65536) ? nullptr : blocks[2 + i].data();  // a block past the end of the file
		}
		for (int r = 0; r < goal.m; ++r) {
			std::vector<uint8_t> got(size, 0xEE), want(size);
			lzgpu::computeParityBlock(goal, r, got.data(), ptrs, 2, size);
			const uint8_t *in[LZO_MAX_PARTS] = {nullptr};
			uint8_t erased[LZO_MAX_PARTS] = {0};
			uint8_t *out[LZO_MAX_PARTS] = {nullptr};
			for (int i = 0; i < goal.k; ++i) in[i] = ptrs[2 + i];
			for (int i = 0; i < goal.m; ++i) erased[goal.k + i] = 1;
			out[goal.k + r] = want.data();
			EXPECT(lzo_rs_recover(goal.k, goal.m, in, erased, out, size) == 0);
			EXPECT(got == want);
		}
	}
	std::printf("computeParityBlock %s: ok\n", text);
}

// Sub-block stripes (WriteCacheBlock::from / to, chunk_writer.cc:479-481): every block of the stripe carries bytes [from, to);
// the sink must get exactly what addDataPacket(writeId, block, from, size, data) gets in the reference — the range's bytes, the
// parity of the range (ChunkWriter::computeParityBlock on `size` bytes), mycrc32 of those `size` bytes, a prefix with
// offset = from and size = to - from — batched with whole-block stripes in the same flush.
static void test_sub_block_stripes(const char *text) {
	lzgpu_goal goal;
	EXPECT(lzgpu_goal_parse(text, &goal) == LZGPU_OK);
	std::mt19937_64 rng(4242);
	lzgpu::StripeBatcher batcher(lzgpu_default_ctx(), goal, 16);
	struct Range { uint32_t from, to; };
	const Range ranges[] = {{0, 65536}, {0, 4096}, {4096, 65536}, {100, 101}, {12345, 54321}, {65535, 65536}};
	std::map<uint32_t, std::vector<std::vector<uint8_t>>> payload;  // stripe -> k payloads of to - from bytes
	uint32_t stripe = 0;
	for (const Range &r : ranges) {
		payload[stripe].resize(goal.k);
		for (int j = 0; j < goal.k; ++j) {
			payload[stripe][j].resize(r.to - r.from);
			for (auto &x : payload[stripe][j]) x = static_cast<uint8_t>(rng());
			EXPECT(batcher.addBlockRange(77, stripe * goal.k + j, r.from, r.to, payload[stripe][j].data()));
		}
		++stripe;
	}
	// a block with another range for a buffered stripe is refused
	bool threw = false;
	try {
		std::vector<uint8_t> x(10);
		batcher.addBlockRange(77, 1 * goal.k, 0, 10, x.data());
	} catch (const std::invalid_argument &) { threw = true; }
	EXPECT(threw);
	std::map<std::pair<uint32_t, int>, Seen> got;
	const size_t n = batcher.flush(9, [&](const lzgpu::PartBlock &pb) {
		got[{pb.block, pb.part}] = Seen{pb, std::vector<uint8_t>(pb.data, pb.data + pb.size), std::vector<uint8_t>(pb.prefix, pb.prefix + LZGPU_WRITE_PREFIX_SIZE)};
	});
	EXPECT(n == sizeof(ranges) / sizeof(ranges[0]));
	EXPECT(got.size() == n * (goal.k + goal.m));
	for (uint32_t s = 0; s < n; ++s) {
		const Range &r = ranges[s];
		const uint32_t size = r.to - r.from;
		// the reference's computeParityBlock on `size` bytes
		const uint8_t *in[LZO_MAX_PARTS] = {nullptr};
		uint8_t erased[LZO_MAX_PARTS] = {0};
		uint8_t *out[LZO_MAX_PARTS] = {nullptr};
		std::vector<std::vector<uint8_t>> par(goal.m, std::vector<uint8_t>(size));
		for (int j = 0; j < goal.k; ++j) in[j] = payload[s][j].data();
		for (int i = 0; i < goal.m; ++i) { erased[goal.k + i] = 1; out[goal.k + i] = par[i].data(); }
		EXPECT(lzo_rs_recover(goal.k, goal.m, in, erased, out, size) == 0);
		for (int part = 0; part < goal.k + goal.m; ++part) {
			auto it = got.find({s, part});
			EXPECT(it != got.end());
			if (it == got.end()) continue;
			const std::vector<uint8_t> &want = part < goal.k ? payload[s][part] : par[part - goal.k];
			const lzgpu::PartBlock &pb = it->second.pb;
			EXPECT(pb.offset == r.from && pb.size == size);
			EXPECT(it->second.data == want);
			EXPECT(pb.crc == lzo_crc32(0, want.data(), size));
			uint8_t prefix[LZO_WRITE_PREFIX_SIZE];
			lzo_write_data_prefix(prefix, 77, pb.write_id, static_cast<uint16_t>(s), r.from, size, pb.crc);
			EXPECT(std::memcmp(it->second.prefix.data(), prefix, sizeof(prefix)) == 0);
		}
	}
	std::printf("sub-block stripes %s: ok\n", text);
}

#ifdef LZ_TEST_CPU_BACKEND
extern "C" int lzgpu_test_fail_next_encode;  // oracle_backend.cc: makes the next lzgpu_encode_chunks call fail
// a failed flush must leave the batcher consistent: the slots were re-ordered (complete stripes first) before the encode, so the
// (chunk, stripe) -> slot map has to follow; the retry must then deliver every stripe with its own data
static void test_failed_flush_keeps_the_index() {
	lzgpu_goal goal;
	EXPECT(lzgpu_goal_parse("ec(3,2)", &goal) == LZGPU_OK);
	lzgpu::StripeBatcher batcher(lzgpu_default_ctx(), goal, 8);
	std::vector<std::vector<uint8_t>> blk(9, std::vector<uint8_t>(B));
	for (size_t i = 0; i < blk.size(); ++i) std::fill(blk[i].begin(), blk[i].end(), static_cast<uint8_t>(0x10 + i));
	// stripe 0 of chunk 1 incomplete (2 of 3), then two complete stripes: flush() moves the complete ones to the front
	EXPECT(batcher.addBlock(1, 0, blk[0].data()) && batcher.addBlock(1, 1, blk[1].data()));
	for (int j = 0; j < 3; ++j) EXPECT(batcher.addBlock(2, j, blk[3 + j].data()));
	for (int j = 0; j < 3; ++j) EXPECT(batcher.addBlock(3, j, blk[6 + j].data()));
	lzgpu_test_fail_next_encode = 1;
	bool threw = false;
	try {
		batcher.flush(0, [](const lzgpu::PartBlock &) {});
	} catch (const std::runtime_error &) { threw = true; }
	EXPECT(threw);
	// the missing block of chunk 1 arrives: it must land in chunk 1's stripe, not in the slot that stripe used to occupy
	EXPECT(batcher.addBlock(1, 2, blk[2].data()));
	std::map<std::pair<uint64_t, int>, uint8_t> first_byte;
	EXPECT(batcher.flush(0, [&](const lzgpu::PartBlock &pb) { if (pb.part < 3) first_byte[{pb.chunk_id, pb.part}] = pb.data[0]; }) == 3);
	for (int j = 0; j < 3; ++j) {
		EXPECT((first_byte[{1, j}] == 0x10 + j));
		EXPECT((first_byte[{2, j}] == 0x13 + j));
		EXPECT((first_byte[{3, j}] == 0x16 + j));
	}
	std::printf("failed flush keeps the index: ok\n");
}
#endif

int main() {
	if (!lzgpu_default_ctx()) {
		std::fprintf(stderr, "no GPU context: %s\n", lzgpu_last_error());
		return 2;
	}
	run("ec(8,2)", 1);
	run("ec(3,2)", 2);
	run("xor3", 3);
	run("ec(5,3)", 4);
	test_compute_parity_block("ec(8,2)");
	test_compute_parity_block("xor3");
	test_compute_parity_block("ec(3,2)");
	test_sub_block_stripes("ec(8,2)");
	test_sub_block_stripes("xor3");
	test_sub_block_stripes("ec(5,3)");
#ifdef LZ_TEST_CPU_BACKEND
	test_failed_flush_keeps_the_index();
#endif
	if (failures) {
		std::fprintf(stderr, "%d failure(s)\n", failures);
		return 1;
	}
	std::printf("stripe batcher: all tests passed\n");
	return 0;
}
