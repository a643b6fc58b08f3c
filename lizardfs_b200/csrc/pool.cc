// pool.cc — lzgpu_pool: ONE process driving several GPUs (include/lzgpu.h "device pool").
//
// The reference's callers are multi-threaded inside one process — ten write workers in the mount
// (src/mount/lizard_client.h:77, src/mount/writedata.cc:645), the chunkserver's background job pool — so the multi-GPU
// form of this engine that a drop-in needs is a pool of per-device contexts inside that process, not one process per GPU.
// Chunks are independent: a batch is cut into one contiguous run of chunks per device ("batch b -> device b", the static
// round-robin of chunk batches of BASELINE.json's north_star with a batch = ceil(n / devices) chunks); each run goes through
// the device's own 3-slot H2D | kernel | D2H pipeline on a worker thread that stays bound to that device.  No collective,
// no peer copy: the only cross-device state is the error code.
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "engine_internal.h"
#include "lzgpu.h"

namespace {

struct Worker {
	lzgpu_ctx *ctx = nullptr;
	std::thread th;
	std::mutex mu;
	std::condition_variable cv;
	std::deque<std::function<void()>> q;
	bool stop = false;

	void loop() {
		cudaSetDevice(ctx->device);
		for (;;) {
			std::function<void()> job;
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&] { return stop || !q.empty(); });
				if (q.empty()) return;
				job = std::move(q.front());
				q.pop_front();
			}
			job();
		}
	}
	void post(std::function<void()> f) {
		{
			std::lock_guard<std::mutex> lk(mu);
			q.push_back(std::move(f));
		}
		cv.notify_one();
	}
};

// completion latch of one pool call
struct Latch {
	std::mutex mu;
	std::condition_variable cv;
	int pending = 0;
	void done() {
		std::lock_guard<std::mutex> lk(mu);
		if (--pending == 0) cv.notify_all();
	}
	void wait() {
		std::unique_lock<std::mutex> lk(mu);
		cv.wait(lk, [&] { return pending == 0; });
	}
};

}  // namespace

struct lzgpu_pool {
	std::vector<Worker *> workers;
};

extern "C" void lzgpu_pool_destroy(lzgpu_pool *pool) {
	if (!pool) return;
	for (Worker *w : pool->workers) {
		if (w->th.joinable()) {
			{
				std::lock_guard<std::mutex> lk(w->mu);
				w->stop = true;
			}
			w->cv.notify_one();
			w->th.join();
		}
		lzgpu_ctx_destroy(w->ctx);
		delete w;
	}
	delete pool;
}

extern "C" int lzgpu_pool_create_list(const int *devices, int n_devices, lzgpu_pool **out) {
	if (!out || !devices || n_devices < 1 || n_devices > 64) return LZGPU_ERR_ARG;
	*out = nullptr;
	auto *pool = new lzgpu_pool();
	for (int i = 0; i < n_devices; ++i) {
		auto *w = new Worker();
		int rc = lzgpu_ctx_create(devices[i], &w->ctx);
		if (rc != LZGPU_OK) {
			delete w;
			lzgpu_pool_destroy(pool);
			return rc;
		}
		pool->workers.push_back(w);
		w->th = std::thread([w] { w->loop(); });
	}
	*out = pool;
	return LZGPU_OK;
}

extern "C" int lzgpu_pool_create(uint64_t device_mask, lzgpu_pool **out) {
	if (!out) return LZGPU_ERR_ARG;
	*out = nullptr;
	const int n = lzgpu_device_count();
	if (n <= 0) {
		lz_set_error("no CUDA device visible: liblzgpu has no CPU fallback");
		return LZGPU_ERR_NO_DEVICE;
	}
	std::vector<int> devs;
	for (int d = 0; d < n && d < 64; ++d)
		if (device_mask == 0 || (device_mask >> d) & 1) devs.push_back(d);
	if (devs.empty()) {
		lz_set_error("device mask 0x%llx selects none of the %d visible devices", static_cast<unsigned long long>(device_mask), n);
		return LZGPU_ERR_ARG;
	}
	return lzgpu_pool_create_list(devs.data(), static_cast<int>(devs.size()), out);
}

extern "C" int lzgpu_pool_size(const lzgpu_pool *pool) { return pool ? static_cast<int>(pool->workers.size()) : 0; }

extern "C" lzgpu_ctx *lzgpu_pool_ctx(lzgpu_pool *pool, int i) {
	if (!pool || i < 0 || i >= static_cast<int>(pool->workers.size())) return nullptr;
	return pool->workers[i]->ctx;
}

// which run of chunks device slot i of G takes: [first, first + count)
extern "C" void lzgpu_pool_share(uint32_t n_chunks, int n_devices, int i, uint32_t *first, uint32_t *count) {
	const uint32_t per = n_devices > 0 ? (n_chunks + n_devices - 1) / n_devices : n_chunks;
	const uint32_t f = std::min<uint64_t>(static_cast<uint64_t>(per) * i, n_chunks);
	if (first) *first = f;
	if (count) *count = std::min<uint32_t>(per, n_chunks - f);
}

// run fn(worker index, first chunk, chunk count) on every device that gets a share; first error by device order wins
static int pool_run(lzgpu_pool *pool, uint32_t n_chunks, const std::function<int(int, uint32_t, uint32_t)> &fn, std::vector<std::string> *errs) {
	const int G = static_cast<int>(pool->workers.size());
	std::vector<int> rcs(G, LZGPU_OK);
	errs->assign(G, std::string());
	Latch latch;
	for (int i = 0; i < G; ++i) {
		uint32_t first, count;
		lzgpu_pool_share(n_chunks, G, i, &first, &count);
		if (count) latch.pending++;
	}
	for (int i = 0; i < G; ++i) {
		uint32_t first, count;
		lzgpu_pool_share(n_chunks, G, i, &first, &count);
		if (!count) continue;
		pool->workers[i]->post([&, i, first, count] {
			rcs[i] = fn(i, first, count);
			if (rcs[i] != LZGPU_OK) (*errs)[i] = lzgpu_last_error();  // thread-local text of the worker
			latch.done();
		});
	}
	latch.wait();
	for (int i = 0; i < G; ++i)
		if (rcs[i] != LZGPU_OK) {
			lz_set_error("device slot %d: %s", i, (*errs)[i].c_str());
			return rcs[i];
		}
	return LZGPU_OK;
}

extern "C" int lzgpu_pool_encode_chunks(lzgpu_pool *pool, const lzgpu_goal *goal, uint32_t n_chunks, uint32_t chunk_len, const uint8_t *data,
                                         size_t chunk_stride, uint8_t *parity, size_t parity_stride, uint32_t *crc, size_t crc_stride) {
	if (!pool || !goal || !data || !parity || !crc) return LZGPU_ERR_ARG;
	if (n_chunks == 0) return LZGPU_OK;
	std::vector<std::string> errs;
	return pool_run(pool, n_chunks, [&](int i, uint32_t first, uint32_t count) {
		return lzgpu_encode_chunks(pool->workers[i]->ctx, goal, count, chunk_len, data + static_cast<size_t>(first) * chunk_stride, chunk_stride,
		                           parity + static_cast<size_t>(first) * parity_stride, parity_stride, crc + static_cast<size_t>(first) * crc_stride,
		                           crc_stride);
	}, &errs);
}

extern "C" int lzgpu_pool_recover_chunks(lzgpu_pool *pool, const lzgpu_goal *goal, uint32_t n_chunks, uint32_t nb, const uint8_t *const *parts,
                                          size_t part_stride, const uint32_t *const *part_crc, const uint8_t *want, uint8_t *const *out,
                                          uint8_t *chunk_out, size_t chunk_out_stride, int64_t *bad) {
	if (!pool || !goal || !parts || !want) return LZGPU_ERR_ARG;
	if (goal->k < 1 || goal->k > LZGPU_MAX_DATA || goal->m < 0 || goal->m > LZGPU_MAX_PARITY) return LZGPU_ERR_ARG;
	if (n_chunks == 0) return LZGPU_OK;
	const int G = static_cast<int>(pool->workers.size()), n = goal->k + goal->m;
	const uint32_t pb = (nb + goal->k - 1) / goal->k;
	std::vector<int64_t> bads(static_cast<size_t>(G) * 3, -1);
	std::vector<std::string> errs;
	int rc = pool_run(pool, n_chunks, [&](int i, uint32_t first, uint32_t count) {
		std::vector<const uint8_t *> p(n, nullptr);
		std::vector<const uint32_t *> pc(n, nullptr);
		std::vector<uint8_t *> o(n, nullptr);
		for (int j = 0; j < n; ++j) {
			if (parts[j]) p[j] = parts[j] + static_cast<size_t>(first) * part_stride;
			if (part_crc && part_crc[j]) pc[j] = part_crc[j] + static_cast<size_t>(first) * pb;
			if (out && out[j]) o[j] = out[j] + static_cast<size_t>(first) * part_stride;
		}
		int r = lzgpu_recover_chunks(pool->workers[i]->ctx, goal, count, nb, p.data(), part_stride, part_crc ? pc.data() : nullptr, want,
		                             out ? o.data() : nullptr, chunk_out ? chunk_out + static_cast<size_t>(first) * chunk_out_stride : nullptr,
		                             chunk_out_stride, &bads[static_cast<size_t>(i) * 3]);
		if (r == LZGPU_ERR_CRC && bads[static_cast<size_t>(i) * 3] >= 0) bads[static_cast<size_t>(i) * 3] += first;
		return r;
	}, &errs);
	if (rc == LZGPU_ERR_CRC && bad) {
		// lowest chunk index over the devices that reported a mismatch (shares are ascending runs of chunks)
		for (int i = 0; i < G; ++i)
			if (bads[static_cast<size_t>(i) * 3] >= 0) {
				std::memcpy(bad, &bads[static_cast<size_t>(i) * 3], 3 * sizeof(int64_t));
				break;
			}
	}
	return rc;
}

// Slice-type conversion (replication) over the devices of the pool: same arguments as lzgpu_convert_chunks, chunks dealt in runs
extern "C" int lzgpu_pool_convert_chunks(lzgpu_pool *pool, const lzgpu_goal *src, const lzgpu_goal *dst, uint32_t n_chunks, uint32_t nb,
                                          const uint8_t *const *parts, size_t part_stride, const uint32_t *const *part_crc, const uint8_t *want,
                                          uint8_t *const *out, size_t out_stride, uint32_t *const *out_crc, int64_t *bad) {
	if (!pool || !src || !dst || !parts || !want || !out) return LZGPU_ERR_ARG;
	if (src->k < 1 || src->k > LZGPU_MAX_DATA || src->m < 0 || src->m > LZGPU_MAX_PARITY || dst->k < 1 || dst->k > LZGPU_MAX_DATA || dst->m < 0 ||
	    dst->m > LZGPU_MAX_PARITY)
		return LZGPU_ERR_ARG;
	if (n_chunks == 0) return LZGPU_OK;
	const int G = static_cast<int>(pool->workers.size()), ns = src->k + src->m, nd = dst->k + dst->m;
	// blocks per chunk of a source / destination part (a standard slice has the single part 0 = the chunk itself)
	const uint32_t pbs = src->kind == LZGPU_KIND_STD ? nb : (nb + src->k - 1) / src->k, pbd = dst->kind == LZGPU_KIND_STD ? nb : (nb + dst->k - 1) / dst->k;
	std::vector<int64_t> bads(static_cast<size_t>(G) * 3, -1);
	std::vector<std::string> errs;
	int rc = pool_run(pool, n_chunks, [&](int i, uint32_t first, uint32_t count) {
		std::vector<const uint8_t *> p(ns, nullptr);
		std::vector<const uint32_t *> pc(ns, nullptr);
		std::vector<uint8_t *> o(nd, nullptr);
		std::vector<uint32_t *> oc(nd, nullptr);
		for (int j = 0; j < ns; ++j) {
			if (parts[j]) p[j] = parts[j] + static_cast<size_t>(first) * part_stride;
			if (part_crc && part_crc[j]) pc[j] = part_crc[j] + static_cast<size_t>(first) * pbs;
		}
		for (int j = 0; j < nd; ++j) {
			if (out[j]) o[j] = out[j] + static_cast<size_t>(first) * out_stride;
			if (out_crc && out_crc[j]) oc[j] = out_crc[j] + static_cast<size_t>(first) * pbd;
		}
		int r = lzgpu_convert_chunks(pool->workers[i]->ctx, src, dst, count, nb, p.data(), part_stride, part_crc ? pc.data() : nullptr, want, o.data(),
		                             out_stride, out_crc ? oc.data() : nullptr, &bads[static_cast<size_t>(i) * 3]);
		if (r == LZGPU_ERR_CRC && bads[static_cast<size_t>(i) * 3] >= 0) bads[static_cast<size_t>(i) * 3] += first;
		return r;
	}, &errs);
	if (rc == LZGPU_ERR_CRC && bad) {
		for (int i = 0; i < G; ++i)
			if (bads[static_cast<size_t>(i) * 3] >= 0) {
				std::memcpy(bad, &bads[static_cast<size_t>(i) * 3], 3 * sizeof(int64_t));
				break;
			}
	}
	return rc;
}

extern "C" int lzgpu_pool_crc_blocks(lzgpu_pool *pool, const uint8_t *data, size_t n_blocks, uint32_t block_len, size_t block_stride,
                                      uint32_t *crc_out) {
	if (!pool || !data || !crc_out) return LZGPU_ERR_ARG;
	if (n_blocks == 0) return LZGPU_OK;
	if (n_blocks > 0xffffffffull) return LZGPU_ERR_ARG;
	std::vector<std::string> errs;
	return pool_run(pool, static_cast<uint32_t>(n_blocks), [&](int i, uint32_t first, uint32_t count) {
		return lzgpu_crc_blocks(pool->workers[i]->ctx, data + static_cast<size_t>(first) * block_stride, count, block_len, block_stride, crc_out + first);
	}, &errs);
}

extern "C" void lzgpu_pool_get_stats(lzgpu_pool *pool, lzgpu_stats *out) {
	if (!pool || !out) return;
	std::memset(out, 0, sizeof(*out));
	double bytes_total = 0.0;
	for (Worker *w : pool->workers) {
		lzgpu_stats s;
		lzgpu_get_stats(w->ctx, &s);
		out->kernel_launches += s.kernel_launches;
		out->bytes_h2d += s.bytes_h2d;
		out->bytes_d2h += s.bytes_d2h;
		out->chunks_encoded += s.chunks_encoded;
		out->chunks_recovered += s.chunks_recovered;
		out->blocks_crc += s.blocks_crc;
		out->batches_timed += s.batches_timed;
		out->batch_ms_total += s.batch_ms_total;
		bytes_total += s.batch_gbps_mean * s.batch_ms_total * 1e6;
		if (s.batch_ms_last > 0.0) {
			out->batch_ms_last = s.batch_ms_last;
			out->batch_bytes_last = s.batch_bytes_last;
			out->batch_gbps_last = s.batch_gbps_last;
		}
	}
	// devices run concurrently: the mean rate of the pool is per-device mean x devices only if they overlap fully, so report
	// bytes / summed device time (a per-device mean) and leave aggregation over wall time to the caller
	out->batch_gbps_mean = out->batch_ms_total > 0.0 ? bytes_total / (out->batch_ms_total * 1e6) : 0.0;
}
