// Internal context of libharcgpu: everything behind the opaque harcgpu_ctx of include/harcgpu.h.
#pragma once
#include "common.cuh"
#include <map>
#include <stdlib.h>
#include <vector>

struct DevBuf {
	void *p = nullptr;
	size_t bytes = 0;
};

struct SetOut { // one read_*.txt.<k> file set, device resident
	u8 *seq = nullptr; u64 seq_bytes = 0; char seq_tail[4]; u32 seq_ntail = 0;
	u8 *pos = nullptr; u64 pos_bytes = 0;
	char *noise = nullptr; u64 noise_bytes = 0;
	u8 *noisepos = nullptr; u64 noisepos_bytes = 0;
	u8 *rev = nullptr; u64 rev_bytes = 0; char rev_tail[8]; u32 rev_ntail = 0;
};

struct harcgpu_ctx {
	int device = 0;
	harcgpu_params p;
	cudaStream_t st = nullptr;
	int L = 0, NW = 0, NW3 = 0;
	std::map<std::string, double> ms;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;

	// ---- stage I
	u32 n = 0;
	u64 *reads = nullptr;   // [n][NW] 2-bit packed (reorder.cpp:188-195)
	DictDev d1[2];
	bool dicts_built = false;
	u32 *claim = nullptr;   // bitmap, 1 = still unclaimed (remainingreads, reorder.cpp:449)
	unsigned long long *tailc = nullptr; // cursor cache of the big bins (walk.cu: advance)
	u32 *bloom1 = nullptr;  // blocked Bloom filter over the keys of both stage I dictionaries, sized to stay in L2 (job.cu: job_bloom_pos)
	u32 bloom1_words = 0;
	long long *gpos = nullptr;
	u64 *counters = nullptr; // 8 x u64
	u32 walkers_used = 0;
	// ---- one job on several GPUs (job.cu, harcgpu_job_*).  Every GPU owns one allocation, its ARENA, which the other GPUs
	// map (CUDA IPC between processes; plain pointers between contexts of one process): a header (barrier flags, the
	// counts of the key exchange), its range of the claim bitmap, a full replica of the packed reads (every GPU packs its
	// slice of the input and stores the result into all replicas), per dictionary the receive buffers of the key
	// exchange, its shard of the key table and of the id lists, and its copy of the Bloom filter over all shards.
	int shard_rank = 0, shard_world = 1;
	u32 shard_n = 0, seg_per = 0;         // reads of the whole job; reads per bitmap range
	u32 job_base = 0, job_nloc = 0;       // slice of the reads this GPU uploads and extracts keys for: [base, base + nloc)
	char *arena[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }; // arena[shard_rank] is local
	u32 *seg[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };    // bitmap range of GPU r (inside arena r)
	bool seg_opened[8] = { false, false, false, false, false, false, false, false };             // mapped through CUDA IPC
	size_t arena_bytes = 0, arena_bitmap_off = 0, arena_reads_off = 0, arena_bloom_off = 0;
	size_t arena_rk_off[2] = { 0, 0 }, arena_ri_off[2] = { 0, 0 }; // received (key, id) pairs of the exchange
	size_t arena_slots_off[2] = { 0, 0 }, arena_ids_off[2] = { 0, 0 };
	bool dicts_sharded = false;
	u32 shard_cap = 0;                    // nominal slots per dictionary shard (power of two)
	u64 shard_nslots = 0;                 // slots reserved per shard: nominal + room for the spill at the end
	u32 recv_cap = 0;                     // pairs a shard has room for
	u32 bloom_seg_words = 0;              // words per Bloom segment (power of two); one segment per shard
	int job_bloom = 1;                    // probe the Bloom filter before a remote table (HARCGPU_JOB_BLOOM=0 turns it off)
	u32 job_epoch = 0;                    // barriers passed so far (the same on every rank)
	bool shard_ready = false, job_reads_loaded = false;
	cudaStream_t st_bcast = nullptr;      // side stream of the broadcast of this GPU's packed slice
	cudaEvent_t ev_packed = nullptr, ev_bcast = nullptr;
	bool bcast_pending = false, bcast_needed = false, bcast_dma = false;
	cudaStream_t st_peer[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr }; // one copy stream per peer
	cudaEvent_t ev_peer[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	int (*pool_exchange)(void *user, void *d_best, uint64_t count) = nullptr;
	void *pool_exchange_user = nullptr;
	int (*job_barrier_hook)(void *user) = nullptr; // ranks that share one GPU (tests): host barrier instead of the barrier kernel
	void *job_barrier_user = nullptr;
	// fused ingest (ingest.cu): reads with N as ASCII lines + their record numbers (input_N.dna, read_order_N.bin)
	char *ing_N = nullptr;
	u32 *ing_orderN = nullptr;
	u32 ing_nN = 0;
	// finalized stage I streams (device)
	bool reordered = false;
	u32 n_matched = 0, n_single = 0, n_unmatched = 0;
	u32 *order = nullptr, *order_s = nullptr;
	u8 *rev = nullptr, *flag = nullptr, *pos = nullptr;

	// ---- stage II inputs
	bool stream_set = false;
	u32 m = 0;              // reads in the reordered stream
	u64 *sreads = nullptr;  // [m][NW] stream reads, already reverse-complemented where flagged
	u32 *s_order = nullptr;
	u8 *s_rev = nullptr, *s_flag = nullptr, *s_pos = nullptr;
	// input_N.dna uploaded ahead of time on a copy stream (harcgpu_stage_nreads)
	cudaStream_t st_copy = nullptr;
	cudaEvent_t ev_staged = nullptr, ev_order = nullptr;
	char *staged_N = nullptr;
	const char *staged_host = nullptr;
	u32 staged_n = 0;
	bool pool_set = false;
	u32 n_s = 0, n_N = 0;   // pool = n_s singletons ++ n_N reads with N
	u64 *pool = nullptr;    // [n_s+n_N][NW] 2-bit codes with N stored as 0
	u64 *poolN = nullptr;   // [n_s+n_N][NW] bit 2i set where base i is N; 3-bit code of encoder.cpp:731-745 = 2*code2 + nflag
	u32 *pool_order = nullptr; // order_s (encoder.cpp:865-870)
	DictDev d2[2];
	u32 *bloom2 = nullptr;  // blocked Bloom filter over the keys of d2[0] and d2[1] (stage2.cu: bloom_pos)
	u32 bloom2_mask = 0;
	// ---- stage II outputs
	bool encoded = false;
	std::vector<SetOut> sets;
	std::vector<void *> s2_keep; // global stage II streams the per-set views point into
	harcgpu_encode_sizes esz;
	u32 *o_order = nullptr, *o_order_N = nullptr;
	u8 *o_single = nullptr; char single_tail[4];
	char *o_inputN = nullptr;

	// Device memory: a caching allocator private to the context.  Every kernel and copy of a context runs on its one
	// stream, so a block can be handed out again as soon as it is released (reuse is stream-ordered by construction).
	// A request takes the smallest cached block that is large enough but at most 30 % larger; otherwise it goes to
	// cudaMalloc.  A pass repeats the sizes of the pass before it, so after the first pass no allocation
	// reaches the driver.  (cudaMallocAsync was measured here to spend hundreds of ms per pass remapping its pool when
	// block sizes vary between calls.)
	struct Block { void *p; size_t bytes; };
	std::vector<Block> live, cached;
	size_t cached_bytes = 0, live_bytes = 0, peak_bytes = 0;
	int alloc_log = -1;
	unsigned long long n_cuda_malloc = 0; // allocations that reached the driver (harcgpu_last_ms(ctx, "cudaMalloc_calls"))
	// Size classes: eight per power of two (512-byte floor), so that a request whose size moves a little from pass to pass
	// (the walk is not deterministic) lands in the class of the pass before, and a block serves requests of its own class
	// or up to two classes below (<= 1.3x) -- small requests must not take the blocks that larger ones will ask for a
	// moment later (an additive slack did that, and in a job on several GPUs, where every cudaMalloc also maps the block
	// for the peers, each miss cost ~20 ms).
	static size_t round_bytes(size_t b)
	{
		if (b <= 512) return 512;
		size_t step = 64;
		while ((step << 4) <= b) step <<= 1; // step = 2^(floor(log2 b) - 3)
		return (b + step - 1) / step * step;
	}
	void trim()
	{
		for (auto &b : cached) cudaFree(b.p);
		cached.clear();
		cached_bytes = 0;
	}
	template <typename T> int alloc(T **out, size_t count)
	{
		size_t bytes = round_bytes((count ? count : 1) * sizeof(T));
		size_t best = (size_t)-1;
		for (size_t i = 0; i < cached.size(); i++)
			if (cached[i].bytes >= bytes && cached[i].bytes <= bytes + bytes / 4 + bytes / 20 &&
			    (best == (size_t)-1 || cached[i].bytes < cached[best].bytes))
				best = i;
		Block b;
		if (best != (size_t)-1) {
			b = cached[best];
			cached[best] = cached.back();
			cached.pop_back();
			cached_bytes -= b.bytes;
		} else {
			void *q = nullptr;
			n_cuda_malloc++;
			if (alloc_log < 0) { const char *ev = getenv("HARCGPU_ALLOC_LOG"); alloc_log = ev && atoi(ev) ? 1 : 0; }
			if (alloc_log) fprintf(stderr, "[harcgpu dev %d] cudaMalloc #%llu of %zu bytes (%zu asked); %zu blocks / %zu MB cached\n", device, n_cuda_malloc,
			                       bytes, (count ? count : 1) * sizeof(T), cached.size(), cached_bytes >> 20);
			cudaError_t e = cudaMalloc(&q, bytes);
			if (e != cudaSuccess) { // give the cached blocks back to the driver and try once more
				cudaGetLastError();
				cudaStreamSynchronize(st);
				trim();
				e = cudaMalloc(&q, bytes);
			}
			if (e != cudaSuccess) {
				cudaGetLastError();
				harcgpu_set_error("cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
				*out = nullptr;
				return -1;
			}
			b.p = q; b.bytes = bytes;
		}
		live.push_back(b);
		live_bytes += b.bytes;
		if (live_bytes > peak_bytes) peak_bytes = live_bytes;
		*out = (T *)b.p;
		return 0;
	}
	void release(void *q)
	{
		if (!q) return;
		for (size_t i = 0; i < live.size(); i++)
			if (live[i].p == q) {
				cached.push_back(live[i]);
				cached_bytes += live[i].bytes;
				live_bytes -= live[i].bytes;
				live[i] = live.back();
				live.pop_back();
				return;
			}
	}
	// tuning aid (HARCGPU_LAPS=1): device time between named points inside a phase, as "lap:<name>" of harcgpu_last_ms
	int laps = -1;
	cudaEvent_t evl0 = nullptr, evl1 = nullptr;
	void lap(const char *name, bool add = false)
	{
		if (laps < 0) { const char *e = getenv("HARCGPU_LAPS"); laps = e && atoi(e) ? 1 : 0; }
		if (!laps) return;
		if (!evl0) { cudaEventCreate(&evl0); cudaEventCreate(&evl1); cudaEventRecord(evl0, st); }
		cudaEventRecord(evl1, st);
		cudaEventSynchronize(evl1);
		float f = 0;
		cudaEventElapsedTime(&f, evl0, evl1);
		if (name) { if (add) ms[std::string("lap:") + name] += f; else ms[std::string("lap:") + name] = f; }
		std::swap(evl0, evl1);
	}
	void tic() { cudaEventRecord(ev0, st); }
	void toc(const char *phase)
	{
		cudaEventRecord(ev1, st);
		cudaEventSynchronize(ev1);
		float f = 0;
		cudaEventElapsedTime(&f, ev0, ev1);
		ms[phase] = f;
	}
};

// sort.cu
int radix_sort_pairs(harcgpu_ctx *c, u64 **keys, u64 **keys_alt, u32 **vals, u32 **vals_alt, size_t n, int begin_bit, int end_bit);
int radix_sort_mixed(harcgpu_ctx *c, u64 **keys, u64 **keys_alt, u32 **vals, u32 **vals_alt, size_t n, bool force_full = false);
// ingest.cu
int ing_ingest(harcgpu_ctx *c, const char *d_fastq, u64 nbytes, u64 *total_reads, u32 *n_clean, u32 *n_N);
int ing_unpack_clean(harcgpu_ctx *c, char *d_out);
// walk.cu
int s1_init_claim(harcgpu_ctx *c, u32 *claim, u32 n);
// job.cu
int job_bloom_insert(harcgpu_ctx *c, const u64 *mixed_keys, u32 nk, int l, int world, u32 seg_words, u32 *bloom);
void job_close(harcgpu_ctx *c);
int job_barrier(harcgpu_ctx *c);
// stage1.cu
int s1_pack_reads(harcgpu_ctx *c, const void *d_ascii, u32 n);
// pack n lines into out[n][NW]
int s1_pack_reads_to(harcgpu_ctx *c, const void *d_ascii, u32 n, u64 *out);
int s1_keys(harcgpu_ctx *c, const u64 *reads, u32 n, int words, int bitpos, int nbits, u32 id0, u64 *keys, u32 *ids);
int s1_packN(harcgpu_ctx *c, const void *d_ascii, u32 n, u64 *out2, u64 *outN);
struct DictShard { // where a dictionary shard of one job on several GPUs is built (inside the arena of ctx.h)
	int rank, world;
	// the (mixed key, id) pairs of the shard, already exchanged (ids ascending among equal keys); null: every key of
	// `reads` is extracted here and the pairs of other shards are dropped (the reads are replicated anyway)
	const u64 *pair_keys = nullptr;
	const u32 *pair_ids = nullptr;
	u32 npairs = 0;
	ulonglong2 *slots;
	u32 cap;   // nominal slots, power of two (fixes the home buckets)
	u64 nslots; // slots the arena has room for: cap + spill
	u32 *ids;  // room for every read id
};
int build_dict(harcgpu_ctx *c, DictDev &d, const u64 *reads, const u64 *readsN, u32 n, int words, int ds, int de, int bits,
               const DictShard *shard = nullptr);
void free_dict(harcgpu_ctx *c, DictDev &d);
int s1_reorder(harcgpu_ctx *c);
int s1_unpack_reads(harcgpu_ctx *c, const u64 *reads, const u32 *order, const u8 *rev, u32 cnt, char *d_out);
// stage2.cu
int s2_set_stream_from_stage1(harcgpu_ctx *c);
int s2_set_stream_host(harcgpu_ctx *c, const char *dna, const char *flag, const u8 *pos, const u32 *order, const char *rev, u32 n);
int s2_load_pool(harcgpu_ctx *c, const char *s_ascii, const u32 *order_s, u32 n_s, const char *N_ascii, u32 n_N);
int s2_load_pool_dev(harcgpu_ctx *c, const void *d_N_ascii, u32 n_N);
int s2_load_pool_ids(harcgpu_ctx *c, const u32 *ids, u32 n_s, const char *N_ascii, u32 n_N);
int s2_encode(harcgpu_ctx *c);
int s2_pack_order(harcgpu_ctx *c, void *h_packed, u32 *h_tail, u64 *packed_bytes, u32 *ntail);
