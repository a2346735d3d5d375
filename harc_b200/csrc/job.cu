// One compression job on several GPUs of one box (include/harcgpu.h, harcgpu_job_*): one process (or, for tests, one
// context) and one ARENA per GPU.  Not in the reference, which is one process; what is split is what its threads
// split (reorder.cpp:284-302 key extraction, 476-497 walker starts, encoder.cpp:169-180 contig ranges).
//
// What crosses NVLink, all of it by kernels of this library over peer memory (no library collective):
//   * packed reads      -- every GPU packs its slice of the input and job_bcast_kernel stores the packed slice into the
//                          replica of every other GPU, on a side stream next to the shard build (which only needs the
//                          pairs): the all-gather of the packed reads, mostly hidden behind the build;
//   * (key, id) pairs   -- every GPU extracts the dictionary keys of its slice, cuts them by owner (one stable radix pass
//                          over the shard bits) and job_push_kernel stores every range straight into the owner's receive
//                          buffer at the offset that follows from the 8 x 8 count matrix (every GPU broadcasts its row):
//                          the all-to-all of the north star, source ranks in rank order so ids stay ascending in a bin;
//   * Bloom segments    -- every GPU builds the filter segment of its shard and copies the other segments from their owners,
//                          so that a walker asks a remote table only for keys that are (almost certainly) there;
//   * probes and claims -- from inside the walk kernel (walk.cu): 32-byte loads from the owner's table, atomicAnd on the
//                          owner's range of the claim bitmap;
//   * barriers          -- job_barrier_kernel: every GPU stores its epoch into the other arenas and spins on its own.
// The two small exchanges of stage II (singleton ids, pool priorities) go through the caller (NCCL in this repo).
#include "ctx.h"
#include <algorithm>
#include <string.h>
#include <stdlib.h>

namespace {
struct JobHdr { // at offset 0 of every arena
	u32 flag[8];       // flag[r] = last barrier epoch rank r has reached (written by rank r)
	u32 error;         // 1: a barrier timed out, 2: a shard overflowed
	u32 pad[7];
	u32 cnt[2][8][8];  // cnt[l][src][dst]: pairs of dictionary l that rank src sends to rank dst (every rank writes its row everywhere)
	u32 bounds[2][9];  // own arena only: where the range for each owner starts in this rank's partitioned pairs
};
constexpr size_t JOB_HDR_BYTES = 1024;
static_assert(sizeof(JobHdr) <= JOB_HDR_BYTES, "job header");
struct HdrPtrs { JobHdr *p[8]; };
struct PushDst { u64 *k[8]; u32 *v[8]; };
struct PullSrc { const uint4 *p[8]; };

__device__ __forceinline__ unsigned long long gtimer()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// Barrier over all GPUs of the job, stream-ordered on each of them: lane r tells rank r "I have reached epoch e" (a store
// into rank r's arena, after a system-wide fence so that everything this GPU wrote before is visible first) and waits
// until rank r has said the same in this GPU's arena.  Epochs only grow, so the flags are never reset.
__global__ void job_barrier_kernel(HdrPtrs H, int me, int world, u32 epoch, unsigned long long timeout_ns)
{
	const int r = threadIdx.x;
	if (r < world) {
		__threadfence_system();
		*((volatile u32 *)&H.p[r]->flag[me]) = epoch;
		volatile u32 *f = &H.p[me]->flag[r];
		const unsigned long long t0 = gtimer();
		while ((int)(*f - epoch) < 0) {
			if (gtimer() - t0 > timeout_ns) { *((volatile u32 *)&H.p[me]->error) = 1u; break; }
			__nanosleep(200);
		}
		__threadfence_system();
	}
}

// keys: this rank's pairs after the stable partition by owner (= top kb bits of the mixed key).  Writes the bounds of the
// owner ranges into the own header and the counts row cnt[l][me][*] into EVERY header.
__global__ void job_counts_kernel(const u64 *__restrict__ keys, u32 n, int kb, int world, int l, int me, HdrPtrs H)
{
	__shared__ u32 b[9];
	const int d = threadIdx.x;
	if (d <= world) {
		u32 lo = 0, hi = n;
		if (d == world) lo = n;
		else if (d > 0) {
			const u64 t = (u64)d << (64 - kb);
			while (lo < hi) {
				const u32 mid = lo + (hi - lo) / 2;
				if (keys[mid] < t) lo = mid + 1; else hi = mid;
			}
		}
		b[d] = lo;
	}
	__syncthreads();
	if (d <= world) H.p[me]->bounds[l][d] = b[d];
	if (d < world) {
		const u32 cnt = b[d + 1] - b[d];
		for (int p = 0; p < world; p++) H.p[p]->cnt[l][me][d] = cnt;
	}
}

// The all-to-all: pair i of this rank goes to its owner d, behind the pairs of the ranks before this one.
__global__ void __launch_bounds__(256) job_push_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ vals, u32 n, int kb, int world,
                                                       int l, int me, const JobHdr *__restrict__ mine, PushDst dst, u32 recv_cap)
{
	__shared__ u32 off[8], bnd[8];
	if ((int)threadIdx.x < world) {
		const int d = threadIdx.x;
		u32 o = 0;
		for (int s = 0; s < me; s++) o += mine->cnt[l][s][d];
		off[d] = o;
		bnd[d] = mine->bounds[l][d];
	}
	__syncthreads();
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 k = keys[i];
	const int d = (int)(k >> (64 - kb));
	const u64 idx = (u64)off[d] + (i - bnd[d]);
	if (idx < recv_cap) { dst.k[d][idx] = k; dst.v[d][idx] = vals[i]; } // (the owner sees the overflow in its own count column)
}

__global__ void __launch_bounds__(256) job_bloom_insert_kernel(const u64 *__restrict__ mixed, u32 nk, int l, int world, u32 seg_words, u32 *bloom)
{
	const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nk) return;
	u32 w, b;
	job_bloom_pos(mixed[i], l, world, seg_words, w, b);
	atomicOr(&bloom[w], b);
}

// segment s of the filter comes from the GPU that built it (blockIdx.y = s)
__global__ void __launch_bounds__(256) job_pull_kernel(PullSrc src, uint4 *__restrict__ dst, u32 seg_vec, int me)
{
	const int s = blockIdx.y;
	if (s == me) return;
	const uint4 *from = src.p[s] + (size_t)s * seg_vec;
	uint4 *to = dst + (size_t)s * seg_vec;
	for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < seg_vec; i += gridDim.x * blockDim.x) to[i] = from[i];
}

// The all-gather of the packed reads: this GPU's slice goes into the replica of every other GPU (plain 16-byte stores into
// NVLink peer memory).  Runs on a side stream next to the dictionary build, which only needs the local slice; the walk
// is the first to need the whole replica.
struct BcastDst { void *p[8]; };
template <typename V> // uint4, or u64 when the slice does not start on a 16-byte boundary (odd words per read and odd base)
__global__ void __launch_bounds__(512) job_bcast_kernel(const V *__restrict__ src, size_t nvec, BcastDst dst, int world, int me)
{
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
		const V v = __ldg(&src[i]);
		for (int r = 0; r < world; r++)
			if (r != me) reinterpret_cast<V *>(dst.p[r])[i] = v;
	}
}

size_t round256(size_t b) { return (b + 255) / 256 * 256; }
JobHdr *hdr(harcgpu_ctx *c, int r) { return (JobHdr *)c->arena[r]; }
HdrPtrs hdrs(harcgpu_ctx *c)
{
	HdrPtrs H;
	for (int r = 0; r < 8; r++) H.p[r] = r < c->shard_world ? hdr(c, r) : nullptr;
	return H;
}
} // namespace

int job_bloom_insert(harcgpu_ctx *c, const u64 *mixed_keys, u32 nk, int l, int world, u32 seg_words, u32 *bloom)
{
	if (!nk) return 0;
	job_bloom_insert_kernel<<<KL + cdiv(nk, 256), 256, 0, c->st>>>(mixed_keys, nk, l, world, seg_words, bloom);
	CK(cudaGetLastError());
	return 0;
}

void job_close(harcgpu_ctx *c)
{
	if (c->st_bcast) {
		cudaStreamSynchronize(c->st_bcast);
		for (int r = 0; r < 8; r++) cudaStreamSynchronize(c->st_peer[r]);
	}
	c->bcast_pending = false;
	if (c->arena[c->shard_rank]) { c->reads = nullptr; c->n = 0; c->dicts_built = false; c->reordered = false; } // the reads lived in the arena
	for (int r = 0; r < 8; r++) {
		if (c->arena[r] && c->seg_opened[r]) cudaIpcCloseMemHandle(c->arena[r]);
		else if (c->arena[r] && r == c->shard_rank) c->release(c->arena[r]);
		c->arena[r] = nullptr; c->seg[r] = nullptr; c->seg_opened[r] = false;
	}
	if (c->dicts_sharded) {
		for (int l = 0; l < 2; l++) free_dict(c, c->d1[l]); // their slots / ids pointed into the arena
	}
	c->shard_world = 1; c->shard_rank = 0; c->shard_n = 0; c->seg_per = 0; c->shard_ready = false; c->job_reads_loaded = false;
	c->dicts_sharded = false; c->shard_cap = 0; c->arena_bytes = 0; c->job_epoch = 0;
}

int job_barrier(harcgpu_ctx *c)
{
	if (c->shard_world < 2) return 0;
	if (c->job_barrier_hook) {
		// ranks that are contexts of one process on ONE GPU: kernels of different streams are not guaranteed to run side by
		// side there (any allocation or memset issued in between serialises them), so the ranks meet on the host instead
		CK(cudaStreamSynchronize(c->st));
		if (c->job_barrier_hook(c->job_barrier_user)) { harcgpu_set_error("the job's host barrier failed"); return -1; }
		return 0;
	}
	unsigned long long timeout_s = 30;
	if (const char *e = getenv("HARCGPU_JOB_TIMEOUT_S")) timeout_s = (unsigned long long)std::max(1, atoi(e));
	c->job_epoch++;
	job_barrier_kernel<<<KL + 1, 32, 0, c->st>>>(hdrs(c), c->shard_rank, c->shard_world, c->job_epoch, timeout_s * 1000000000ull);
	CK(cudaGetLastError());
	return 0;
}

// Starts the broadcast of this GPU's packed slice (behind everything queued on the compute stream so far).  Two routes:
// job_bcast_kernel (default: plain stores into the peers' replicas), or the copy engines (HARCGPU_JOB_BCAST=dma: one
// cudaMemcpyAsync per peer, each on a stream of its own).  Measured on 8 GPUs: the kernel route 61.7 ms per step at
// configs[2] and 255.8 ms at configs[4], the copy engines 71.2 and 310.2 ms -- seven peer copies per GPU do not reach the
// NVLink egress that 74 blocks of storing threads do, and the walk ends up waiting for them.  Neither hides completely:
// the same per-GPU dictionary work takes 45 ms next to a 4 GB broadcast (N = 2) and 84 ms next to a 28 GB one (N = 8).
static int job_bcast_start(harcgpu_ctx *c)
{
	if (!c->bcast_needed) return 0;
	c->bcast_needed = false;
	const u32 n_local = c->job_nloc;
	u64 *mine = (u64 *)(c->arena[c->shard_rank] + c->arena_reads_off) + (size_t)c->job_base * c->NW;
	if (!c->st_bcast) {
		CK(cudaStreamCreateWithFlags(&c->st_bcast, cudaStreamNonBlocking));
		CK(cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&c->ev_bcast, cudaEventDisableTiming));
		for (int r = 0; r < 8; r++) {
			CK(cudaStreamCreateWithFlags(&c->st_peer[r], cudaStreamNonBlocking));
			CK(cudaEventCreateWithFlags(&c->ev_peer[r], cudaEventDisableTiming));
		}
		const char *e = getenv("HARCGPU_JOB_BCAST");
		c->bcast_dma = e && !strcmp(e, "dma");
	}
	CK(cudaEventRecord(c->ev_packed, c->st));
	const size_t words = (size_t)n_local * c->NW;
	if (c->bcast_dma) {
		for (int r = 0; r < c->shard_world; r++) {
			if (r == c->shard_rank) continue;
			void *dst = (void *)((u64 *)(c->arena[r] + c->arena_reads_off) + (size_t)c->job_base * c->NW);
			CK(cudaStreamWaitEvent(c->st_peer[r], c->ev_packed, 0));
			CK(cudaMemcpyAsync(dst, mine, words * 8, cudaMemcpyDefault, c->st_peer[r]));
			CK(cudaEventRecord(c->ev_peer[r], c->st_peer[r]));
		}
	} else {
		CK(cudaStreamWaitEvent(c->st_bcast, c->ev_packed, 0));
		BcastDst dst;
		for (int r = 0; r < 8; r++)
			dst.p[r] = r < c->shard_world ? (void *)((u64 *)(c->arena[r] + c->arena_reads_off) + (size_t)c->job_base * c->NW) : nullptr;
		if (((uintptr_t)mine & 15) == 0 && words % 2 == 0)
			job_bcast_kernel<uint4><<<KL + 74, 512, 0, c->st_bcast>>>((const uint4 *)mine, words / 2, dst, c->shard_world, c->shard_rank);
		else
			job_bcast_kernel<u64><<<KL + 74, 512, 0, c->st_bcast>>>((const u64 *)mine, words, dst, c->shard_world, c->shard_rank);
		CK(cudaGetLastError());
		CK(cudaEventRecord(c->ev_bcast, c->st_bcast));
	}
	c->bcast_pending = true;
	return 0;
}

// the compute stream goes on only after this GPU's broadcast of its slice is complete
static int job_bcast_join(harcgpu_ctx *c)
{
	if (job_bcast_start(c)) return -1; // (not started yet: a caller that skipped the dictionary build)
	if (c->bcast_pending) {
		if (c->bcast_dma) {
			for (int r = 0; r < c->shard_world; r++)
				if (r != c->shard_rank) CK(cudaStreamWaitEvent(c->st, c->ev_peer[r], 0));
		} else CK(cudaStreamWaitEvent(c->st, c->ev_bcast, 0));
		c->bcast_pending = false;
	}
	return 0;
}

// error word of the own header (synchronises the stream)
static int job_check(harcgpu_ctx *c, const char *where)
{
	u32 e = 0;
	CK(cudaMemcpyAsync(&e, &hdr(c, c->shard_rank)->error, 4, cudaMemcpyDeviceToHost, c->st));
	CK(cudaStreamSynchronize(c->st));
	if (e == 1) { harcgpu_set_error("%s: a barrier between the GPUs of the job timed out (is every rank making the same calls?)", where); return -1; }
	if (e) { harcgpu_set_error("%s: job error %u", where, e); return -1; }
	return 0;
}

extern "C" {

int harcgpu_job_init(harcgpu_ctx *c, int rank, int world, uint32_t n_total, uint32_t base, uint32_t n_local, void *ipc_handle_out,
                     void **local_ptr_out)
{
	if (!c || world < 1 || world > 8 || rank < 0 || rank >= world) { harcgpu_set_error("bad job arguments (1..8 GPUs)"); return -1; }
	if ((u64)base + n_local > n_total) { harcgpu_set_error("slice [%u, %u + %u) exceeds the %u reads of the job", base, base, n_local, n_total); return -1; }
	if (c->p.shard_dicts != 0 && world > 1 && (world & (world - 1))) { harcgpu_set_error("sharded dictionaries need 2, 4 or 8 GPUs"); return -1; }
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->st));
	job_close(c);
	// a context that was used on its own before: its reads and what was derived from them go
	c->release(c->reads); c->release(c->claim);
	c->reads = nullptr; c->claim = nullptr;
	for (int l = 0; l < 2; l++) free_dict(c, c->d1[l]);
	c->dicts_built = false; c->reordered = false; c->stream_set = false; c->pool_set = false; c->encoded = false;
	c->shard_rank = rank; c->shard_world = world; c->shard_n = n_total; c->job_base = base; c->job_nloc = n_local;
	c->seg_per = (uint32_t)((((uint64_t)n_total + world - 1) / world + 31) / 32 * 32);
	if (c->seg_per == 0) c->seg_per = 32;
	c->dicts_sharded = c->p.shard_dicts != 0 && world > 1;
	c->job_bloom = 1;
	if (const char *e = getenv("HARCGPU_JOB_BLOOM")) c->job_bloom = atoi(e);
	// arena layout (the same on every rank)
	size_t off = JOB_HDR_BYTES;
	c->arena_bitmap_off = off; off += round256((size_t)c->seg_per / 8);
	c->arena_reads_off = off;  off += round256((size_t)n_total * c->NW * 8);
	if (c->dicts_sharded) {
		const u64 per = ((u64)n_total + world - 1) / world;
		const u64 rcap = per + per / 4 + 4096; // room for a shard 25 % above the mean
		if (rcap > 0x7fffffffull) { harcgpu_set_error("dictionary shard too large"); return -1; }
		c->recv_cap = (u32)rcap;
		u64 cap = 16;
		while (cap < 2 * per) cap <<= 1; // load factor <= 0.5 for a shard of mean size (<= 0.63 for one 25 % above it)
		c->shard_cap = (u32)cap;
		c->shard_nslots = cap + cap / 8 + 1024;
		u64 bw = 1024;
		while (bw * 32 < 8 * 2 * rcap) bw <<= 1; // >= 8 bits per key of both dictionaries
		c->bloom_seg_words = (u32)bw;
		for (int l = 0; l < c->p.numdict; l++) {
			c->arena_rk_off[l] = off;    off += round256((size_t)rcap * 8);
			c->arena_ri_off[l] = off;    off += round256((size_t)rcap * 4);
			c->arena_slots_off[l] = off; off += (size_t)c->shard_nslots * sizeof(ulonglong2);
			c->arena_ids_off[l] = off;   off += round256((size_t)rcap * 4);
		}
		c->arena_bloom_off = off; off += (size_t)world * bw * 4;
	}
	c->arena_bytes = off;
	char *arena = nullptr;
	if (c->alloc(&arena, off)) return -1;
	c->arena[rank] = arena;
	c->seg[rank] = (u32 *)(arena + c->arena_bitmap_off);
	CK(cudaMemsetAsync(arena, 0, JOB_HDR_BYTES, c->st));
	c->reads = (u64 *)(arena + c->arena_reads_off);
	c->n = n_total;
	c->job_reads_loaded = false;
	if (c->alloc(&c->claim, ((size_t)n_total + 31) / 32)) return -1; // this GPU's hint bitmap over all reads
	CK(cudaStreamSynchronize(c->st));
	if (ipc_handle_out) {
		static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
		cudaIpcMemHandle_t h;
		CK(cudaIpcGetMemHandle(&h, arena));
		memcpy(ipc_handle_out, &h, sizeof h);
	}
	if (local_ptr_out) *local_ptr_out = arena;
	return 0;
}

int harcgpu_job_connect(harcgpu_ctx *c, const void *handles, void *const *local_ptrs)
{
	if (!c || (!handles && !local_ptrs) || !c->arena[c->shard_rank]) { harcgpu_set_error("harcgpu_job_init first"); return -1; }
	CK(cudaSetDevice(c->device));
	for (int r = 0; r < c->shard_world; r++) {
		if (r == c->shard_rank) continue;
		if (local_ptrs) c->arena[r] = (char *)local_ptrs[r]; // a context of this process
		else {
			cudaIpcMemHandle_t h;
			memcpy(&h, (const char *)handles + 64 * (size_t)r, sizeof h);
			void *p = nullptr;
			CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
			c->arena[r] = (char *)p;
			c->seg_opened[r] = true;
		}
		c->seg[r] = (u32 *)(c->arena[r] + c->arena_bitmap_off);
	}
	return 0;
}

int harcgpu_job_load_reads_device(harcgpu_ctx *c, const void *d_ascii, uint32_t n_local)
{
	if (!c || (!d_ascii && n_local) || !c->arena[c->shard_rank]) { harcgpu_set_error("harcgpu_job_init first"); return -1; }
	if (n_local != c->job_nloc) { harcgpu_set_error("this rank's slice has %u reads, %u were given", c->job_nloc, n_local); return -1; }
	if ((uintptr_t)d_ascii & 15) { harcgpu_set_error("the device buffer must be 16-byte aligned"); return -1; }
	CK(cudaSetDevice(c->device));
	for (int r = 0; r < c->shard_world; r++)
		if (!c->arena[r]) { harcgpu_set_error("harcgpu_job_connect first"); return -1; }
	for (int l = 0; l < 2; l++) if (!c->dicts_sharded) free_dict(c, c->d1[l]);
	c->dicts_built = false; c->reordered = false; c->stream_set = false; c->pool_set = false; c->encoded = false;
	if (job_bcast_join(c) || job_barrier(c)) return -1; // nobody still reads the replicas of the pass before
	c->tic();
	u64 *mine = (u64 *)(c->arena[c->shard_rank] + c->arena_reads_off) + (size_t)c->job_base * c->NW;
	if (s1_pack_reads_to(c, d_ascii, n_local, mine)) return -1;
	c->toc("pack");
	c->bcast_needed = c->shard_world > 1 && n_local;
	c->job_reads_loaded = true;
	return 0;
}

int harcgpu_job_load_reads(harcgpu_ctx *c, const char *ascii, uint32_t n_local)
{
	if (!c || (!ascii && n_local)) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	char *d = nullptr;
	const size_t bytes = (size_t)n_local * (c->L + 1);
	if (c->alloc(&d, bytes + 16)) return -1;
	CK(cudaMemcpyAsync(d, ascii, bytes, cudaMemcpyHostToDevice, c->st));
	const int rc = harcgpu_job_load_reads_device(c, d, n_local);
	CK(cudaStreamSynchronize(c->st));
	c->release(d);
	return rc;
}

// reorder.cpp:277-394 for one job on several GPUs.  Sharded: keys of this rank's slice -> all-to-all by owner -> the shard
// is sorted and built where it lives.  Replicated (params.shard_dicts == 0): every GPU builds both dictionaries over all reads.
int harcgpu_job_build_dicts(harcgpu_ctx *c)
{
	if (!c || !c->arena[c->shard_rank] || !c->job_reads_loaded) { harcgpu_set_error("harcgpu_job_load_reads first"); return -1; }
	CK(cudaSetDevice(c->device));
	cudaStream_t st = c->st;
	const int me = c->shard_rank, world = c->shard_world;
	c->tic();
	if (!c->dicts_sharded) {
		if (job_bcast_start(c) || job_bcast_join(c) || job_barrier(c)) return -1; // every slice has arrived in this GPU's replica
		for (int l = 0; l < c->p.numdict; l++)
			if (build_dict(c, c->d1[l], c->reads, nullptr, c->n, c->NW, c->p.dict_start[l], c->p.dict_end[l], 2, nullptr)) return -1;
		c->toc("dict");
		c->dicts_built = true;
		return job_check(c, "harcgpu_job_build_dicts");
	}
	int kb = 0;
	while ((1 << kb) < world) kb++;
	const u32 nl = c->job_nloc;
	c->lap(nullptr);
	u64 *k[2] = { nullptr, nullptr }, *ka[2] = { nullptr, nullptr };
	u32 *v[2] = { nullptr, nullptr }, *va[2] = { nullptr, nullptr };
	const size_t na = nl ? nl : 1;
	for (int l = 0; l < c->p.numdict; l++) {
		if (c->alloc(&k[l], na) || c->alloc(&ka[l], na) || c->alloc(&v[l], na) || c->alloc(&va[l], na)) return -1;
		const int bitpos = 2 * c->p.dict_start[l], nbits = 2 * (c->p.dict_end[l] - c->p.dict_start[l] + 1);
		if (s1_keys(c, c->reads + (size_t)c->job_base * c->NW, nl, c->NW, bitpos, nbits, c->job_base, k[l], v[l])) return -1;
		// the shard of a key is the top kb bits of its mixed value: one stable pass cuts the pairs into one range per owner
		if (radix_sort_pairs(c, &k[l], &ka[l], &v[l], &va[l], nl, 64 - kb, 64)) return -1;
		job_counts_kernel<<<KL + 1, 32, 0, st>>>(k[l], nl, kb, world, l, me, hdrs(c));
		CK(cudaGetLastError());
	}
	c->lap("dict_keys_partition");
	if (job_barrier(c)) return -1; // the count matrix is complete everywhere (and so are the replicas of the reads)
	c->lap("dict_barrier1");
	for (int l = 0; l < c->p.numdict; l++) {
		PushDst dst;
		for (int r = 0; r < 8; r++) {
			dst.k[r] = r < world ? (u64 *)(c->arena[r] + c->arena_rk_off[l]) : nullptr;
			dst.v[r] = r < world ? (u32 *)(c->arena[r] + c->arena_ri_off[l]) : nullptr;
		}
		if (nl) job_push_kernel<<<KL + cdiv(nl, 256), 256, 0, st>>>(k[l], v[l], nl, kb, world, l, me, hdr(c, me), dst, c->recv_cap);
		CK(cudaGetLastError());
	}
	c->lap("dict_push");
	// the pairs are on their way: now the packed slice follows, on the side stream, under the shard build (the pairs go
	// first because the build waits for them, while the replicas are only needed by the walk)
	if (job_bcast_start(c)) return -1;
	if (job_barrier(c)) return -1; // every pair has arrived
	c->lap("dict_barrier2");
	JobHdr h;
	CK(cudaMemcpyAsync(&h, hdr(c, me), sizeof h, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	for (int l = 0; l < c->p.numdict; l++) { c->release(k[l]); c->release(ka[l]); c->release(v[l]); c->release(va[l]); }
	if (h.error == 1) { harcgpu_set_error("harcgpu_job_build_dicts: a barrier between the GPUs of the job timed out"); return -1; }
	u32 *bloom = (u32 *)(c->arena[me] + c->arena_bloom_off);
	CK(cudaMemsetAsync(bloom + (size_t)me * c->bloom_seg_words, 0, (size_t)c->bloom_seg_words * 4, st));
	for (int l = 0; l < c->p.numdict; l++) {
		u64 nrecv = 0;
		for (int s = 0; s < world; s++) nrecv += h.cnt[l][s][me];
		if (nrecv > c->recv_cap) {
			harcgpu_set_error("dictionary %d: shard %d receives %llu keys, it has room for %u (keys too unevenly spread for a split by hash range)",
			                  l, me, (unsigned long long)nrecv, c->recv_cap);
			return -1;
		}
		DictShard sh;
		sh.rank = me; sh.world = world; sh.cap = c->shard_cap; sh.nslots = c->shard_nslots;
		sh.slots = (ulonglong2 *)(c->arena[me] + c->arena_slots_off[l]);
		sh.ids = (u32 *)(c->arena[me] + c->arena_ids_off[l]);
		sh.pair_keys = (const u64 *)(c->arena[me] + c->arena_rk_off[l]);
		sh.pair_ids = (const u32 *)(c->arena[me] + c->arena_ri_off[l]);
		sh.npairs = (u32)nrecv;
		if (build_dict(c, c->d1[l], c->reads, nullptr, (u32)nrecv, c->NW, c->p.dict_start[l], c->p.dict_end[l], 2, &sh)) return -1;
		if (c->d1[l].numkeys)
			job_bloom_insert_kernel<<<KL + cdiv(c->d1[l].numkeys, 256), 256, 0, st>>>(c->d1[l].keys, c->d1[l].numkeys, l, world, c->bloom_seg_words, bloom);
		CK(cudaGetLastError());
	}
	c->lap("dict_shard_build");
	if (job_barrier(c)) return -1; // every shard and every filter segment is complete
	c->lap("dict_barrier3");
	{
		PullSrc src;
		for (int r = 0; r < 8; r++) src.p[r] = r < world ? (const uint4 *)(c->arena[r] + c->arena_bloom_off) : nullptr;
		const u32 seg_vec = c->bloom_seg_words / 4;
		job_pull_kernel<<<dim3(KL + std::min<u32>(cdiv(seg_vec, 256), 148 * 4), world), 256, 0, st>>>(src, (uint4 *)bloom, seg_vec, me);
		CK(cudaGetLastError());
	}
	c->lap("dict_bloom_pull");
	c->toc("dict");
	c->dicts_built = true;
	return job_check(c, "harcgpu_job_build_dicts");
}

// reorder.cpp:434-703 on this GPU's share of the walkers: its range of the bitmap is armed, all GPUs meet, walk, meet again.
int harcgpu_job_reorder(harcgpu_ctx *c)
{
	if (!c || !c->arena[c->shard_rank] || !c->dicts_built) { harcgpu_set_error("harcgpu_job_build_dicts first"); return -1; }
	CK(cudaSetDevice(c->device));
	if (c->shard_world > 1) {
		if (job_bcast_join(c)) return -1; // this GPU's slice has reached every replica before it says so at the barrier
		const u64 lo = std::min<u64>((u64)c->shard_rank * c->seg_per, c->shard_n), hi = std::min<u64>(lo + c->seg_per, c->shard_n);
		CK(cudaMemsetAsync(c->seg[c->shard_rank], 0, (size_t)c->seg_per / 8, c->st));
		if (s1_init_claim(c, c->seg[c->shard_rank], (u32)(hi - lo))) return -1;
		// (the barrier "every range of the bitmap is armed before any walker claims" is inside s1_reorder, right in front of
		// the walk kernel)
	}
	c->shard_ready = true;
	if (s1_reorder(c)) return -1;
	if (job_barrier(c)) return -1;     // nobody re-arms its range or rebuilds its shard while a peer still walks
	return job_check(c, "harcgpu_job_reorder");
}

int harcgpu_job_set_barrier(harcgpu_ctx *c, int (*fn)(void *), void *user)
{
	if (!c) { harcgpu_set_error("null argument"); return -1; }
	c->job_barrier_hook = fn; c->job_barrier_user = user;
	return 0;
}

int harcgpu_set_pool_exchange(harcgpu_ctx *c, int (*fn)(void *, void *, uint64_t), void *user)
{
	if (!c) { harcgpu_set_error("null argument"); return -1; }
	c->pool_exchange = fn; c->pool_exchange_user = user;
	return 0;
}

} // extern "C"
