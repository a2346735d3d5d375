// Stable LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass, single sweep per pass ("onesweep") -- the sort
// behind the dictionary build (reorder.cpp:305 std::sort of the keys + the CSR fill of 344-391 in one go: sorting
// (key, read id) pairs stably leaves the ids ascending inside a bin), the chunk order of stage I's finalize and the
// merge order of stage II.
//
// rs_ghist_kernel reads the keys once and counts the digits of ALL passes; their scans are the global starts of every
// digit.  A pass is then one kernel: a block takes the next tile of 4096 pairs (ticket), ranks it (warp-level match of
// equal digits, item by item, so equal keys keep their order), publishes its digit counts and adds up the counts of the
// tiles before it by decoupled look-back (a tile publishes first its own counts, then the inclusive prefix; a later
// tile walks back until it meets an inclusive one), stages the tile in shared memory in digit order and writes every
// digit's run contiguously.  HBM traffic per pass and pair: 12 B read + 12 B written, plus 8 B once for the histograms.
// The two buffers are used in turn and swapped, so the caller finds the result in *keys / *vals.
#include "ctx.h"
#include <algorithm>
#include <utility>

namespace {
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS; // 4096
constexpr int RS_WARPS = RS_THREADS / 32;

constexpr int RS_MAXPASS = 8;
constexpr u64 RS_PARTIAL = 1ull << 62, RS_INCLUSIVE = 1ull << 63, RS_COUNT = RS_PARTIAL - 1;

// digit counts of every pass: ghist[pass][digit]
__global__ void __launch_bounds__(RS_THREADS) rs_ghist_kernel(const u64 *__restrict__ keys, size_t n, int begin_bit, int end_bit, int passes,
                                                              u32 *__restrict__ ghist)
{
	__shared__ u32 cnt[RS_MAXPASS][256];
	for (int k = threadIdx.x; k < RS_MAXPASS * 256; k += RS_THREADS) (&cnt[0][0])[k] = 0;
	__syncthreads();
	for (size_t idx = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; idx < n; idx += (size_t)gridDim.x * RS_THREADS) {
		const u64 k = __ldg(&keys[idx]);
#pragma unroll
		for (int p = 0; p < RS_MAXPASS; p++) {
			if (p < passes) {
				const int shift = begin_bit + 8 * p;
				const u32 dmask = end_bit - shift >= 8 ? 255u : (1u << (end_bit - shift)) - 1u;
				atomicAdd(&cnt[p][(u32)(k >> shift) & dmask], 1u);
			}
		}
	}
	__syncthreads();
	for (int k = threadIdx.x; k < passes * 256; k += RS_THREADS) {
		const u32 v = (&cnt[0][0])[k];
		if (v) atomicAdd(&ghist[k], v);
	}
}
// gstart[pass][digit] = exclusive scan of ghist[pass][*]
__global__ void __launch_bounds__(256) rs_gscan_kernel(const u32 *__restrict__ ghist, u32 *__restrict__ gstart)
{
	__shared__ u32 wsum[8];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const u32 v = ghist[blockIdx.x * 256 + threadIdx.x];
	u32 incl = v;
	for (int o = 1; o < 32; o <<= 1) {
		const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) wsum[w] = incl;
	__syncthreads();
	u32 before = incl - v;
	for (int k = 0; k < w; k++) before += wsum[k];
	gstart[blockIdx.x * 256 + threadIdx.x] = before;
}

struct RsSmem {
	u64 key[RS_TILE];
	union { // the per-warp digit counters are dead once every pair knows its place in the tile
		u32 val[RS_TILE];
		u32 wh[RS_WARPS][256]; // per warp: count, then start inside the tile's run of the digit
	};
	u32 dstart[256];       // start of the digit's run inside the tile
	u64 gbase[256];        // global index of the first pair of the digit's run of this tile
	u32 wsum[8];
	u32 ticket;
};

__global__ void __launch_bounds__(RS_THREADS, 4) rs_scatter_kernel(const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
                                                                u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, size_t n, int shift,
                                                                u32 dmask, const u32 *__restrict__ gstart, u64 *lookback, u32 *ticket_ctr)
{
	extern __shared__ uint4 rs_raw[];
	RsSmem &s = *reinterpret_cast<RsSmem *>(rs_raw);
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	// tiles are taken in ticket order, so every tile a block looks back at belongs to a block that already runs
	if (tid == 0) s.ticket = atomicAdd(ticket_ctr, 1u);
	for (int k = tid; k < RS_WARPS * 256; k += RS_THREADS) (&s.wh[0][0])[k] = 0;
	__syncthreads();
	const u32 tile = s.ticket;
	const size_t tbase = (size_t)tile * RS_TILE;
	const u32 tile_n = (u32)min((size_t)RS_TILE, n - tbase);
	// warp-striped: warp w owns positions [512 w, 512 w + 512) of the tile, item i of lane l is position 512 w + 32 i + l,
	// so (item, lane) order is position order
	u64 key[RS_ITEMS];
	u32 rank[RS_ITEMS];
	const u32 lt = (1u << lane) - 1u;
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 p = (u32)w * (32 * RS_ITEMS) + 32 * i + lane;
		key[i] = p < tile_n ? __ldg(&keys_in[tbase + p]) : 0ull;
	}
	// the tile's digit counts first (plain shared-memory atomics), published at once: the tiles behind this one can
	// add them up while this one is still ranking
	if (tid < 256) s.dstart[tid] = 0;
	__syncthreads();
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 p = (u32)w * (32 * RS_ITEMS) + 32 * i + lane;
		if (p < tile_n) atomicAdd(&s.dstart[(u32)(key[i] >> shift) & dmask], 1u);
	}
	__syncthreads();
	volatile u64 *lb = lookback;
	if (tid < 256) lb[(size_t)tile * 256 + tid] = (u64)s.dstart[tid] | RS_PARTIAL;
	// the matches of all items first (independent of each other), then the counter updates in item order
	u32 peers[RS_ITEMS];
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 p = (u32)w * (32 * RS_ITEMS) + 32 * i + lane;
		const u32 d = p < tile_n ? ((u32)(key[i] >> shift) & dmask) : (256u + lane); // pairs behind the end match nobody
		peers[i] = __match_any_sync(0xffffffffu, d);
	}
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 p = (u32)w * (32 * RS_ITEMS) + 32 * i + lane;
		const bool valid = p < tile_n;
		const u32 d = (u32)(key[i] >> shift) & dmask;
		const int leader = __ffs(peers[i]) - 1;
		u32 old = 0;
		if (lane == leader && valid) { old = s.wh[w][d]; s.wh[w][d] = old + __popc(peers[i]); }
		old = __shfl_sync(0xffffffffu, old, leader);
		rank[i] = old + __popc(peers[i] & lt);
		__syncwarp(); // the leader of the next item may be another lane that reads the counter this one has just written
	}
	__syncthreads();
	// digit tid (the first 256 threads): starts of the warps' runs inside the digit's run, look-back, then the start of
	// the digit's run inside the tile
	u32 run = 0, incl = 0;
	if (tid < 256) {
#pragma unroll
		for (int k = 0; k < RS_WARPS; k++) { const u32 t = s.wh[k][tid]; s.wh[k][tid] = run; run += t; }
		// decoupled look-back over the tiles before this one, for digit tid
		u64 excl = 0;
		for (long long t = (long long)tile - 1; t >= 0; t--) {
			u64 v;
			do { v = lb[(size_t)t * 256 + tid]; } while ((v & (RS_PARTIAL | RS_INCLUSIVE)) == 0);
			excl += v & RS_COUNT;
			if (v & RS_INCLUSIVE) break;
		}
		lb[(size_t)tile * 256 + tid] = (excl + run) | RS_INCLUSIVE;
		s.gbase[tid] = (u64)gstart[tid] + excl;
		incl = run;
		for (int o = 1; o < 32; o <<= 1) {
			const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl += t;
		}
		if (lane == 31) s.wsum[w] = incl;
	}
	__syncthreads();
	if (tid < 256) {
		u32 before = incl - run;
		for (int k = 0; k < w; k++) before += s.wsum[k];
		s.dstart[tid] = before;
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) { // place of the pair in the tile, in digit order
		const u32 p = (u32)w * (32 * RS_ITEMS) + 32 * i + lane;
		const u32 d = (u32)(key[i] >> shift) & dmask;
		if (p < tile_n) rank[i] += s.dstart[d] + s.wh[w][d];
	}
	__syncthreads(); // the counters are dead: their memory takes the values now
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 p = (u32)w * (32 * RS_ITEMS) + 32 * i + lane;
		if (p < tile_n) {
			s.key[rank[i]] = key[i];
			s.val[rank[i]] = __ldg(&vals_in[tbase + p]);
		}
	}
	__syncthreads();
	for (u32 q = tid; q < tile_n; q += RS_THREADS) {
		const u64 k = s.key[q];
		const u32 d = (u32)(k >> shift) & dmask;
		const size_t dst = (size_t)(s.gbase[d] + (q - s.dstart[d]));
		keys_out[dst] = k;
		vals_out[dst] = s.val[q];
	}
}
// ---- full 64-bit order from a sort by the top 32 bits.  Mixed keys (common.cuh) are spread over all 64 bits, so after a
// stable sort by bits 32..63 a run of DIFFERENT keys that share their top half is a birthday collision: a few elements
// long, and rare (n^2 / 2^33 pairs).  One thread per run head puts such a run in order of the low half with a stable
// insertion sort; equal keys (the bins of a dictionary) are left as they are.  A run longer than RUN_MAX is reported and
// the caller falls back to the full sort.
constexpr int RUN_MAX = 32;
// After the four passes over the top half, pairs that share a top half are in input order.  Most such runs are already in
// key order -- above all the dictionary bins: runs of EQUAL keys, of any length on repetitive genomes -- and need
// nothing, so the work is driven by the inversions: a pair that is smaller than its predecessor in the same run walks
// back to the head of its run (a step or two for the chance collisions of random keys) and marks it; only marked runs
// are looked at again.  A bin of 10^5 equal keys costs nothing this way; one thread per run head scanning its whole run
// (round 1) cost 74 ms on a 3 M-read input with a long poly-A stretch.
__global__ void __launch_bounds__(256) runmark_kernel(const u64 *__restrict__ keys, size_t n, u32 *__restrict__ fixbits)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0 || i >= n) return;
	const u64 k = keys[i], kp = keys[i - 1];
	if ((k >> 32) != (kp >> 32) || k >= kp) return;
	size_t h = i - 1;
	while (h > 0 && (keys[h - 1] >> 32) == (k >> 32)) h--;
	atomicOr(&fixbits[h >> 5], 1u << (h & 31));
}
__global__ void __launch_bounds__(256) runfix_kernel(u64 *__restrict__ keys, u32 *__restrict__ vals, size_t n, const u32 *__restrict__ fixbits,
                                                     u32 *__restrict__ too_long)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !((fixbits[i >> 5] >> (i & 31)) & 1u)) return; // not the head of a run with an inversion
	const u64 k0 = keys[i];
	u64 k[RUN_MAX];
	u32 v[RUN_MAX];
	int len = 0;
	while (i + len < n && (keys[i + len] >> 32) == (k0 >> 32)) {
		if (len == RUN_MAX) { atomicMax(too_long, 1u); return; } // a long run that is out of order: the full sort takes over
		k[len] = keys[i + len];
		v[len] = vals[i + len];
		len++;
	}
	for (int a = 1; a < len; a++) { // stable insertion sort by the whole key
		const u64 ka = k[a];
		const u32 va = v[a];
		int b = a - 1;
		while (b >= 0 && k[b] > ka) { k[b + 1] = k[b]; v[b + 1] = v[b]; b--; }
		k[b + 1] = ka;
		v[b + 1] = va;
	}
	for (int a = 0; a < len; a++) { keys[i + a] = k[a]; vals[i + a] = v[a]; }
}
} // namespace

// Sorts the n pairs (*keys, *vals) by the key bits [begin_bit, end_bit), stably.  *keys / *vals and *keys_alt / *vals_alt
// are two buffers of n elements each; they are used in turn and swapped so that on return *keys / *vals hold the result.
// n < 2^32 (offsets are 32-bit).
int radix_sort_pairs(harcgpu_ctx *c, u64 **keys, u64 **keys_alt, u32 **vals, u32 **vals_alt, size_t n, int begin_bit, int end_bit)
{
	if (n == 0 || end_bit <= begin_bit) return 0;
	if (n >= 0xffffffffull) { harcgpu_set_error("radix sort: too many pairs"); return -1; }
	cudaStream_t st = c->st;
	const u32 tiles = (u32)((n + RS_TILE - 1) / RS_TILE);
	const int passes = (end_bit - begin_bit + 7) / 8;
	if (passes > RS_MAXPASS) { harcgpu_set_error("radix sort: more than 64 key bits"); return -1; }
	u32 *ghist = nullptr, *gstart = nullptr, *ticket = nullptr;
	u64 *lookback = nullptr;
	const size_t lbn = (size_t)passes * tiles * 256;
	if (c->alloc(&ghist, passes * 256) || c->alloc(&gstart, passes * 256) || c->alloc(&ticket, passes) || c->alloc(&lookback, lbn)) return -1;
	CK(cudaMemsetAsync(ghist, 0, sizeof(u32) * passes * 256, st));
	CK(cudaMemsetAsync(ticket, 0, sizeof(u32) * passes, st));
	CK(cudaMemsetAsync(lookback, 0, sizeof(u64) * lbn, st));
	CK(cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
	rs_ghist_kernel<<<KL + (unsigned)std::min<size_t>(tiles, 148 * 8), RS_THREADS, 0, st>>>(*keys, n, begin_bit, end_bit, passes, ghist);
	rs_gscan_kernel<<<KL + passes, 256, 0, st>>>(ghist, gstart);
	CK(cudaGetLastError());
	for (int p = 0; p < passes; p++) {
		const int shift = begin_bit + 8 * p;
		const u32 dmask = end_bit - shift >= 8 ? 255u : (1u << (end_bit - shift)) - 1u; // only bits below end_bit count
		rs_scatter_kernel<<<KL + tiles, RS_THREADS, sizeof(RsSmem), st>>>(*keys, *vals, *keys_alt, *vals_alt, n, shift, dmask, gstart + p * 256,
		                                                                  lookback + (size_t)p * tiles * 256, ticket + p);
		CK(cudaGetLastError());
		std::swap(*keys, *keys_alt);
		std::swap(*vals, *vals_alt);
	}
	c->release(ghist); c->release(gstart); c->release(ticket); c->release(lookback);
	return 0;
}

// Stable sort by all 64 key bits for keys that are spread over the whole 64-bit range (mixed keys): four passes over the
// top half, then the run fix-up; the full eight passes only if a run of different keys with one top half is longer than
// RUN_MAX (not expected: it would be a 33-fold birthday collision in 2^32).  force_full (test hook) takes the slow path.
int radix_sort_mixed(harcgpu_ctx *c, u64 **keys, u64 **keys_alt, u32 **vals, u32 **vals_alt, size_t n, bool force_full)
{
	if (n == 0) return 0;
	cudaStream_t st = c->st;
	if (!force_full) {
		u32 *flag = nullptr, *fixbits = nullptr, h = 0;
		const size_t nw = (n + 31) / 32;
		if (c->alloc(&flag, 1) || c->alloc(&fixbits, nw)) return -1;
		CK(cudaMemsetAsync(flag, 0, 4, st));
		CK(cudaMemsetAsync(fixbits, 0, 4 * nw, st));
		if (radix_sort_pairs(c, keys, keys_alt, vals, vals_alt, n, 32, 64)) return -1;
		runmark_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(*keys, n, fixbits);
		runfix_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(*keys, *vals, n, fixbits, flag);
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		c->release(flag); c->release(fixbits);
		if (!h) return 0;
	}
	// every step so far was stable, so the full sort can start from whatever order the pairs are in now
	return radix_sort_pairs(c, keys, keys_alt, vals, vals_alt, n, 0, 64);
}
