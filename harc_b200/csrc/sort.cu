// Stable LSD radix sort of (u64 key, u32 value) pairs, 8 bits per pass -- the sort behind the dictionary build
// (reorder.cpp:305 std::sort of the keys + the CSR fill of 344-391 in one go: sorting (key, read id) pairs stably leaves
// the ids ascending inside a bin), the chunk order of stage I's finalize and the merge order of stage II.
//
// Per pass: rs_hist_kernel counts the digits of every tile of 4096 pairs into hist[digit][tile]; one exclusive scan over
// that array (digit-major) is at once the global start of every digit and the offset of every tile inside it;
// rs_scatter_kernel ranks the tile again (warp-level match of equal digits, item by item, so equal keys keep their
// order), stages the tile in shared memory in digit order and writes every digit's run contiguously.  HBM traffic per
// pass and pair: 8 B (histogram) + 12 B read + 12 B written.  The two buffers are used in turn; the caller gets told
// which one holds the result.
#include "ctx.h"
#include <utility>

namespace {
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS; // 4096
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const u64 *__restrict__ keys, size_t n, int shift, u32 dmask, u32 tiles,
                                                             u32 *__restrict__ hist)
{
	__shared__ u32 cnt[256];
	cnt[threadIdx.x] = 0;
	__syncthreads();
	const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const size_t idx = base + (size_t)i * RS_THREADS + threadIdx.x;
		if (idx < n) atomicAdd(&cnt[(u32)(__ldg(&keys[idx]) >> shift) & dmask], 1u);
	}
	__syncthreads();
	hist[(size_t)threadIdx.x * tiles + blockIdx.x] = cnt[threadIdx.x];
}

struct RsSmem {
	u64 key[RS_TILE];
	u32 val[RS_TILE];
	u32 wh[RS_WARPS][256]; // per warp: count, then start inside the tile's run of the digit
	u32 dstart[256];       // start of the digit's run inside the tile
	u32 gbase[256];        // global index of the first pair of the digit's run of this tile
	u32 wsum[RS_WARPS];
};

__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in,
                                                                u64 *__restrict__ keys_out, u32 *__restrict__ vals_out, size_t n, int shift,
                                                                u32 dmask, u32 tiles, const u32 *__restrict__ offs)
{
	extern __shared__ uint4 rs_raw[];
	RsSmem &s = *reinterpret_cast<RsSmem *>(rs_raw);
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const size_t tbase = (size_t)blockIdx.x * RS_TILE;
	const u32 tile_n = (u32)min((size_t)RS_TILE, n - tbase);
	for (int k = tid; k < RS_WARPS * 256; k += RS_THREADS) (&s.wh[0][0])[k] = 0;
	s.gbase[tid] = offs[(size_t)tid * tiles + blockIdx.x];
	__syncthreads();
	// warp-striped: warp w owns positions [512 w, 512 w + 512) of the tile, item i of lane l is position 512 w + 32 i + l,
	// so (item, lane) order is position order
	u64 key[RS_ITEMS];
	u32 rank[RS_ITEMS];
	const u32 lt = (1u << lane) - 1u;
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 p = (u32)w * (32 * RS_ITEMS) + 32 * i + lane;
		const bool valid = p < tile_n;
		key[i] = valid ? __ldg(&keys_in[tbase + p]) : 0ull;
		const u32 d = valid ? ((u32)(key[i] >> shift) & dmask) : (256u + lane); // pairs behind the end match nobody
		const u32 peers = __match_any_sync(0xffffffffu, d);
		const int leader = __ffs(peers) - 1;
		u32 old = 0;
		if (lane == leader && valid) { old = s.wh[w][d]; s.wh[w][d] = old + __popc(peers); }
		old = __shfl_sync(0xffffffffu, old, leader);
		rank[i] = old + __popc(peers & lt);
	}
	__syncthreads();
	// digit tid: starts of the warps' runs inside the digit's run, then the start of the digit's run inside the tile
	u32 run = 0;
#pragma unroll
	for (int k = 0; k < RS_WARPS; k++) { const u32 t = s.wh[k][tid]; s.wh[k][tid] = run; run += t; }
	u32 incl = run;
	for (int o = 1; o < 32; o <<= 1) {
		const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) s.wsum[w] = incl;
	__syncthreads();
	u32 before = incl - run;
	for (int k = 0; k < w; k++) before += s.wsum[k];
	s.dstart[tid] = before;
	__syncthreads();
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const u32 p = (u32)w * (32 * RS_ITEMS) + 32 * i + lane;
		if (p < tile_n) {
			const u32 d = (u32)(key[i] >> shift) & dmask;
			const u32 q = s.dstart[d] + s.wh[w][d] + rank[i];
			s.key[q] = key[i];
			s.val[q] = __ldg(&vals_in[tbase + p]);
		}
	}
	__syncthreads();
	for (u32 q = tid; q < tile_n; q += RS_THREADS) {
		const u64 k = s.key[q];
		const u32 d = (u32)(k >> shift) & dmask;
		const size_t dst = (size_t)s.gbase[d] + (q - s.dstart[d]);
		keys_out[dst] = k;
		vals_out[dst] = s.val[q];
	}
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
	const size_t hn = (size_t)256 * tiles;
	u32 *hist = nullptr, *offs = nullptr;
	u64 *scan_tmp = nullptr;
	if (c->alloc(&hist, hn) || c->alloc(&offs, hn) || c->alloc(&scan_tmp, scan_tmp_elems(hn))) return -1;
	CK(cudaFuncSetAttribute(rs_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
	for (int shift = begin_bit; shift < end_bit; shift += 8) {
		const u32 dmask = end_bit - shift >= 8 ? 255u : (1u << (end_bit - shift)) - 1u; // only bits below end_bit count
		rs_hist_kernel<<<KL + tiles, RS_THREADS, 0, st>>>(*keys, n, shift, dmask, tiles, hist);
		CK(cudaGetLastError());
		if (exclusive_scan_u32(hist, offs, hn, scan_tmp, nullptr, st)) return -1;
		rs_scatter_kernel<<<KL + tiles, RS_THREADS, sizeof(RsSmem), st>>>(*keys, *vals, *keys_alt, *vals_alt, n, shift, dmask, tiles, offs);
		CK(cudaGetLastError());
		std::swap(*keys, *keys_alt);
		std::swap(*vals, *vals_alt);
	}
	c->release(hist); c->release(offs); c->release(scan_tmp);
	return 0;
}
