// Stage II of HARC -- consensus encoding (reference: src/encoder.cpp) -- as segmented sm_100a kernels.
//
// The reference walks the reordered stream contig by contig with std::list bookkeeping.  Here every contig is laid
// out on ONE global column axis (contig c occupies [base_c, base_c + L + sum of its position deltas)), so that
//   * G[i], the column where stream read i starts, is an inclusive prefix sum                   (encoder.cpp:246-251, 627-637)
//   * the consensus of a column is a vote over the reads G[lo..hi) that cover it                 (buildcontig, 619-652)
//   * re-aligning the pool (singletons ++ reads with N) is one probe per (column window, direction, dictionary);
//     a pool read goes to the smallest (column, probe) that accepts it = the order in which the sequential reference
//     would have claimed it (atomicMin on a priority word)                                       (encode, 231-418)
//   * the final read order is a merge of two sorted lists by column                              (list inserts 254-268, 310-313)
//   * noise / noisepos / pos / rev / order streams are scans + scatters over the merged list    (writecontig, 654-717)
//   * packbits is a bit shuffle of the 2-bit consensus array                                     (512-616)
// Consensus, reads and pool all live in the stage I 2-bit layout (A0 G1 C2 T3); the reference's 3-bit code is
// 2*code2 + nflag, so its 3-bit Hamming distance is popc(x2 ^ y2) + popc(nflags).
//
// Known deviation (documented in DESIGN.md): a pool dictionary bin with more than maxsearch (1000) live reads is
// scanned over its top 1000 ids only, without tracking removals; the result is still lossless.
#include "ctx.h"
#include <string.h>
#include <stdlib.h>
#include <utility>

namespace {
constexpr u32 CAP1 = 10000001u; // encoder.cpp:226: a contig is cut once list_size > 10000000
// priority word of a pool read: rank << 56 | column << 2 | probe kind.  Both sentinels are positive as int64 so that the
// multi-GPU exchange can be a plain all-reduce(min) over int64.
constexpr u64 NOBEST = 0x7fffffffffffffffull;    // no contig took the read
constexpr u64 ELSEWHERE = 0x7ffffffffffffffeull; // another GPU's contig took it (or: unaligned, and rank 0 writes those)
constexpr int RANK_SHIFT = 56;

__device__ __forceinline__ u64 revpairs64(u64 x)
{
	u64 y = __brevll(x);
	return ((y & 0x5555555555555555ull) << 1) | ((y >> 1) & 0x5555555555555555ull);
}
__device__ __forceinline__ u64 swappairs64(u64 x) // bit code (A0 G1 C2 T3) <-> file code (A0 C1 G2 T3)
{
	return ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
}
// reverse the base order of an L-base sequence (2 bits/base) held in NW words; no complement
template <int NW>
__device__ __forceinline__ void reverse2(const u64 (&in)[NW], int L, u64 (&out)[NW])
{
	u64 t[NW + 1];
#pragma unroll
	for (int k = 0; k < NW; k++) t[k] = revpairs64(in[NW - 1 - k]);
	t[NW] = 0;
	const int s = 2 * (32 * NW - L);
#pragma unroll
	for (int k = 0; k < NW; k++) out[k] = s ? (t[k] >> s) | (t[k + 1] << (64 - s)) : t[k];
}
template <int NW>
__device__ __forceinline__ void load_words(const u64 *__restrict__ p, u64 (&w)[NW])
{
#pragma unroll
	for (int k = 0; k < NW; k++) w[k] = __ldg(&p[k]);
}
// reverse complement of a pool read (N stays N): r2 = codes, rn = N flags
template <int NW>
__device__ __forceinline__ void revcomp_pool(u64 (&r2)[NW], u64 (&rn)[NW], int L)
{
	u64 a[NW], b[NW];
	reverse2<NW>(r2, L, a);
	reverse2<NW>(rn, L, b);
#pragma unroll
	for (int k = 0; k < NW; k++) {
		u64 valid = lowmask(2 * L - 64 * k);
		r2[k] = (a[k] ^ valid) & ~(b[k] | (b[k] << 1));
		rn[k] = b[k];
	}
}
// bits [2g, 2g+2L) of the consensus bit array (two zero pad words follow the array)
template <int NW>
__device__ __forceinline__ void load_window(const u64 *__restrict__ cons2, u64 g, int L, u64 (&w)[NW])
{
	const u64 bit = 2 * g;
	const size_t q = (size_t)(bit >> 6);
	const int r = (int)(bit & 63);
#pragma unroll
	for (int k = 0; k < NW; k++) {
		u64 lo = __ldg(&cons2[q + k]), hi = __ldg(&cons2[q + k + 1]);
		w[k] = (r ? (lo >> r) | (hi << (64 - r)) : lo) & lowmask(2 * L - 64 * k);
	}
}
__device__ __forceinline__ u64 getbits_g(const u64 *__restrict__ w, u64 bit, int n)
{
	const size_t q = (size_t)(bit >> 6);
	const int r = (int)(bit & 63);
	u64 v = __ldg(&w[q]) >> r;
	if (r) v |= __ldg(&w[q + 1]) << (64 - r);
	if (n < 64) v &= (1ull << n) - 1;
	return v;
}
// first index in [0,n) with a[i] > v
__device__ __forceinline__ u32 upper_bound64(const u64 *__restrict__ a, u32 n, u64 v)
{
	u64 lo = 0, hi = n;
	while (lo < hi) {
		u64 mid = (lo + hi) >> 1;
		if (__ldg(&a[mid]) <= v) lo = mid + 1; else hi = mid;
	}
	return (u32)lo;
}

// ---------------------------------------------------------------------------------------------- contig layout
__global__ void __launch_bounds__(256) natstart_kernel(const u8 *__restrict__ flag, u32 m, u32 per, u32 *__restrict__ ns)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	ns[i] = (flag[i] == '0' || i % per == 0) ? 1u : 0u; // flag '0' or first read of a thread range (encoder.cpp:226, 169-180)
}
__global__ void __launch_bounds__(256) scatter_idx_kernel(const u32 *__restrict__ f, const u32 *__restrict__ ex, u32 m, u32 *__restrict__ out)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < m && f[i]) out[ex[i]] = i;
}
__global__ void __launch_bounds__(256) cstart_kernel(const u32 *__restrict__ ns, const u32 *__restrict__ ex_ns, const u32 *__restrict__ nat_idx,
                                                     const u8 *__restrict__ pos, u32 m, int L, u32 *__restrict__ cs, u64 *__restrict__ inc)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	u32 c = ex_ns[i] + ns[i] - 1;
	u32 idx = i - nat_idx[c];
	u32 s = (idx % CAP1 == 0) ? 1u : 0u;
	cs[i] = s;
	inc[i] = s ? (i ? (u64)L : 0ull) : (u64)pos[i];
}
__global__ void __launch_bounds__(256) finish_layout_kernel(const u32 *__restrict__ cs, const u32 *__restrict__ ex_cs, const u64 *__restrict__ inc,
                                                            u64 *__restrict__ G, u32 m, u32 *__restrict__ cid, u32 *__restrict__ cstart)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	G[i] += inc[i]; // exclusive -> inclusive
	u32 c = ex_cs[i] + cs[i] - 1;
	cid[i] = c;
	if (cs[i]) cstart[c] = i;
}

// T[k] = first stream read that starts at column >= TILE_C * k: lets a block find the reads around its columns with two
// loads instead of two 25-step binary searches by one thread
constexpr int TILE_C = 128;
__global__ void __launch_bounds__(256) tile_index_kernel(const u64 *__restrict__ G, u32 m, u32 nt, u32 *__restrict__ T)
{
	u32 k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nt) return;
	const u64 v = (u64)TILE_C * k;
	u64 lo = 0, hi = m;
	while (lo < hi) {
		u64 mid = (lo + hi) >> 1;
		if (__ldg(&G[mid]) < v) lo = mid + 1; else hi = mid;
	}
	T[k] = (u32)lo;
}

// ---------------------------------------------------------------------------------------------- consensus
// buildcontig (encoder.cpp:619-652).  A block owns CONS_T consecutive columns.  The stream reads that cover them are a
// contiguous index range (G is non-decreasing); they are staged in shared memory in chunks (start column relative to
// the tile + packed bases), and every warp then runs over the reads that touch its 32 columns with all lanes on the
// same read: the start column is a broadcast, the packed words of a read are read by neighbouring lanes from the same
// one or two words, and a lane adds the base that falls on its column to four 16-bit counters packed in one u64
// (flushed to 32-bit counters after every chunk, so nothing can overflow).
constexpr int CONS_T = 256;   // columns per block
constexpr int CONS_R = 512;   // reads per staged chunk
__global__ void __launch_bounds__(CONS_T) consensus_kernel(const u64 *__restrict__ G, const u32 *__restrict__ T, const u32 *__restrict__ sreads32,
                                                           u32 m, int L, int W2, u64 TOT, u64 *__restrict__ cons2)
{
	extern __shared__ u32 cons_smem[];
	int *sG = reinterpret_cast<int *>(cons_smem);      // [CONS_R] start column - c0
	u32 *sW = cons_smem + CONS_R;                      // [CONS_R][W2]
	const int tid = threadIdx.x, lane = tid & 31;
	const u64 c0 = (u64)blockIdx.x * CONS_T;
	// reads that may cover the tile: from the first read of the TILE_C-column cell that holds column c0 - L + 1 (a few
	// more than needed) to the first read that starts behind the tile
	const u32 rlo = c0 >= (u64)L ? __ldg(&T[(c0 - L + 1) / TILE_C]) : 0u;
	const u32 rhi = __ldg(&T[(c0 + CONS_T) / TILE_C]);
	const int wc0 = tid & ~31;                         // first column of this warp, relative to c0
	u32 cA = 0, cG = 0, cC = 0, cT = 0;
	for (u32 base = rlo; base < rhi; base += CONS_R) {
		const u32 nch = min((u32)CONS_R, rhi - base);
		for (u32 k = tid; k < nch; k += CONS_T) sG[k] = (int)(long long)(__ldg(&G[base + k]) - c0);
		for (u32 k = tid; k < nch * (u32)W2; k += CONS_T) sW[k] = __ldg(&sreads32[(size_t)base * W2 + k]);
		__syncthreads();
		// reads of the chunk that touch this warp's columns: start in (wc0 - L, wc0 + 31]
		u32 klo = 0, khi = nch;
		{
			u32 lo = 0, hi = nch;
			while (lo < hi) { u32 mid = (lo + hi) >> 1; if (sG[mid] <= wc0 - L) lo = mid + 1; else hi = mid; }
			klo = lo;
			hi = nch;
			while (lo < hi) { u32 mid = (lo + hi) >> 1; if (sG[mid] <= wc0 + 31) lo = mid + 1; else hi = mid; }
			khi = lo;
		}
		u64 cnt = 0; // four 16-bit counters, bit-code order A G C T; a chunk adds at most CONS_R votes
#pragma unroll 4
		for (u32 k = klo; k < khi; k++) {
			const u32 o = (u32)(tid - sG[k]);
			if (o < (u32)L) {
				const u32 v = (sW[k * W2 + (o >> 4)] >> (2 * (o & 15))) & 3u;
				cnt += 1ull << (16 * v);
			}
		}
		cA += (u32)cnt & 0xffffu; cG += (u32)(cnt >> 16) & 0xffffu; cC += (u32)(cnt >> 32) & 0xffffu; cT += (u32)(cnt >> 48);
		__syncthreads();
	}
	const u64 g = c0 + tid;
	u32 code = 0;
	if (g < TOT) {
		// ties -> A < C < G < T, strict '>' from max = 0 (encoder.cpp:642-648)
		u32 mx = 0;
		if (cA > mx) { mx = cA; code = 0; }
		if (cC > mx) { mx = cC; code = 2; }
		if (cG > mx) { mx = cG; code = 1; }
		if (cT > mx) { mx = cT; code = 3; }
	}
	u32 b0 = __ballot_sync(0xffffffffu, code & 1u), b1 = __ballot_sync(0xffffffffu, (code >> 1) & 1u);
	if (lane == 0 && (g < TOT)) {
		u64 lo = b0, hi = b1, v = 0;
		// interleave
		lo = (lo | (lo << 16)) & 0x0000FFFF0000FFFFull; lo = (lo | (lo << 8)) & 0x00FF00FF00FF00FFull;
		lo = (lo | (lo << 4)) & 0x0F0F0F0F0F0F0F0Full; lo = (lo | (lo << 2)) & 0x3333333333333333ull;
		lo = (lo | (lo << 1)) & 0x5555555555555555ull;
		hi = (hi | (hi << 16)) & 0x0000FFFF0000FFFFull; hi = (hi | (hi << 8)) & 0x00FF00FF00FF00FFull;
		hi = (hi | (hi << 4)) & 0x0F0F0F0F0F0F0F0Full; hi = (hi | (hi << 2)) & 0x3333333333333333ull;
		hi = (hi | (hi << 1)) & 0x5555555555555555ull;
		v = lo | (hi << 1);
		cons2[g >> 5] = v;
	}
}

// The same vote, bit-sliced: one THREAD owns 16 consecutive columns (one u32 of the packed consensus).  A read that
// touches them contributes a 32-bit window of its packed bases (one funnel shift), which is split into four one-hot
// masks (one per base code, on the even bit positions); the masks are added to bit-sliced counters -- three low planes
// per base, flushed every seven reads into twelve planes -- so a read costs ~80 instructions for 16 columns instead of
// ~14 per column.  The argmax with ties A < C < G < T is a bit-sliced comparison.  Columns with more than 4000 covering
// reads raise `deep`; the caller then runs consensus_kernel (exact for any depth) instead.
constexpr int CS_THREADS = 128;
constexpr int CS_COLS = 16;
constexpr int CS_BLOCK_COLS = CS_THREADS * CS_COLS; // 2048
constexpr int CS_SG = 4096;   // staged read starts per block
constexpr int CS_PLANES = 12;
__global__ void __launch_bounds__(CS_THREADS) consensus_sliced_kernel(const u64 *__restrict__ G, const u32 *__restrict__ T, u32 nt,
                                                                      const u32 *__restrict__ sreads32, u32 m, int L, int W2, u64 TOT,
                                                                      u32 *__restrict__ cons32, u32 *__restrict__ deep)
{
	__shared__ int sG[CS_SG];
	const int tid = threadIdx.x;
	const u64 cb = (u64)blockIdx.x * CS_BLOCK_COLS;
	const u32 rlo = cb >= (u64)L ? __ldg(&T[(cb - L + 1) / TILE_C]) : 0u;
	const u32 rhi = __ldg(&T[min((u64)nt - 1, (cb + CS_BLOCK_COLS) / TILE_C)]);
	const bool staged = rhi - rlo <= (u32)CS_SG;
	if (staged)
		for (u32 k = tid; k < rhi - rlo; k += CS_THREADS) sG[k] = (int)(long long)(__ldg(&G[rlo + k]) - cb);
	__syncthreads();
	const u64 c0 = cb + (u64)CS_COLS * tid;
	if (c0 >= (TOT + 31) / 32 * 32) return;
	const int r0 = CS_COLS * tid;
	// reads that touch the 16 columns: start in (c0 - L, c0 + 15]
	u32 lo, hi;
	if (staged) {
		u32 a = 0, b = rhi - rlo;
		while (a < b) { u32 mid = (a + b) >> 1; if (sG[mid] <= r0 - L) a = mid + 1; else b = mid; }
		lo = a;
		b = rhi - rlo;
		while (a < b) { u32 mid = (a + b) >> 1; if (sG[mid] <= r0 + CS_COLS - 1) a = mid + 1; else b = mid; }
		hi = a;
		lo += rlo; hi += rlo;
	} else {
		lo = c0 >= (u64)L ? upper_bound64(G, m, c0 - L) : 0u;
		hi = upper_bound64(G, m, c0 + CS_COLS - 1);
	}
	if (hi - lo > 4000u) { atomicOr(deep, 1u); return; }
	const u32 M = 0x55555555u;
	u32 f[4][CS_PLANES], l[4][3]; // base code order A G C T
#pragma unroll
	for (int b = 0; b < 4; b++) {
#pragma unroll
		for (int p = 0; p < CS_PLANES; p++) f[b][p] = 0;
		l[b][0] = l[b][1] = l[b][2] = 0;
	}
	auto flush = [&]() {
#pragma unroll
		for (int b = 0; b < 4; b++) {
			u32 a = f[b][0], cy = a & l[b][0];
			f[b][0] = a ^ l[b][0];
#pragma unroll
			for (int p = 1; p < 3; p++) {
				a = f[b][p];
				const u32 x = a ^ l[b][p];
				f[b][p] = x ^ cy;
				cy = (a & l[b][p]) | (cy & x);
			}
#pragma unroll
			for (int p = 3; p < CS_PLANES; p++) { a = f[b][p]; f[b][p] = a ^ cy; cy &= a; }
			l[b][0] = l[b][1] = l[b][2] = 0;
		}
	};
	int since = 0;
	for (u32 i = lo; i < hi; i++) {
		const int off = staged ? r0 - sG[i - rlo] : (int)((long long)c0 - (long long)__ldg(&G[i])); // column c0 is base `off` of the read
		const int bit = 2 * off, q = bit >> 5, r = bit & 31;
		const u32 *rw = sreads32 + (size_t)i * W2;
		const u32 w0 = (q >= 0 && q < W2) ? __ldg(&rw[q]) : 0u;
		const u32 w1 = (q + 1 >= 0 && q + 1 < W2) ? __ldg(&rw[q + 1]) : 0u;
		const u32 x = __funnelshift_r(w0, w1, r);
		const int klo = max(0, -off), khi = min(CS_COLS, L - off); // columns klo..khi-1 of the tile lie on the read
		const u32 hm = khi >= 16 ? ~0u : ((1u << (2 * khi)) - 1u), lm = (1u << (2 * klo)) - 1u;
		const u32 vm = M & hm & ~lm;
		const u32 e = x & vm, o = (x >> 1) & vm;
		const u32 h[4] = { vm & ~(e | o), e & ~o, o & ~e, e & o };
#pragma unroll
		for (int b = 0; b < 4; b++) {
			const u32 t0 = l[b][0] & h[b];
			l[b][0] ^= h[b];
			const u32 t1 = l[b][1] & t0;
			l[b][1] ^= t0;
			l[b][2] ^= t1;
		}
		if (++since == 7) { flush(); since = 0; }
	}
	flush();
	// argmax, ties -> A < C < G < T with strict '>' (encoder.cpp:642-648): start from A, then C (code 2), G (1), T (3)
	u32 best[CS_PLANES], clo = 0, chi = 0;
#pragma unroll
	for (int p = 0; p < CS_PLANES; p++) best[p] = f[0][p];
	const int order[3] = { 2, 1, 3 }; // indices into f: C, G, T
#pragma unroll
	for (int s3 = 0; s3 < 3; s3++) {
		const int b = order[s3];
		u32 gt = 0, eq = M;
#pragma unroll
		for (int p = CS_PLANES - 1; p >= 0; p--) {
			gt |= eq & f[b][p] & ~best[p];
			eq &= ~(f[b][p] ^ best[p]);
		}
#pragma unroll
		for (int p = 0; p < CS_PLANES; p++) best[p] = (f[b][p] & gt) | (best[p] & ~gt);
		// code of base index b in the packed layout is b itself (A0 G1 C2 T3)
		clo = (b & 1) ? (clo | gt) : (clo & ~gt);
		chi = (b & 2) ? (chi | gt) : (chi & ~gt);
	}
	cons32[c0 / CS_COLS] = clo | (chi << 1);
}

// ---------------------------------------------------------------------------------------------- pool re-alignment
struct PoolArgs {
	const u64 *G; const u32 *cid; const u32 *cstart;
	u32 m, NC, per;
	u64 TOT;
	const u64 *cons2;
	const u64 *pool, *poolN;
	DictView d[2];
	int L, thresh_s, maxsearch;
	u64 *best;
	u32 *flags;    // [0]: a scan of a bin beyond maxsearch ended with entries left; [1]: such a scan lowered a priority in this pass
	unsigned long long *pcur; // cursor cache of the bins beyond 8 x maxsearch
	u32 pcur_mask;
	u64 rank_bits; // rank << RANK_SHIFT
	const u32 *bloom; u32 bloom_mask;
	const u32 *T; // tile index
	u32 nt;       // its entries
	u64 cwords;   // words of cons2 incl. the two zero words at its end
};

// Blocked Bloom filter over the keys of both pool dictionaries (one 32-bit word per key, two bits in it): most window
// keys of the consensus are in neither dictionary, and the filter (a few MB, L2 resident) answers those without
// touching the key tables in DRAM.
__device__ __forceinline__ void bloom_pos(u64 key, int l, u32 mask, u32 &word, u32 &bits)
{
	u64 h = (key ^ (l ? 0x9E3779B97F4A7C15ull : 0ull)) * 0xD6E8FEB86659FD93ull;
	h ^= h >> 32;
	word = (u32)(h >> 10) & mask;
	bits = (1u << ((u32)h & 31u)) | (1u << (((u32)h >> 5) & 31u));
}
__global__ void __launch_bounds__(256) bloom_insert_kernel(const u64 *__restrict__ keys, u32 nk, int l, u32 *bloom, u32 mask)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nk) return;
	u32 w, b;
	bloom_pos(key_unmix(keys[i]), l, mask, w, b); // the dictionary stores mixed keys (common.cuh)
	atomicOr(&bloom[w], b);
}

// One thread per PP_CPT consecutive window starts.  The consensus bits of the block's span and the start columns of the
// reads around it are staged in shared memory; a thread computes the four window keys of its first column from scratch
// and then ROLLS them: a forward key drops its first base and takes the next consensus base on top, a reverse key (the
// reverse complement of a window that also moves right) shifts up and takes the complemented base at the bottom -- a
// handful of instructions per key and column instead of two 64-bit extractions, a pair reversal and a five-step spread.
constexpr int PP_CPT = 8;
constexpr int PP_THREADS = 128;
constexpr int PP_COLS = PP_CPT * PP_THREADS; // 1024 window starts per block = 8 cells of the tile index
constexpr int PP_R = 2048;                   // stream reads whose start columns are staged per block
constexpr int PP_CW = (PP_COLS + 256) / 32 + 2; // consensus words of the block's span (columns + read length)
template <int NW>
__global__ void __launch_bounds__(PP_THREADS) pool_probe_kernel(PoolArgs a)
{
	__shared__ int sG[PP_R];
	__shared__ u64 sC[PP_CW];
	const int tid = threadIdx.x;
	const u64 gb = (u64)blockIdx.x * PP_COLS;
	const int L = a.L;
	const u64 glast = a.TOT - L; // last window start; the launch guarantees TOT >= L
	for (int k = tid; k < PP_CW; k += PP_THREADS) {
		const u64 wi = gb / 32 + k;
		sC[k] = wi < a.cwords ? __ldg(&a.cons2[wi]) : 0ull;
	}
	// r = last read with G[r] <= g tells which contig column g belongs to: the start columns of the reads around the block
	// (from the last read before it to the first one behind it, via the tile index) are staged and searched
	const u32 cell = blockIdx.x * (PP_COLS / TILE_C);
	const u32 t0 = __ldg(&a.T[min(cell, a.nt - 1)]);
	const u32 r0 = t0 ? t0 - 1 : 0u, r1 = __ldg(&a.T[min(cell + PP_COLS / TILE_C, a.nt - 1)]);
	const bool staged = r1 - r0 <= (u32)PP_R;
	if (staged)
		for (u32 k = tid; k < r1 - r0; k += PP_THREADS) sG[k] = (int)(long long)(__ldg(&a.G[r0 + k]) - gb);
	__syncthreads();
	const int rel0 = PP_CPT * tid;
	if (gb + rel0 > glast) return;
	u32 r;
	if (staged) {
		u32 lo = 0, hi = r1 - r0;
		while (lo < hi) { u32 mid = (lo + hi) >> 1; if (sG[mid] <= rel0) lo = mid + 1; else hi = mid; }
		r = r0 + lo - 1;
	} else r = upper_bound64(a.G, a.m, gb + rel0) - 1;
	auto base_at = [&](int rel) -> u64 { return (sC[rel >> 5] >> (2 * (rel & 31))) & 3ull; }; // consensus base of column gb + rel
	auto bits_at = [&](int rel, int n) -> u64 { // 2n bits from column gb + rel on (n <= 32)
		const int q = rel >> 5, sh = 2 * (rel & 31);
		u64 v = sC[q] >> sh;
		if (sh) v |= sC[q + 1] << (64 - sh);
		return n < 32 ? v & ((1ull << (2 * n)) - 1) : v;
	};
	const int nb0 = a.d[0].dend - a.d[0].dstart + 1, nb1 = a.d[1].dend - a.d[1].dstart + 1;
	u64 k3s[4] = { 0, 0, 0, 0 };
	u32 c_cached = 0xffffffffu;
	bool col_ok = false;
	u64 g_limit = 0; // a window may start at g only if g + L <= g_limit (start of the next contig)
	for (int t = 0; t < PP_CPT; t++) {
		const int rel = rel0 + t;
		const u64 g = gb + rel;
		if (g > glast) break;
		// the four window keys: forward dict 0, forward dict 1, reverse dict 0, reverse dict 1 (encoder.cpp:270, 338)
#pragma unroll
		for (int q = 0; q < 4; q++) {
			const int l = q & 1, nb = l ? nb1 : nb0, ds = a.d[l].dstart, de = a.d[l].dend;
			if (t == 0) {
				if (q < 2) k3s[q] = spread2to3(bits_at(rel + ds, nb), nb);
				else {
					const u64 k2 = bits_at(rel + L - 1 - de, nb);
					k3s[q] = spread2to3(revpairs64(~k2 & lowmask(2 * nb)) >> (64 - 2 * nb), nb);
				}
			} else if (q < 2) k3s[q] = (k3s[q] >> 3) | (base_at(rel + de) << (3 * (nb - 1) + 1));
			else k3s[q] = ((k3s[q] << 3) & ((1ull << (3 * nb)) - 1)) | ((base_at(rel + L - 1 - ds) ^ 3ull) << 1);
		}
		// may a window start here?  (advance r to the last read that starts at or before g)
		if (staged) { while (r + 1 < r1 && sG[r + 1 - r0] <= rel) r++; }
		else { while (r + 1 < a.m && __ldg(&a.G[r + 1]) <= g) r++; }
		const u32 c = __ldg(&a.cid[r]);
		if (c != c_cached) {
			c_cached = c;
			const u32 next = c + 1 < a.NC ? __ldg(&a.cstart[c + 1]) : a.m;
			// the last contig of a thread range is written without alignment (encoder.cpp:438-441)
			col_ok = !(next == a.m || next % a.per == 0);
			g_limit = col_ok ? __ldg(&a.G[next]) : 0ull;
		}
		if (!col_ok || g + L > g_limit) continue; // j <= ref.size()-readlen (encoder.cpp:252)
		// Bloom words of the four keys first, so that the four L2 round trips overlap
		u32 bword[4], bbits[4];
#pragma unroll
		for (int q = 0; q < 4; q++) {
			u32 bw;
			bloom_pos(k3s[q], q & 1, a.bloom_mask, bw, bbits[q]);
			bword[q] = __ldg(&a.bloom[bw]);
		}
		u64 w[NW], rc[NW];
		bool have_w = false, have_rc = false;
#pragma unroll
		for (int q = 0; q < 4; q++) {
			if ((bword[q] & bbits[q]) != bbits[q]) continue;
			const int l = q & 1;
			const bool rev = q >= 2;
			const DictView &dv = a.d[l];
			u32 bstart, bsize;
			if (!dict_lookup(dv, k3s[q], bstart, bsize)) continue;
			if (!have_w) { load_window<NW>(a.cons2, g, L, w); have_w = true; }
			if (rev && !have_rc) {
				u64 tt[NW];
				reverse2<NW>(w, L, tt);
#pragma unroll
				for (int k = 0; k < NW; k++) rc[k] = tt[k] ^ lowmask(2 * L - 64 * k);
				have_rc = true;
			}
			const u64 myprio = a.rank_bits | (g << 2) | (u64)q;
			if (bsize <= (u32)a.maxsearch) {
				for (u32 e = bsize; e-- > 0u;) { // from the tail, no break: every read of the bin within thresh_s is taken (293-317)
					const u32 rid = bin_entry(dv, bstart, bsize, e);
					const u64 *p2 = a.pool + (size_t)rid * NW, *pn = a.poolN + (size_t)rid * NW;
					int d = 0;
#pragma unroll
					for (int k = 0; k < NW; k++) d += __popcll((rev ? rc[k] : w[k]) ^ __ldg(&p2[k])) + __popcll(__ldg(&pn[k]));
					if (d <= a.thresh_s) atomicMin(&a.best[rid], myprio);
				}
			} else {
				// A bin beyond maxsearch.  The reference removes a read from its bins the moment a window takes it
				// (encoder.cpp:1010-1031), so a later window scans the last maxsearch reads that are still LIVE (encoder.cpp:293).
				// Here all windows probe at once: a read counts as live for this window unless a window of higher priority
				// (= earlier in the reference's sequential order) holds it.  The host repeats the probe until no priority
				// moves; priorities only fall and never below the sequential result, so the fixed point is that result.
				// (A window gives up after 8 x maxsearch entries: the reference never meets the dead ones because it compacts its
				// bins, here they are stepped over one by one, and a bin of 10^5 reads probed from 10^4 windows -- a poly-A run --
				// would otherwise cost minutes.  Bins up to 8 x maxsearch are exact.)
				// Bins beyond 8 x maxsearch (where the result is approximate anyway) also keep a cursor in a small direct-mapped
				// cache: an index above which every read has been taken by SOME window, so the scan starts below the part of the
				// bin that is used up instead of stepping over it read by read, window after window and pass after pass.
				int live = 0;
				u32 e = bsize;
				const bool huge = bsize > 8u * (u32)a.maxsearch && bsize < (1u << 28);
				const u64 ctag = ((u64)bstart << 32) | ((u64)l << 31);
				unsigned long long *ce = a.pcur + ((bstart * 2u + (u32)l) & a.pcur_mask);
				if (huge) {
					const u64 cv = *((volatile unsigned long long *)ce);
					if ((cv & ~0x0fffffffull) == ctag) e = min(e, (u32)(cv & 0x0fffffffull));
				}
				const u32 top = e;
				bool used_up = huge; // every read from `top` down to here is taken
				const u32 stop = e > 8u * (u32)a.maxsearch ? e - 8u * (u32)a.maxsearch : 0u;
				while (e-- > stop && live < a.maxsearch) {
					const u32 rid = bin_entry(dv, bstart, bsize, e);
					const u64 owner = *((volatile u64 *)&a.best[rid]);
					if (used_up && owner == NOBEST) {
						used_up = false;
						if (e + 1u < top) *((volatile unsigned long long *)ce) = ctag | (u64)(e + 1u);
					}
					if (owner < myprio) continue; // taken earlier: not in the bin any more
					live++;
					const u64 *p2 = a.pool + (size_t)rid * NW, *pn = a.poolN + (size_t)rid * NW;
					int d = 0;
#pragma unroll
					for (int k = 0; k < NW; k++) d += __popcll((rev ? rc[k] : w[k]) ^ __ldg(&p2[k])) + __popcll(__ldg(&pn[k]));
					if (d <= a.thresh_s && atomicMin(&a.best[rid], myprio) > myprio) a.flags[1] = 1u;
				}
				if (used_up && e + 1u < top && e != 0xffffffffu) *((volatile unsigned long long *)ce) = ctag | (u64)(e + 1u);
				if (live >= a.maxsearch && e != 0xffffffffu && e + 1u > stop) a.flags[0] = 1u;
			}
		}
	}
}

__global__ void __launch_bounds__(256) fill64_kernel(u64 *p, size_t n, u64 v)
{
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = v;
}
// after the exchange: keep the reads this GPU's contigs won (rank bits stripped), mark the others
__global__ void __launch_bounds__(256) best_localize_kernel(u64 *best, u32 P, int rank)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P) return;
	u64 b = best[i];
	if (b == NOBEST) { if (rank != 0) best[i] = ELSEWHERE; }
	else if ((int)(b >> RANK_SHIFT) != rank) best[i] = ELSEWHERE;
	else best[i] = b & ((1ull << RANK_SHIFT) - 1);
}
// aligned pool reads in DESCENDING id order (the bin scan order, encoder.cpp:293), then stably sorted by priority
__global__ void __launch_bounds__(256) aligned_flag_kernel(const u64 *__restrict__ best, u32 P, u32 *__restrict__ af)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < P) af[i] = best[P - 1 - i] < ELSEWHERE;
}
__global__ void __launch_bounds__(256) aligned_compact_kernel(const u64 *__restrict__ best, const u32 *__restrict__ af, const u32 *__restrict__ ex,
                                                              u32 P, u64 *__restrict__ prio, u32 *__restrict__ rid)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P || !af[i]) return;
	prio[ex[i]] = best[P - 1 - i];
	rid[ex[i]] = P - 1 - i;
}

// ---------------------------------------------------------------------------------------------- merge
__global__ void __launch_bounds__(256) place_orig_kernel(const u64 *__restrict__ G, u32 m, const u64 *__restrict__ iprio, u32 M,
                                                         u32 *__restrict__ f_src, u8 *__restrict__ f_kind, u64 *__restrict__ f_col)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= m) return;
	u64 g = G[i];
	u64 lo = 0, hi = M; // inserted reads with column < g come first; at equal column the original comes first (254-268)
	while (lo < hi) {
		u64 mid = (lo + hi) >> 1;
		if ((__ldg(&iprio[mid]) >> 2) < g) lo = mid + 1; else hi = mid;
	}
	size_t e = (size_t)i + lo;
	f_src[e] = i; f_kind[e] = 0; f_col[e] = g;
}
__global__ void __launch_bounds__(256) place_ins_kernel(const u64 *__restrict__ G, u32 m, const u64 *__restrict__ iprio, const u32 *__restrict__ irid,
                                                        u32 M, u32 *__restrict__ f_src, u8 *__restrict__ f_kind, u64 *__restrict__ f_col)
{
	u32 k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= M) return;
	u64 pr = iprio[k], col = pr >> 2;
	size_t e = (size_t)k + upper_bound64(G, m, col);
	f_src[e] = irid[k]; f_kind[e] = (u8)(1u | (((pr & 3ull) >= 2) ? 2u : 0u)); f_col[e] = col;
}

// ---------------------------------------------------------------------------------------------- emission
struct EmitArgs {
	const u32 *f_src; const u8 *f_kind; const u64 *f_col; u64 F;
	const u64 *sreads; const u32 *cs; const u32 *s_order; const u8 *s_rev;
	const u64 *pool, *poolN; const u32 *pool_order; u32 n_s;
	const u64 *cons2;
	int L;
	// pass 1 outputs
	u64 *nm1; u8 *posb; u8 *revc; u32 *isN; u32 *ordv;
	// pass 2 inputs / outputs
	const u64 *noff; const u32 *exN;
	char *noise; u8 *noisepos; u32 *o_order; u32 *o_order_N;
};

template <int NW>
__device__ __forceinline__ void entry_mismatches(const EmitArgs &a, u64 e, u64 (&rw)[NW], u64 (&rn)[NW], u64 (&cw)[NW], u64 (&mm)[NW])
{
	const u32 src = a.f_src[e];
	const u8 kind = a.f_kind[e];
	if (kind == 0) {
		load_words<NW>(a.sreads + (size_t)src * NW, rw);
#pragma unroll
		for (int k = 0; k < NW; k++) rn[k] = 0;
	} else {
		load_words<NW>(a.pool + (size_t)src * NW, rw);
		load_words<NW>(a.poolN + (size_t)src * NW, rn);
		if (kind & 2) revcomp_pool<NW>(rw, rn, a.L); // stored as reverse_complement(read) (encoder.cpp:375)
	}
	load_window<NW>(a.cons2, a.f_col[e], a.L, cw);
#pragma unroll
	for (int k = 0; k < NW; k++) {
		u64 x = rw[k] ^ cw[k];
		mm[k] = ((x | (x >> 1)) & 0x5555555555555555ull) | rn[k];
	}
}

template <int NW>
__global__ void __launch_bounds__(128) emit_count_kernel(EmitArgs a)
{
	const u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= a.F) return;
	u64 rw[NW], rn[NW], cw[NW], mm[NW];
	entry_mismatches<NW>(a, e, rw, rn, cw, mm);
	int nm = 0;
#pragma unroll
	for (int k = 0; k < NW; k++) nm += __popcll(mm[k]);
	a.nm1[e] = (u64)nm + 1;
	const u32 src = a.f_src[e];
	const u8 kind = a.f_kind[e];
	// first read of a contig: readlen; otherwise delta to the previous read of the merged list (661-663, 682-683, 707-708)
	a.posb[e] = (kind == 0 && a.cs[src]) ? (u8)a.L : (u8)(a.f_col[e] - a.f_col[e - 1]);
	a.revc[e] = kind == 0 ? a.s_rev[src] : ((kind & 2) ? 'r' : 'd');
	a.isN[e] = (kind != 0 && src >= a.n_s) ? 1u : 0u; // only reads from input_N.dna contain 'N' (684-687)
	a.ordv[e] = kind == 0 ? a.s_order[src] : a.pool_order[src];
}

// enc_noise[ref][read] (encoder.cpp:752-771), indexed by the 2-bit code A0 G1 C2 T3 and N=4
__device__ __constant__ char kNoise[4][5] = {
	{ 0, '1', '0', '2', '3' }, // ref A: C0 G1 T2 N3
	{ '1', 0, '2', '0', '3' }, // ref G: T0 A1 C2 N3
	{ '0', '1', 0, '2', '3' }, // ref C: A0 G1 T2 N3
	{ '2', '0', '1', 0, '3' }, // ref T: G0 C1 A2 N3
};

template <int NW>
__global__ void __launch_bounds__(128) emit_write_kernel(EmitArgs a)
{
	const u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= a.F) return;
	u64 rw[NW], rn[NW], cw[NW], mm[NW];
	entry_mismatches<NW>(a, e, rw, rn, cw, mm);
	const u64 no = a.noff[e];
	char *np = a.noise + no;
	u8 *pp = a.noisepos + (no - e);
	int t = 0, prev = 0;
#pragma unroll
	for (int k = 0; k < NW; k++) {
		u64 mk = mm[k];
		while (mk) {
			int b = __ffsll((long long)mk) - 1;
			mk &= mk - 1;
			int p = 32 * k + (b >> 1);
			u32 rf = (u32)(cw[k] >> b) & 3u;
			u32 rd = ((rn[k] >> b) & 1ull) ? 4u : ((u32)(rw[k] >> b) & 3u);
			np[t] = kNoise[rf][rd];
			pp[t] = (u8)(p - prev); // delta from the previous noise position of this read, first one absolute (677-679)
			prev = p;
			t++;
		}
	}
	np[t] = '\n';
	if (a.isN[e]) a.o_order_N[a.exN[e]] = a.ordv[e];
	else a.o_order[e - a.exN[e]] = a.ordv[e];
}

// ---------------------------------------------------------------------------------------------- packbits (512-616)
// seq: 4 bases per byte, base k of a group at bits 2k, A0 C1 G2 T3
__global__ void __launch_bounds__(256) pack_seq_kernel(const u64 *__restrict__ cons2, u64 col0, u64 nbytes, u8 *__restrict__ out)
{
	u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nbytes) return;
	out[b] = (u8)swappairs64(getbits_g(cons2, 2 * (col0 + 4 * b), 8));
}
// rev: 8 flags per byte, LSB first, d=0 r=1
__global__ void __launch_bounds__(256) pack_rev_kernel(const u8 *__restrict__ revc, u64 e0, u64 nbytes, u8 *__restrict__ out)
{
	u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nbytes) return;
	u32 v = 0;
#pragma unroll
	for (int k = 0; k < 8; k++) v |= (revc[e0 + 8 * b + k] == 'r' ? 1u : 0u) << k;
	out[b] = (u8)v;
}
// the <4 leftover bases / <8 leftover flags go to the .tail files as ASCII
__global__ void tails_kernel(const u64 *__restrict__ cons2, u64 colend, u32 nseq, const u8 *__restrict__ revc, u64 eend, u32 nrev, char *__restrict__ out)
{
	if (threadIdx.x < nseq) out[threadIdx.x] = "AGCT"[getbits_g(cons2, 2 * (colend - nseq + threadIdx.x), 2)];
	if (threadIdx.x < nrev) out[4 + threadIdx.x] = (char)revc[eend - nrev + threadIdx.x];
}

// ---------------------------------------------------------------------------------------------- unaligned pool reads (476-499)
__global__ void __launch_bounds__(256) unaligned_flag_kernel(const u64 *__restrict__ best, u32 P, u32 *__restrict__ uf)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < P) uf[i] = best[i] == NOBEST;
}
__global__ void __launch_bounds__(256) unaligned_kernel(const u32 *__restrict__ uf, const u32 *__restrict__ exU, u32 P, u32 n_s, u32 U_s,
                                                        const u32 *__restrict__ pool_order, u32 *__restrict__ ulist,
                                                        u32 *__restrict__ o_order_tail, u32 *__restrict__ o_order_N_tail)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P || !uf[i]) return;
	u32 u = exU[i];
	ulist[u] = i;
	if (i < n_s) o_order_tail[u] = pool_order[i];
	else o_order_N_tail[u - U_s] = pool_order[i];
}
__global__ void __launch_bounds__(256) pack_singleton_kernel(const u64 *__restrict__ pool, const u32 *__restrict__ ulist, int L, int NW,
                                                             u64 nbases, u64 nbytes, u8 *__restrict__ out, char *__restrict__ tail)
{
	u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (b > nbytes) return;
	u32 v = 0;
	for (int k = 0; k < 4; k++) {
		u64 t = 4 * b + k;
		if (t >= nbases) break;
		u32 rid = ulist[t / L];
		int off = (int)(t % L);
		u32 c = (u32)(__ldg(&pool[(size_t)rid * NW + (off >> 5)]) >> (2 * (off & 31))) & 3u;
		if (b == nbytes) tail[k] = "AGCT"[c];
		v |= (((c & 1u) << 1) | (c >> 1)) << (2 * k);
	}
	if (b < nbytes) out[b] = (u8)v;
}
__global__ void __launch_bounds__(256) unaligned_N_kernel(const u64 *__restrict__ pool, const u64 *__restrict__ poolN, const u32 *__restrict__ ulist,
                                                          int L, int NW, u64 nbytes, char *__restrict__ out)
{
	u64 b = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nbytes) return;
	u64 line = b / (L + 1);
	int c = (int)(b % (L + 1));
	char ch = '\n';
	if (c < L) {
		u32 rid = ulist[line];
		u64 w2 = __ldg(&pool[(size_t)rid * NW + (c >> 5)]), wn = __ldg(&poolN[(size_t)rid * NW + (c >> 5)]);
		ch = ((wn >> (2 * (c & 31))) & 1ull) ? 'N' : "AGCT"[(w2 >> (2 * (c & 31))) & 3ull];
	}
	out[b] = ch;
}

// ---------------------------------------------------------------------------------------------- pack_order (pack_order.cpp:20-77)
// Blocks of 32 entries, `numbits` bits each, as one little-endian bit stream of `numbits` u32 words per block.
// One thread per output word: word k of a block holds bits [32k, 32k+32) of the block's stream.
__global__ void __launch_bounds__(256) pack_order_kernel(const u32 *__restrict__ order, u64 nwords, int numbits, u32 *__restrict__ out)
{
	const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nwords) return;
	const u64 blk = t / numbits;
	const int k = (int)(t % numbits);
	const u32 *o = order + blk * 32;
	const int b0 = 32 * k, b1 = b0 + 32; // bit range of this word
	u32 v = 0;
	for (int j = b0 / numbits; j < 32 && j * numbits < b1; j++) {
		const int sh = j * numbits - b0;
		const u32 x = __ldg(&o[j]);
		v |= sh >= 0 ? x << sh : x >> (-sh);
	}
	out[t] = v;
}

// ---------------------------------------------------------------------------------------------- hand-off from stage I
template <int NW>
__global__ void __launch_bounds__(128) gather_stream_kernel(const u64 *__restrict__ reads, const u32 *__restrict__ order, const u8 *__restrict__ rev,
                                                            u32 cnt, int L, u64 *__restrict__ out)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= cnt) return;
	u64 r[NW];
	load_words<NW>(reads + (size_t)order[i] * NW, r);
	if (rev && rev[i] == 'r') { // temp.dna holds the read already reverse-complemented (reorder.cpp:744-752)
		u64 t[NW];
		reverse2<NW>(r, L, t);
#pragma unroll
		for (int k = 0; k < NW; k++) r[k] = t[k] ^ lowmask(2 * L - 64 * k);
	}
#pragma unroll
	for (int k = 0; k < NW; k++) out[(size_t)i * NW + k] = r[k];
}
__global__ void __launch_bounds__(256) iota_off_kernel(u32 *v, u32 n)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) v[i] = i;
}
} // namespace

#define DISPATCH_NW(NWv, CALL)                                                                    \
	switch (NWv) {                                                                                \
	case 1: { constexpr int NW = 1; CALL; } break;                                                \
	case 2: { constexpr int NW = 2; CALL; } break;                                                \
	case 3: { constexpr int NW = 3; CALL; } break;                                                \
	case 4: { constexpr int NW = 4; CALL; } break;                                                \
	case 5: { constexpr int NW = 5; CALL; } break;                                                \
	case 6: { constexpr int NW = 6; CALL; } break;                                                \
	case 7: { constexpr int NW = 7; CALL; } break;                                                \
	case 8: { constexpr int NW = 8; CALL; } break;                                                \
	default: harcgpu_set_error("unsupported read length"); return -1;                              \
	}

static void free_stream(harcgpu_ctx *c)
{
	c->release(c->sreads); c->release(c->s_order); c->release(c->s_rev); c->release(c->s_flag); c->release(c->s_pos);
	c->sreads = nullptr; c->s_order = nullptr; c->s_rev = c->s_flag = c->s_pos = nullptr;
	c->stream_set = false; c->encoded = false;
}
static void free_pool(harcgpu_ctx *c)
{
	c->release(c->pool); c->release(c->poolN); c->release(c->pool_order);
	c->pool = c->poolN = nullptr; c->pool_order = nullptr;
	for (int l = 0; l < 2; l++) free_dict(c, c->d2[l]);
	c->release(c->bloom2);
	c->bloom2 = nullptr;
	c->pool_set = false; c->encoded = false;
}

int s2_set_stream_from_stage1(harcgpu_ctx *c)
{
	free_stream(c);
	cudaStream_t st = c->st;
	const u32 m = c->n_matched;
	c->m = m;
	if (c->alloc(&c->sreads, (size_t)m * c->NW) || c->alloc(&c->s_order, m) || c->alloc(&c->s_rev, m) || c->alloc(&c->s_flag, m) ||
	    c->alloc(&c->s_pos, m))
		return -1;
	if (m) {
		DISPATCH_NW(c->NW, (gather_stream_kernel<NW><<<KL + cdiv(m, 128), 128, 0, st>>>(c->reads, c->order, c->rev, m, c->L, c->sreads)));
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(c->s_order, c->order, 4 * (size_t)m, cudaMemcpyDeviceToDevice, st));
		CK(cudaMemcpyAsync(c->s_rev, c->rev, m, cudaMemcpyDeviceToDevice, st));
		CK(cudaMemcpyAsync(c->s_flag, c->flag, m, cudaMemcpyDeviceToDevice, st));
		CK(cudaMemcpyAsync(c->s_pos, c->pos, m, cudaMemcpyDeviceToDevice, st));
	}
	c->stream_set = true;
	return 0;
}

int s2_set_stream_host(harcgpu_ctx *c, const char *dna, const char *flag, const u8 *pos, const u32 *order, const char *rev, u32 m)
{
	free_stream(c);
	cudaStream_t st = c->st;
	c->m = m;
	if (c->alloc(&c->sreads, (size_t)m * c->NW) || c->alloc(&c->s_order, m) || c->alloc(&c->s_rev, m) || c->alloc(&c->s_flag, m) ||
	    c->alloc(&c->s_pos, m))
		return -1;
	if (m) {
		char *d = nullptr;
		size_t bytes = (size_t)m * (c->L + 1);
		if (c->alloc(&d, bytes + 16)) return -1;
		CK(cudaMemcpyAsync(d, dna, bytes, cudaMemcpyHostToDevice, st));
		if (s1_packN(c, d, m, c->sreads, nullptr)) return -1;
		CK(cudaMemcpyAsync(c->s_order, order, 4 * (size_t)m, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(c->s_rev, rev, m, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(c->s_flag, flag, m, cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(c->s_pos, pos, m, cudaMemcpyHostToDevice, st));
		CK(cudaStreamSynchronize(st));
		c->release(d);
	}
	c->stream_set = true;
	return 0;
}

// encoder.cpp:823-872 + 132-148.  d_N: device buffer with the N reads (16-byte aligned) or null; h_N: host buffer or null.
// by_id: the singletons are given as ids into the reads resident on this context (order_s on the host, s_ascii null)
static int load_pool_impl(harcgpu_ctx *c, const char *s_ascii, const u32 *order_s, u32 n_s, const char *h_N, const void *d_N, u32 n_N,
                          bool by_id = false)
{
	free_pool(c);
	cudaStream_t st = c->st;
	const bool from_stage1 = !by_id && (s_ascii == nullptr && order_s == nullptr && n_s == 0 && c->reordered);
	if (from_stage1) n_s = c->n_single;
	else if (by_id) { if (n_s && (!order_s || !c->reads)) { harcgpu_set_error("singleton ids need the reads of this context"); return -1; } }
	else if (n_s && (!s_ascii || !order_s)) { harcgpu_set_error("singleton reads and their order are both needed"); return -1; }
	const u64 P64 = (u64)n_s + n_N;
	if (P64 > 0xfffffff0ull) { harcgpu_set_error("pool too large"); return -1; }
	const u32 P = (u32)P64;
	c->n_s = n_s; c->n_N = n_N;
	if (c->alloc(&c->pool, (size_t)P * c->NW) || c->alloc(&c->poolN, (size_t)P * c->NW) || c->alloc(&c->pool_order, P)) return -1;
	if (P) CK(cudaMemsetAsync(c->poolN, 0, (size_t)P * c->NW * 8, st));
	const size_t line = (size_t)c->L + 1;
	if (n_s && by_id) {
		CK(cudaMemcpyAsync(c->pool_order, order_s, 4 * (size_t)n_s, cudaMemcpyDefault, st)); // host or device ids
		DISPATCH_NW(c->NW, (gather_stream_kernel<NW><<<KL + cdiv(n_s, 128), 128, 0, st>>>(c->reads, c->pool_order, nullptr, n_s, c->L, c->pool)));
		CK(cudaGetLastError());
		CK(cudaStreamSynchronize(st));
	} else if (n_s) {
		if (from_stage1) {
			DISPATCH_NW(c->NW, (gather_stream_kernel<NW><<<KL + cdiv(n_s, 128), 128, 0, st>>>(c->reads, c->order_s, nullptr, n_s, c->L, c->pool)));
			CK(cudaGetLastError());
			CK(cudaMemcpyAsync(c->pool_order, c->order_s, 4 * (size_t)n_s, cudaMemcpyDeviceToDevice, st));
		} else {
			char *d = nullptr;
			if (c->alloc(&d, n_s * line + 16)) return -1;
			CK(cudaMemcpyAsync(d, s_ascii, n_s * line, cudaMemcpyHostToDevice, st));
			if (s1_packN(c, d, n_s, c->pool, nullptr)) return -1;
			CK(cudaMemcpyAsync(c->pool_order, order_s, 4 * (size_t)n_s, cudaMemcpyHostToDevice, st));
			CK(cudaStreamSynchronize(st));
			c->release(d);
		}
	}
	if (n_N) {
		char *d = nullptr;
		if (!d_N && c->staged_N && c->staged_host == h_N && c->staged_n == n_N) {
			// uploaded ahead of time by harcgpu_stage_nreads: only wait for that copy
			CK(cudaStreamWaitEvent(st, c->ev_staged, 0));
			d = c->staged_N;
			c->staged_N = nullptr; c->staged_host = nullptr; c->staged_n = 0;
		} else if (!d_N) {
			if (c->alloc(&d, n_N * line + 16)) return -1;
			CK(cudaMemcpyAsync(d, h_N, n_N * line, cudaMemcpyDefault, st)); // host (or, for the bench, device) buffer
		}
		if (s1_packN(c, d_N ? d_N : d, n_N, c->pool + (size_t)n_s * c->NW, c->poolN + (size_t)n_s * c->NW)) return -1;
		iota_off_kernel<<<KL + cdiv(n_N, 256), 256, 0, st>>>(c->pool_order + n_s, n_N); // order_s[i] = i - numreads_s (869-870)
		CK(cudaGetLastError());
		if (d) { CK(cudaStreamSynchronize(st)); c->release(d); }
	}
	// dictionary windows of encoder.cpp:132-145
	int ds[2], de[2];
	const int L = c->L;
	if (L > 50) { ds[0] = 0; de[0] = 20; ds[1] = 21; de[1] = 41; }
	else { ds[0] = 0; de[0] = 20 * L / 50; ds[1] = 20 * L / 50 + 1; de[1] = 41 * L / 50; }
	c->tic();
	for (int l = 0; l < 2; l++)
		if (build_dict(c, c->d2[l], c->pool, c->poolN, P, c->NW, ds[l], de[l], 3)) return -1;
	{
		// Bloom filter over both dictionaries: >= 8 bits per key (it has to stay in L2 next to the streams of the kernel)
		const u64 nk = (u64)c->d2[0].numkeys + c->d2[1].numkeys;
		u64 words = 1024;
		while (words * 32 < 8 * nk) words <<= 1;
		c->release(c->bloom2);
		c->bloom2 = nullptr;
		if (c->alloc(&c->bloom2, words)) return -1;
		c->bloom2_mask = (u32)(words - 1);
		CK(cudaMemsetAsync(c->bloom2, 0, words * 4, st));
		for (int l = 0; l < 2; l++)
			if (c->d2[l].numkeys) bloom_insert_kernel<<<KL + cdiv(c->d2[l].numkeys, 256), 256, 0, st>>>(c->d2[l].keys, c->d2[l].numkeys, l, c->bloom2, c->bloom2_mask);
		CK(cudaGetLastError());
	}
	c->toc("pooldict");
	c->pool_set = true;
	return 0;
}
int s2_load_pool(harcgpu_ctx *c, const char *s_ascii, const u32 *order_s, u32 n_s, const char *N_ascii, u32 n_N)
{
	return load_pool_impl(c, s_ascii, order_s, n_s, N_ascii, nullptr, n_N);
}
int s2_load_pool_ids(harcgpu_ctx *c, const u32 *ids, u32 n_s, const char *N_ascii, u32 n_N)
{
	return load_pool_impl(c, nullptr, ids, n_s, N_ascii, nullptr, n_N, true);
}
int s2_load_pool_dev(harcgpu_ctx *c, const void *d_N_ascii, u32 n_N) { return load_pool_impl(c, nullptr, nullptr, 0, nullptr, d_N_ascii, n_N); }

int s2_encode(harcgpu_ctx *c)
{
	cudaStream_t st = c->st;
	const u32 m = c->m, P = c->n_s + c->n_N, n_s = c->n_s;
	const int L = c->L, NWv = c->NW, K = c->p.file_sets;
	c->encoded = false;
	// drop previous outputs
	for (auto &s : c->sets) { c->release(s.seq); c->release(s.rev); }
	c->sets.clear();
	c->release(c->o_order); c->release(c->o_order_N); c->release(c->o_single); c->release(c->o_inputN);
	c->o_order = c->o_order_N = nullptr; c->o_single = nullptr; c->o_inputN = nullptr;
	for (void *q : c->s2_keep) c->release(q);
	c->s2_keep.clear();
	c->tic();
	c->lap(nullptr);

	const u32 per = m ? 1 + (m - 1) / K : 1; // encoder.cpp:171
	u32 *ns = nullptr, *ex = nullptr, *nat_idx = nullptr, *cs = nullptr, *cid = nullptr, *cstart = nullptr, *d_tot32 = nullptr;
	u64 *inc = nullptr, *G = nullptr, *scan_tmp = nullptr, *d_tot64 = nullptr, *cons2 = nullptr;
	u32 *tile_idx = nullptr;
	u32 NC = 0, nt_host = 0;
	u64 TOT = 0;
	size_t scan_n = std::max<size_t>(std::max<size_t>(m, P), 1);
	if (c->alloc(&d_tot32, 4) || c->alloc(&d_tot64, 2)) return -1;
	if (m) {
		if (c->alloc(&ns, m) || c->alloc(&ex, m) || c->alloc(&nat_idx, m) || c->alloc(&cs, m) || c->alloc(&cid, m) || c->alloc(&cstart, m) ||
		    c->alloc(&inc, m) || c->alloc(&G, m))
			return -1;
	}
	if (c->alloc(&scan_tmp, scan_tmp_elems(2 * scan_n + 64))) return -1;
	if (m) {
		natstart_kernel<<<KL + cdiv(m, 256), 256, 0, st>>>(c->s_flag, m, per, ns);
		if (exclusive_scan_u32(ns, ex, m, scan_tmp, nullptr, st)) return -1;
		scatter_idx_kernel<<<KL + cdiv(m, 256), 256, 0, st>>>(ns, ex, m, nat_idx);
		cstart_kernel<<<KL + cdiv(m, 256), 256, 0, st>>>(ns, ex, nat_idx, c->s_pos, m, L, cs, inc);
		if (exclusive_scan_u32(cs, ex, m, scan_tmp, d_tot32, st)) return -1;
		if (exclusive_scan_u64(inc, G, m, scan_tmp, d_tot64, st)) return -1;
		finish_layout_kernel<<<KL + cdiv(m, 256), 256, 0, st>>>(cs, ex, inc, G, m, cid, cstart);
		CK(cudaGetLastError());
		u64 tot = 0;
		CK(cudaMemcpyAsync(&NC, d_tot32, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&tot, d_tot64, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		TOT = tot + L;
	}
	c->lap("s2_layout");
	const size_t cwords = (size_t)((TOT + 31) / 32);
	if (c->alloc(&cons2, cwords + 2)) return -1;
	CK(cudaMemsetAsync(cons2 + cwords, 0, 16, st));
	if (m) {
		// tile index: entries 0 .. TOT/TILE_C + CONS_T/TILE_C (the consensus tiles look one tile past their end)
		const u32 nt = (u32)(TOT / TILE_C + CONS_T / TILE_C + 2);
		nt_host = nt;
		if (c->alloc(&tile_idx, nt)) return -1;
		tile_index_kernel<<<KL + cdiv(nt, 256), 256, 0, st>>>(G, m, nt, tile_idx);
		// bit-sliced vote; the column-per-lane kernel only if some column is covered by more reads than the sliced counters hold
		u32 h_deep = 0;
		CK(cudaMemsetAsync(d_tot32, 0, 4, st));
		consensus_sliced_kernel<<<KL + cdiv(TOT, CS_BLOCK_COLS), CS_THREADS, 0, st>>>(G, tile_idx, nt, reinterpret_cast<const u32 *>(c->sreads), m, L,
		                                                                            2 * NWv, TOT, reinterpret_cast<u32 *>(cons2), d_tot32);
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(&h_deep, d_tot32, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		if (h_deep || getenv("HARCGPU_CONSENSUS_COLUMNS")) {
			const size_t csm = (size_t)CONS_R * (1 + 2 * NWv) * sizeof(u32);
			consensus_kernel<<<KL + cdiv(TOT, CONS_T), CONS_T, csm, st>>>(G, tile_idx, reinterpret_cast<const u32 *>(c->sreads), m, L, 2 * NWv, TOT, cons2);
			CK(cudaGetLastError());
		}
	}

	c->lap("s2_consensus");
	// ---- pool re-alignment
	u64 *best = nullptr, *prio_u = nullptr, *iprio = nullptr;
	u32 *af = nullptr, *exa = nullptr, *rid_u = nullptr, *irid = nullptr;
	u32 M = 0, M_single = 0;
	u32 *pflags = nullptr;
	unsigned long long *pcur = nullptr;
	constexpr u32 PCUR_ENTRIES = 1u << 16;
	PoolArgs pa;
	bool have_probe = false;
	if (c->alloc(&best, (size_t)P + 1) || c->alloc(&pflags, 2) || c->alloc(&pcur, PCUR_ENTRIES)) return -1;
	CK(cudaMemsetAsync(pcur, 0xff, 8 * (size_t)PCUR_ENTRIES, st)); // no bin has this tag
	if (P) {
		fill64_kernel<<<KL + cdiv(P, 256), 256, 0, st>>>(best, P, NOBEST);
		CK(cudaGetLastError());
	}
	if (P && m && TOT >= (u64)L) {
		PoolArgs a;
		a.G = G; a.cid = cid; a.cstart = cstart; a.m = m; a.NC = NC; a.per = per; a.TOT = TOT; a.cons2 = cons2;
		a.pool = c->pool; a.poolN = c->poolN;
		for (int l = 0; l < 2; l++) {
			a.d[l].slots = c->d2[l].slots; a.d[l].ids = c->d2[l].ids; a.d[l].slot_shift = c->d2[l].slot_shift;
			a.d[l].dstart = c->d2[l].bitpos / 3; a.d[l].dend = a.d[l].dstart + c->d2[l].nbits / 3 - 1;
			a.d[l].world = 0;
		}
		a.L = L; a.thresh_s = c->p.thresh_s; a.maxsearch = c->p.maxsearch; a.best = best; a.flags = pflags; a.pcur = pcur; a.pcur_mask = PCUR_ENTRIES - 1;
		a.rank_bits = (u64)(c->shard_world > 1 ? c->shard_rank : 0) << RANK_SHIFT;
		a.bloom = c->bloom2; a.bloom_mask = c->bloom2_mask; a.T = tile_idx; a.nt = nt_host; a.cwords = cwords + 2;
		pa = a;
		have_probe = true;
	}
	// One pass of the probe settles everything unless a bin beyond maxsearch was cut short; then the probe is repeated until
	// no priority moves (see pool_probe_kernel).  One job on several GPUs: every GPU probed the same pool against its own
	// contigs and the smallest priority over all GPUs wins -- all-reduce(min) after every pass, done by the caller's exchange
	// hook (NCCL through torch.distributed in this repo); the word behind the priorities carries "somebody goes on".
	for (int pass = 0; P && (have_probe || c->shard_world > 1); pass++) {
		CK(cudaMemsetAsync(pflags, 0, 8, st));
		if (have_probe) {
			const u64 nwin = TOT - L + 1;
			DISPATCH_NW(NWv, (pool_probe_kernel<NW><<<KL + cdiv(nwin, PP_COLS), PP_THREADS, 0, st>>>(pa)));
			CK(cudaGetLastError());
		}
		u32 hf[2] = { 0, 0 };
		CK(cudaMemcpyAsync(hf, pflags, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		// (the first pass also lowers priorities through the small bins, which the flag does not see: a cut-short scan always
		// gets a second look)
		bool again = hf[0] && (hf[1] || pass == 0);
		if (c->shard_world > 1) {
			if (!c->pool_exchange) { harcgpu_set_error("encode of one job on several GPUs needs harcgpu_set_pool_exchange"); return -1; }
			const u64 word = again ? 0ull : NOBEST;
			CK(cudaMemcpyAsync(best + P, &word, 8, cudaMemcpyHostToDevice, st));
			CK(cudaStreamSynchronize(st));
			if (c->pool_exchange(c->pool_exchange_user, best, (u64)P + 1)) { harcgpu_set_error("pool exchange hook failed"); return -1; }
			u64 got = 0;
			CK(cudaMemcpyAsync(&got, best + P, 8, cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			again = got == 0ull;
		}
		c->ms["pool_passes"] = pass + 1;
		if (!again) break;
		if (pass >= 15) break; // sixteen passes: what is settled by then is a valid (lossless) assignment, if not the reference's
	}
	if (P && c->shard_world > 1) {
		best_localize_kernel<<<KL + cdiv(P, 256), 256, 0, st>>>(best, P, c->shard_rank);
		CK(cudaGetLastError());
	}
	if (P && ((m && TOT >= (u64)L) || c->shard_world > 1)) {
		if (c->alloc(&af, P) || c->alloc(&exa, P) || c->alloc(&prio_u, P) || c->alloc(&rid_u, P)) return -1;
		aligned_flag_kernel<<<KL + cdiv(P, 256), 256, 0, st>>>(best, P, af);
		if (exclusive_scan_u32(af, exa, P, scan_tmp, d_tot32, st)) return -1;
		aligned_compact_kernel<<<KL + cdiv(P, 256), 256, 0, st>>>(best, af, exa, P, prio_u, rid_u);
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(&M, d_tot32, 4, cudaMemcpyDeviceToHost, st));
		u32 before = 0; // aligned reads among the ids >= n_s (the flags run from the highest id down)
		if (n_s && n_s < P) CK(cudaMemcpyAsync(&before, exa + (P - n_s), 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		M_single = n_s ? M - (n_s < P ? before : 0u) : 0u;
	}
	if (c->alloc(&iprio, M) || c->alloc(&irid, M)) return -1;
	if (M) {
		int end_bit = 3;
		while (end_bit < 64 && (TOT >> (end_bit - 2)) != 0) end_bit++;
		if (radix_sort_pairs(c, &prio_u, &iprio, &rid_u, &irid, M, 0, end_bit)) return -1;
		std::swap(prio_u, iprio); // iprio / irid = sorted
		std::swap(rid_u, irid);
		CK(cudaStreamSynchronize(st));
	}

	c->lap("s2_pool");
	// ---- merged list
	const u64 F = (u64)m + M;
	u32 *f_src = nullptr, *isN = nullptr, *exN = nullptr, *ordv = nullptr;
	u8 *f_kind = nullptr, *posb = nullptr, *revc = nullptr, *noisepos = nullptr;
	u64 *f_col = nullptr, *nm1 = nullptr, *noff = nullptr;
	char *noise = nullptr;
	if (c->alloc(&f_src, F) || c->alloc(&f_kind, F) || c->alloc(&f_col, F) || c->alloc(&nm1, F) || c->alloc(&noff, F + 1) ||
	    c->alloc(&posb, F) || c->alloc(&revc, F + 8) || c->alloc(&isN, F) || c->alloc(&exN, F) || c->alloc(&ordv, F))
		return -1;
	c->release(scan_tmp);
	if (c->alloc(&scan_tmp, scan_tmp_elems(2 * std::max<u64>(std::max<u64>(F, P), 1) + 64))) return -1;
	u64 NB = 0;
	u32 FN = 0;
	if (m) {
		place_orig_kernel<<<KL + cdiv(m, 256), 256, 0, st>>>(G, m, iprio, M, f_src, f_kind, f_col);
		CK(cudaGetLastError());
	}
	if (M) {
		place_ins_kernel<<<KL + cdiv(M, 256), 256, 0, st>>>(G, m, iprio, irid, M, f_src, f_kind, f_col);
		CK(cudaGetLastError());
	}
	EmitArgs ea;
	ea.f_src = f_src; ea.f_kind = f_kind; ea.f_col = f_col; ea.F = F; ea.sreads = c->sreads; ea.cs = cs; ea.s_order = c->s_order;
	ea.s_rev = c->s_rev; ea.pool = c->pool; ea.poolN = c->poolN; ea.pool_order = c->pool_order; ea.n_s = n_s; ea.cons2 = cons2; ea.L = L;
	ea.nm1 = nm1; ea.posb = posb; ea.revc = revc; ea.isN = isN; ea.ordv = ordv; ea.noff = noff; ea.exN = exN;
	ea.noise = nullptr; ea.noisepos = nullptr; ea.o_order = nullptr; ea.o_order_N = nullptr;
	if (F) {
		DISPATCH_NW(NWv, (emit_count_kernel<NW><<<KL + cdiv(F, 128), 128, 0, st>>>(ea)));
		CK(cudaGetLastError());
		if (exclusive_scan_u64(nm1, noff, F, scan_tmp, d_tot64, st)) return -1;
		if (exclusive_scan_u32(isN, exN, F, scan_tmp, d_tot32, st)) return -1;
		CK(cudaMemcpyAsync(&NB, d_tot64, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&FN, d_tot32, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		CK(cudaMemcpyAsync(noff + F, d_tot64, 8, cudaMemcpyDeviceToDevice, st));
	} else {
		CK(cudaMemsetAsync(noff, 0, 8, st));
	}
	c->lap("s2_merge_emit");
	// ---- unaligned pool reads
	u32 *uf = nullptr, *exU = nullptr, *ulist = nullptr;
	u32 U = 0, U_s = 0;
	if (c->alloc(&uf, P) || c->alloc(&exU, (size_t)P + 1) || c->alloc(&ulist, P)) return -1;
	if (P) {
		unaligned_flag_kernel<<<KL + cdiv(P, 256), 256, 0, st>>>(best, P, uf);
		if (exclusive_scan_u32(uf, exU, P, scan_tmp, d_tot32, st)) return -1;
		CK(cudaMemcpyAsync(exU + P, d_tot32, 4, cudaMemcpyDeviceToDevice, st));
		CK(cudaMemcpyAsync(&U, d_tot32, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&U_s, exU + n_s, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	}
	const u32 U_N = U - U_s;
	const u64 Fc = F - FN;
	const u64 n_order = Fc + U_s, n_order_N = (u64)FN + U_N;
	if (c->alloc(&c->o_order, n_order) || c->alloc(&c->o_order_N, n_order_N) || c->alloc(&noise, NB) || c->alloc(&noisepos, NB - F + 1))
		return -1;
	if (F) {
		ea.noise = noise; ea.noisepos = noisepos; ea.o_order = c->o_order; ea.o_order_N = c->o_order_N;
		DISPATCH_NW(NWv, (emit_write_kernel<NW><<<KL + cdiv(F, 128), 128, 0, st>>>(ea)));
		CK(cudaGetLastError());
	}
	const u64 sbases = (u64)U_s * L, sbytes = sbases / 4, stail = sbases % 4;
	const u64 nbytesN = (u64)U_N * (L + 1);
	char *d_tail = nullptr;
	if (c->alloc(&c->o_single, sbytes) || c->alloc(&c->o_inputN, nbytesN) || c->alloc(&d_tail, 16 * (size_t)(K + 1))) return -1;
	CK(cudaMemsetAsync(d_tail, 0, 16 * (size_t)(K + 1), st));
	if (P) {
		unaligned_kernel<<<KL + cdiv(P, 256), 256, 0, st>>>(uf, exU, P, n_s, U_s, c->pool_order, ulist, c->o_order + Fc, c->o_order_N + FN);
		CK(cudaGetLastError());
		if (sbases) {
			pack_singleton_kernel<<<KL + cdiv(sbytes + 1, 256), 256, 0, st>>>(c->pool, ulist, L, NWv, sbases, sbytes, c->o_single, d_tail + 16 * (size_t)K);
			CK(cudaGetLastError());
		}
		if (nbytesN) {
			unaligned_N_kernel<<<KL + cdiv(nbytesN, 256), 256, 0, st>>>(c->pool, c->poolN, ulist + U_s, L, NWv, nbytesN, c->o_inputN);
			CK(cudaGetLastError());
		}
	}

	c->lap("s2_unaligned");
	// ---- per file set views (encoder.cpp:169-196, 512-581)
	c->sets.resize(K);
	std::vector<u64> h_fs(K + 1), h_col(K + 1), h_no(K + 1);
	for (int k = 0; k <= K; k++) {
		u64 a = std::min<u64>((u64)k * per, m);
		if (k == K || a >= m) { h_fs[k] = F; h_col[k] = TOT; h_no[k] = NB; continue; }
		u64 g = 0;
		CK(cudaMemcpyAsync(&g, G + a, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		h_col[k] = g;
		// inserted reads with column < g precede original a; none can sit at column >= g of an earlier contig
		u64 lo = 0;
		if (M) {
			// binary search on the device array through small copies
			u64 l0 = 0, hi = M;
			while (l0 < hi) {
				u64 mid = (l0 + hi) >> 1, pv = 0;
				CK(cudaMemcpyAsync(&pv, iprio + mid, 8, cudaMemcpyDeviceToHost, st));
				CK(cudaStreamSynchronize(st));
				if ((pv >> 2) < g) l0 = mid + 1; else hi = mid;
			}
			lo = l0;
		}
		h_fs[k] = a + lo;
		CK(cudaMemcpyAsync(&h_no[k], noff + h_fs[k], 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	}
	for (int k = 0; k < K; k++) {
		SetOut &s = c->sets[k];
		const u64 e0 = h_fs[k], e1 = h_fs[k + 1], cnt = e1 - e0;
		const u64 col0 = h_col[k], col1 = h_col[k + 1], ncol = col1 - col0;
		s.pos = posb + e0; s.pos_bytes = cnt;
		s.noise = noise + h_no[k]; s.noise_bytes = h_no[k + 1] - h_no[k];
		s.noisepos = noisepos + (h_no[k] - e0); s.noisepos_bytes = s.noise_bytes - cnt;
		s.seq_bytes = ncol / 4; s.seq_ntail = (u32)(ncol % 4);
		s.rev_bytes = cnt / 8; s.rev_ntail = (u32)(cnt % 8);
		if (c->alloc(&s.seq, s.seq_bytes) || c->alloc(&s.rev, s.rev_bytes)) return -1;
		if (s.seq_bytes) pack_seq_kernel<<<KL + cdiv(s.seq_bytes, 256), 256, 0, st>>>(cons2, col0, s.seq_bytes, s.seq);
		if (s.rev_bytes) pack_rev_kernel<<<KL + cdiv(s.rev_bytes, 256), 256, 0, st>>>(revc, e0, s.rev_bytes, s.rev);
		if (s.seq_ntail || s.rev_ntail) tails_kernel<<<KL + 1, 32, 0, st>>>(cons2, col1, s.seq_ntail, revc, e1, s.rev_ntail, d_tail + 16 * (size_t)k);
		CK(cudaGetLastError());
	}
	std::vector<char> h_tail(16 * (size_t)(K + 1));
	CK(cudaMemcpyAsync(h_tail.data(), d_tail, h_tail.size(), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	for (int k = 0; k < K; k++) {
		memcpy(c->sets[k].seq_tail, &h_tail[16 * (size_t)k], 4);
		memcpy(c->sets[k].rev_tail, &h_tail[16 * (size_t)k + 4], 8);
	}
	memcpy(c->single_tail, &h_tail[16 * (size_t)K], 4);
	c->lap("s2_sets");
	c->toc("encode");

	c->esz.n_order = (u32)n_order; c->esz.n_order_N = (u32)n_order_N;
	c->esz.singleton_bytes = sbytes; c->esz.singleton_tail = stail; c->esz.input_N_bytes = nbytesN;
	c->esz.aligned_singletons = n_s - U_s; c->esz.aligned_N = c->n_N - U_N;
	if (c->shard_world > 1) { // count what THIS GPU's contigs took
		c->esz.aligned_singletons = M_single;
		c->esz.aligned_N = M - M_single;
	}
	c->s2_keep.push_back(posb); c->s2_keep.push_back(noise); c->s2_keep.push_back(noisepos);
	void *tmp[] = { pflags, pcur, tile_idx, ns, ex, nat_idx, cs, cid, cstart, inc, G, scan_tmp, d_tot32, d_tot64, cons2, best, prio_u, iprio, af, exa, rid_u, irid,
	                f_src, f_kind, f_col, nm1, noff, revc, isN, exN, ordv, uf, exU, ulist, d_tail };
	for (void *q : tmp) c->release(q);
	c->encoded = true;
	return 0;
}

// pack_order.cpp:20-77 on the order stream of the last encode: header {int numbits, u32 numreads}, packed blocks, tail
int s2_pack_order(harcgpu_ctx *c, void *h_packed, u32 *h_tail, u64 *packed_bytes, u32 *ntail)
{
	const u32 n = c->esz.n_order;
	const int numbits = n ? 32 - __builtin_clz(n) : 0; // (int)(log2(n) + 1)
	const u64 nblk = n / 32, nwords = nblk * (u64)numbits;
	if (packed_bytes) *packed_bytes = 8 + 4 * nwords;
	if (ntail) *ntail = n % 32;
	if (h_packed) {
		int hdr[2] = { numbits, (int)n };
		memcpy(h_packed, hdr, 8);
		if (nwords) {
			u32 *d = nullptr;
			if (c->alloc(&d, nwords)) return -1;
			pack_order_kernel<<<KL + cdiv(nwords, 256), 256, 0, c->st>>>(c->o_order, nwords, numbits, d);
			CK(cudaGetLastError());
			CK(cudaMemcpyAsync((char *)h_packed + 8, d, 4 * nwords, cudaMemcpyDeviceToHost, c->st));
			CK(cudaStreamSynchronize(c->st));
			c->release(d);
		}
	}
	if (h_tail && n % 32) {
		CK(cudaMemcpyAsync(h_tail, c->o_order + nblk * 32, 4 * (size_t)(n % 32), cudaMemcpyDeviceToHost, c->st));
		CK(cudaStreamSynchronize(c->st));
	}
	return 0;
}
