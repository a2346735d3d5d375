// Stage I of HARC -- hash-based read reordering (reference: src/reorder.cpp) -- as sm_100a kernels.
//
//   K1 pack2_kernel      ASCII lines -> 2 bits/base                     reorder.cpp:203-209, 240-263
//   K2 keys_kernel       (read & mask) >> 2*dict_start                  reorder.cpp:284-302
//   K3 radix sort + heads/scan/scatter + table insert                   reorder.cpp:305-390 (sort, dedup, MPHF, CSR fill)
//   K4 walk_kernel       warp-per-chain greedy walk                     reorder.cpp:434-703, 863-915
//   K5 finalize (chunk sort, partition) + unpack_kernel                 reorder.cpp:722-830
//
// Data layout in HBM: reads[n][NW] u64 (NW = ceil(2L/64)); per dictionary the canonical CSR (keys, start, ids) and a
// 16-byte-slot open-addressing table at load factor <= 0.5; a claim bitmap (1 bit/read); a chunked record log.
#include "ctx.h"
#include <cub/device/device_radix_sort.cuh>

// ------------------------------------------------------------------------------------------------ K1 pack
namespace {
constexpr int PACK_RPB = 64; // reads per block: 64*(L+1) bytes is a multiple of 16 for every L

__device__ __forceinline__ u32 code2_of(unsigned char ch)
{
	// (ch>>1)&3: A->0 C->1 T->2 G->3; reference code (reorder.cpp:188-195): A=0 G=1 C=2 T=3
	return (0x78u >> (2 * ((ch >> 1) & 3))) & 3u;
}
// One block stages 64 lines in shared memory with 16-byte loads, then one thread per output word packs it.
// WITH_N (stage II pool, encoder.cpp:731-745): 'N' is stored as code 0 and flagged in a second word array at the
// low bit of the base's pair; the reference's 3-bit code is then 2*code2 + nflag.
template <bool WITH_N>
__global__ void __launch_bounds__(256) pack_kernel(const char *__restrict__ ascii, u32 n, int L, int NWo, u64 *__restrict__ out,
                                                   u64 *__restrict__ outN)
{
	extern __shared__ uint4 stage4[];
	char *stage = reinterpret_cast<char *>(stage4);
	const size_t line = (size_t)L + 1;
	const size_t r0 = (size_t)blockIdx.x * PACK_RPB;
	const u32 nr = (u32)min((size_t)PACK_RPB, (size_t)n - r0);
	const size_t bytes = (size_t)nr * line;
	const char *src = ascii + r0 * line;
	const size_t nvec = bytes / 16;
	const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
	for (size_t v = threadIdx.x; v < nvec; v += blockDim.x) stage4[v] = __ldg(&src4[v]);
	for (size_t b = nvec * 16 + threadIdx.x; b < bytes; b += blockDim.x) stage[b] = src[b];
	__syncthreads();
	for (u32 w = threadIdx.x; w < nr * (u32)NWo; w += blockDim.x) {
		u32 r = w / NWo, k = w % NWo;
		const unsigned char *s = reinterpret_cast<const unsigned char *>(stage + (size_t)r * line);
		u64 v = 0, vn = 0;
		int b0 = k * 32;
#pragma unroll 8
		for (int c = 0; c < 32; c++)
			if (b0 + c < L) {
				unsigned char ch = s[b0 + c];
				if (WITH_N && ch == 'N') vn |= 1ull << (2 * c);
				else v |= (u64)code2_of(ch) << (2 * c);
			}
		out[(r0 + r) * NWo + k] = v;
		if (WITH_N) outN[(r0 + r) * NWo + k] = vn;
	}
}

// ------------------------------------------------------------------------------------------------ K5 unpack
// out line i = read[order[i]] (reverse-complemented if rev[i]=='r') as ASCII + '\n'   (reorder.cpp:740-760, 832-861)
__global__ void __launch_bounds__(256) unpack_kernel(const u64 *__restrict__ reads, const u32 *__restrict__ order,
                                                     const u8 *__restrict__ rev, u32 cnt, int L, int NW, char *__restrict__ out)
{
	const size_t line = (size_t)L + 1;
	const size_t r0 = (size_t)blockIdx.x * PACK_RPB;
	const u32 nr = (u32)min((size_t)PACK_RPB, (size_t)cnt - r0);
	const size_t bytes = (size_t)nr * line;
	for (size_t b = threadIdx.x; b < bytes; b += blockDim.x) {
		u32 r = (u32)(b / line);
		int c = (int)(b % line);
		char ch = '\n';
		if (c < L) {
			u32 id = order[r0 + r];
			bool rc = rev && rev[r0 + r] == 'r';
			int src = rc ? L - 1 - c : c;
			u32 v = (u32)(__ldg(&reads[(size_t)id * NW + (src >> 5)]) >> (2 * (src & 31))) & 3u;
			if (rc) v ^= 3u;
			ch = "AGCT"[v];
		}
		out[r0 * line + b] = ch;
	}
}

// ------------------------------------------------------------------------------------------------ K2/K3 dictionary
__global__ void __launch_bounds__(256) keys_kernel(const u64 *__restrict__ reads, u32 n, int words, int bitpos, int nbits,
                                                   u64 *__restrict__ keys, u32 *__restrict__ ids)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 *r = reads + (size_t)i * words;
	int q = bitpos >> 6, sh = bitpos & 63;
	u64 v = __ldg(&r[q]) >> sh;
	if (sh && q + 1 < words) v |= __ldg(&r[q + 1]) << (64 - sh);
	if (nbits < 64) v &= (1ull << nbits) - 1;
	keys[i] = v;
	ids[i] = i;
}

// stage II keys (encoder.cpp:893-911): the reference's 3-bit code of base b is 2*code2 + nflag
__global__ void __launch_bounds__(256) keys3_kernel(const u64 *__restrict__ r2, const u64 *__restrict__ rN, u32 n, int words, int ds,
                                                    int de, u64 *__restrict__ keys, u32 *__restrict__ ids)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	u64 key = 0;
	for (int b = ds; b <= de; b++) {
		u64 c2 = (__ldg(&r2[(size_t)i * words + (b >> 5)]) >> (2 * (b & 31))) & 3ull;
		u64 nf = (__ldg(&rN[(size_t)i * words + (b >> 5)]) >> (2 * (b & 31))) & 1ull;
		key |= ((c2 << 1) | nf) << (3 * (b - ds));
	}
	keys[i] = key;
	ids[i] = i;
}

__global__ void __launch_bounds__(256) heads_kernel(const u64 *__restrict__ ks, u32 n, u32 *__restrict__ head)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	head[i] = (i == 0 || ks[i] != ks[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) bins_kernel(const u64 *__restrict__ ks, const u32 *__restrict__ head,
                                                   const u32 *__restrict__ binidx, u32 n, u32 numkeys,
                                                   u64 *__restrict__ keys, u32 *__restrict__ start)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i == 0) start[numkeys] = n;
	if (i >= n || !head[i]) return;
	u32 b = binidx[i];
	keys[b] = ks[i];
	start[b] = i;
}

// One thread per bin: linear probing, CAS on the (start,size) half of the slot; the table is read-only afterwards.
__global__ void __launch_bounds__(256) insert_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ start, u32 numkeys,
                                                     ulonglong2 *slots, u32 mask)
{
	u32 b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= numkeys) return;
	u64 key = keys[b];
	u32 s = start[b], sz = start[b + 1] - s;
	u64 val = (u64)s | ((u64)sz << 32);
	u32 h = (u32)mix64(key) & mask;
	while (true) {
		u64 old = atomicCAS(&slots[h].y, 0ull, val);
		if (old == 0ull) { slots[h].x = key; return; }
		h = (h + 1) & mask;
	}
}
} // namespace

int s1_pack_reads(harcgpu_ctx *c, const void *d_ascii, u32 n)
{
	if (n == 0) return 0;
	size_t smem = (size_t)PACK_RPB * (c->L + 1);
	smem = (smem + 15) / 16 * 16;
	pack_kernel<false><<<KL + cdiv(n, PACK_RPB), 256, smem, c->st>>>((const char *)d_ascii, n, c->L, c->NW, c->reads, nullptr);
	CK(cudaGetLastError());
	return 0;
}
// d_ascii must be 16-byte aligned.  out2/outN: [n][NW]
int s1_packN(harcgpu_ctx *c, const void *d_ascii, u32 n, u64 *out2, u64 *outN)
{
	if (n == 0) return 0;
	size_t smem = (size_t)PACK_RPB * (c->L + 1);
	smem = (smem + 15) / 16 * 16;
	if (outN) pack_kernel<true><<<KL + cdiv(n, PACK_RPB), 256, smem, c->st>>>((const char *)d_ascii, n, c->L, c->NW, out2, outN);
	else pack_kernel<false><<<KL + cdiv(n, PACK_RPB), 256, smem, c->st>>>((const char *)d_ascii, n, c->L, c->NW, out2, nullptr);
	CK(cudaGetLastError());
	return 0;
}
int s1_unpack_reads(harcgpu_ctx *c, const u64 *reads, const u32 *order, const u8 *rev, u32 cnt, char *d_out)
{
	if (cnt == 0) return 0;
	unpack_kernel<<<KL + cdiv(cnt, PACK_RPB), 256, 0, c->st>>>(reads, order, rev, cnt, c->L, c->NW, d_out);
	CK(cudaGetLastError());
	return 0;
}

void free_dict(harcgpu_ctx *c, DictDev &d)
{
	c->release(d.keys); c->release(d.start); c->release(d.ids); c->release(d.slots);
	d = DictDev();
}

// bits == 2: stage I key = bits [2*ds, 2*(de+1)) of the packed read; bits == 3: stage II key from (reads, readsN)
int build_dict(harcgpu_ctx *c, DictDev &d, const u64 *reads, const u64 *readsN, u32 n, int words, int ds, int de, int bits)
{
	cudaStream_t st = c->st;
	free_dict(c, d);
	const int bitpos = bits * ds, nbits = bits * (de - ds + 1);
	d.bitpos = bitpos; d.nbits = nbits;
	u64 *k_in = nullptr, *k_out = nullptr, *scan_tmp = nullptr;
	u32 *id_in = nullptr, *head = nullptr, *binidx = nullptr, *d_total = nullptr;
	void *cub_tmp = nullptr;
	if (c->alloc(&d.ids, n)) return -1;
	if (n == 0) {
		d.numkeys = 0;
		if (c->alloc(&d.keys, 1) || c->alloc(&d.start, 1)) return -1;
		CK(cudaMemsetAsync(d.start, 0, 4, st));
		d.slot_mask = 15;
		if (c->alloc(&d.slots, 16)) return -1;
		CK(cudaMemsetAsync(d.slots, 0, 16 * sizeof(ulonglong2), st));
		return 0;
	}
	if (c->alloc(&k_in, n) || c->alloc(&k_out, n) || c->alloc(&id_in, n) || c->alloc(&head, n) || c->alloc(&binidx, n) ||
	    c->alloc(&scan_tmp, scan_tmp_elems(n)) || c->alloc(&d_total, 1))
		return -1;
	if (bits == 2) keys_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(reads, n, words, bitpos, nbits, k_in, id_in);
	else keys3_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(reads, readsN, n, words, ds, de, k_in, id_in);
	CK(cudaGetLastError());
	size_t tb = 0;
	CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in, k_out, id_in, d.ids, (int64_t)n, 0, nbits, st));
	if (c->alloc((char **)&cub_tmp, tb)) return -1;
	// LSD radix sort is stable and ids start ascending, so ids stay ascending inside a bin (reorder.cpp:371-384)
	CK(cub::DeviceRadixSort::SortPairs(cub_tmp, tb, k_in, k_out, id_in, d.ids, (int64_t)n, 0, nbits, st));
	heads_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(k_out, n, head);
	CK(cudaGetLastError());
	if (exclusive_scan_u32(head, binidx, n, scan_tmp, d_total, st)) return -1;
	u32 nk = 0;
	CK(cudaMemcpyAsync(&nk, d_total, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	d.numkeys = nk;
	if (c->alloc(&d.keys, nk) || c->alloc(&d.start, (size_t)nk + 1)) return -1;
	bins_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(k_out, head, binidx, n, nk, d.keys, d.start);
	CK(cudaGetLastError());
	u64 cap = 16;
	while (cap < 2ull * nk) cap <<= 1;
	d.slot_mask = (u32)(cap - 1);
	if (c->alloc(&d.slots, cap)) return -1;
	CK(cudaMemsetAsync(d.slots, 0, cap * sizeof(ulonglong2), st));
	insert_kernel<<<KL + cdiv(nk, 256), 256, 0, st>>>(d.keys, d.start, nk, d.slots, d.slot_mask);
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(st));
	c->release(k_in); c->release(k_out); c->release(id_in); c->release(head); c->release(binidx);
	c->release(scan_tmp); c->release(d_total); c->release(cub_tmp);
	return 0;
}

// ------------------------------------------------------------------------------------------------ K4 walk
namespace {
constexpr int WALK_WARPS = 4;  // warps (= walkers) per block
constexpr int CHUNK = 32;      // records per log chunk

// record: rid | pos<<32 | rev<<40 | matched<<41 | singleton<<42
__device__ __forceinline__ u64 mkrec(u32 rid, u32 pos, u32 rev, u32 matched, u32 single)
{
	return (u64)rid | ((u64)pos << 32) | ((u64)rev << 40) | ((u64)matched << 41) | ((u64)single << 42);
}

struct WalkArgs {
	const u64 *reads;
	u32 n;
	int L, maxmatch, thresh, maxsearch, numdict;
	DictView d[2];
	int kbits[2];
	u32 *claim;
	u32 *stripe_done;
	u32 walkers;
	// record log
	u64 *recs;
	u64 *chunk_key;
	u32 *chunk_fill;
	u32 *chunk_ctr;
	u32 max_chunks;
	u64 *counters;
};

template <int NW>
struct alignas(16) WalkSmem {
	u64 ref[NW];   // consensus of the current window, 2 bits/base (reorder.cpp:466)
	u64 rref[NW];  // its reverse complement
	u64 cur[NW + (NW & 1)]; // the read just appended (padded so the vote counts that follow stay 16-byte aligned)
};

template <int NW>
__device__ __forceinline__ void load_read(const u64 *__restrict__ reads, u32 rid, u64 (&rw)[NW])
{
	const u64 *r = reads + (size_t)rid * NW;
	if (NW % 2 == 0) {
		const ulonglong2 *r2 = reinterpret_cast<const ulonglong2 *>(r);
#pragma unroll
		for (int k = 0; k < NW / 2; k++) { ulonglong2 v = __ldg(&r2[k]); rw[2 * k] = v.x; rw[2 * k + 1] = v.y; }
	} else {
#pragma unroll
		for (int k = 0; k < NW; k++) rw[k] = __ldg(&r[k]);
	}
}

__device__ __forceinline__ u64 revpairs64_w(u64 x) // reverse the order of the 32 base pairs of a word
{
	u64 y = __brevll(x);
	return ((y & 0x5555555555555555ull) << 1) | ((y >> 1) & 0x5555555555555555ull);
}

// updaterefcount (reorder.cpp:863-915).  Vote counts live in a circular buffer (origin `head`) of uint4 {A,C,G,T} per
// position instead of being shifted; lane handles bases lane, lane+32, ...; the consensus word t is assembled with
// two warp OR-reductions (lanes 0-15 -> low half, 16-31 -> high half); the reverse complement is derived from the
// finished words by lanes 0..NW-1 (pair reversal + shift + complement).
template <int NW>
__device__ __forceinline__ void update_ref(WalkSmem<NW> &s, uint4 *cnt, int L, int lane, bool reset, bool rev, int shift, int &head)
{
	if (reset) head = 0;
	else { head += shift; if (head >= L) head -= L; }
	const int sh = 2 * (lane & 15);
#pragma unroll
	for (int t = 0; t < NW; t++) {
		const int i = lane + 32 * t;
		u32 out = 0;
		if (i < L) {
			const int src = rev ? L - 1 - i : i;
			u32 cc = (u32)(s.cur[src >> 5] >> (2 * (src & 31))) & 3u;
			if (rev) cc ^= 3u;
			int slot = head + i;
			if (slot >= L) slot -= L;
			// counts are kept in chartoint order A,C,G,T (reorder.cpp:139-142); the bit code is A0 G1 C2 T3
			uint4 v = (reset || i >= L - shift) ? make_uint4(0, 0, 0, 0) : cnt[slot];
			v.x += cc == 0; v.y += cc == 2; v.z += cc == 1; v.w += cc == 3;
			cnt[slot] = v;
			// argmax, ties -> A < C < G < T with strict '>' from max = 0 (reorder.cpp:893-899)
			u32 mx = v.x;
			if (v.y > mx) { mx = v.y; out = 2; }
			if (v.z > mx) { mx = v.z; out = 1; }
			if (v.w > mx) { mx = v.w; out = 3; }
		}
		const u32 piece = out << sh;
		const u32 lo = __reduce_or_sync(0xffffffffu, lane < 16 ? piece : 0u);
		const u32 hi = __reduce_or_sync(0xffffffffu, lane >= 16 ? piece : 0u);
		if (lane == 0) s.ref[t] = (u64)lo | ((u64)hi << 32);
	}
	__syncwarp();
	if (lane < NW) {
		const int sft = 2 * (32 * NW - L);
		const u64 t0 = revpairs64_w(s.ref[NW - 1 - lane]);
		const u64 t1 = lane + 1 < NW ? revpairs64_w(s.ref[NW - 2 - lane]) : 0ull;
		const u64 r = sft ? (t0 >> sft) | (t1 << (64 - sft)) : t0;
		s.rref[lane] = r ^ lowmask(2 * L - 64 * lane);
	}
	__syncwarp();
}

// popcount(ref ^ (read & mask[j])) with ref >>= 2j (forward, reorder.cpp:543) or
// popcount(revref ^ (read & revmask[j])) with revref <<= 2j (reverse, reorder.cpp:608)
template <int NW>
__device__ __forceinline__ int hamming(const WalkSmem<NW> &s, const u64 (&rw)[NW], int L, int j, bool rev)
{
	const int sh = 2 * j, q = sh >> 6, r = sh & 63;
	int d = 0;
	if (!rev) {
		const int nb = 2 * (L - j);
#pragma unroll
		for (int k = 0; k < NW; k++) {
			u64 lo = k + q < NW ? s.ref[k + q] : 0ull, hi = k + q + 1 < NW ? s.ref[k + q + 1] : 0ull;
			u64 x = r ? (lo >> r) | (hi << (64 - r)) : lo;
			d += __popcll(x ^ (rw[k] & lowmask(nb - 64 * k)));
		}
	} else {
#pragma unroll
		for (int k = 0; k < NW; k++) {
			u64 lo = k - q >= 0 ? s.rref[k - q] : 0ull, hi = k - q - 1 >= 0 ? s.rref[k - q - 1] : 0ull;
			u64 x = r ? (lo << r) | (hi >> (64 - r)) : lo;
			x &= lowmask(2 * L - 64 * k);
			d += __popcll(x ^ (rw[k] & ~lowmask(sh - 64 * k)));
		}
	}
	return d;
}

__device__ __forceinline__ bool is_unclaimed(const u32 *claim, u32 rid)
{
	u32 w = *((const volatile u32 *)&claim[rid >> 5]);
	return (w >> (rid & 31)) & 1u;
}
__device__ __forceinline__ bool try_claim(u32 *claim, u32 rid)
{
	u32 bit = 1u << (rid & 31);
	u32 old = atomicAnd(&claim[rid >> 5], ~bit);
	return (old & bit) != 0;
}

template <int NW>
__global__ void __launch_bounds__(WALK_WARPS * 32) walk_kernel(WalkArgs a)
{
	extern __shared__ uint4 smem_raw[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u32 wid = blockIdx.x * WALK_WARPS + warp;
	if (wid >= a.walkers) return;
	const int L = a.L, LP = (L + 31) & ~31;
	// per-warp shared state
	const size_t per_warp = sizeof(WalkSmem<NW>) + (size_t)16 * LP;
	char *base = reinterpret_cast<char *>(smem_raw) + per_warp * warp;
	WalkSmem<NW> &s = *reinterpret_cast<WalkSmem<NW> *>(base);
	uint4 *cnt = reinterpret_cast<uint4 *>(base + sizeof(WalkSmem<NW>));

	u64 c_steps = 0, c_probes = 0, c_hits = 0, c_cmp = 0, c_fail = 0, c_restart = 0;

	// record log state (lane 0 only)
	u32 chunk = 0xffffffffu, fill = CHUNK, seq = 0;
	auto emit = [&](u64 rec) {
		if (fill == CHUNK) {
			if (chunk != 0xffffffffu) a.chunk_fill[chunk] = CHUNK;
			chunk = atomicAdd(a.chunk_ctr, 1u);
			if (chunk >= a.max_chunks) { chunk = a.max_chunks - 1; } // cannot happen: max_chunks = n/CHUNK + walkers + 1
			a.chunk_key[chunk] = ((u64)wid << 32) | seq++;
			fill = 0;
		}
		a.recs[(size_t)chunk * CHUNK + fill++] = rec;
	};

	// reorder.cpp:476-497: walker t starts at read t*(n/T); gives up at once if that read is taken
	u32 current = (u32)((u64)wid * (a.n / a.walkers));
	int ok = 0;
	if (lane == 0) ok = try_claim(a.claim, current);
	ok = __shfl_sync(0xffffffffu, ok, 0);
	if (!ok) return;
	c_restart++;
	int head = 0;
	if (lane < NW) s.cur[lane] = __ldg(&a.reads[(size_t)current * NW + lane]);
	__syncwarp();
	update_ref<NW>(s, cnt, L, lane, true, false, 0, head);
	bool prev_unmatched = true;
	u32 prev = current;
	// restart state: current stripe, downward cursor inside it, stripes visited
	u32 stripe = wid, stripes_tried = 0;
	long long cursor = (long long)((((u64)wid + 1) * a.n) / a.walkers) - 1;

	const int kind = lane & 3;          // 0: fwd dict0, 1: fwd dict1, 2: rev dict0, 3: rev dict1 (reorder.cpp:517-643 order)
	const bool rev = kind >= 2;
	const int l = kind & 1;
	const bool dict_on = l < a.numdict;
	const DictView dv = a.d[l];
	const int kb = a.kbits[l];

	while (true) {
		c_steps++;
		// ---- search: all (shift, direction, dictionary) probes of 8 consecutive shifts in flight at once; the lowest
		// lane with a claimable candidate wins, which is exactly the sequential order of reorder.cpp:517-649.
		bool found = false;
		u32 k_rid = 0;
		int k_j = 0, k_rev = 0;
		for (int jb = 0; jb < a.maxmatch && !found; jb += 8) {
			const int j = jb + (lane >> 2);
			bool valid = dict_on && j < a.maxmatch && (rev ? dv.dstart > j : dv.dend + j < L);
			u32 bstart = 0, bsize = 0;
			bool hit = false;
			if (valid) {
				u64 key = rev ? getbits(s.rref, NW, 2 * (dv.dstart - j), kb) : getbits(s.ref, NW, 2 * (dv.dstart + j), kb);
				hit = dict_lookup(dv, key, bstart, bsize);
			}
			c_probes += __popc(__ballot_sync(0xffffffffu, valid));
			c_hits += __popc(__ballot_sync(0xffffffffu, hit));
			// candidate scan state: next index to look at (descending), live entries seen so far
			u32 left = hit ? bsize : 0u; // entries of the bin not looked at yet (scanned from the tail, reorder.cpp:540)
			int seen = 0;
			u32 cand = 0xffffffffu;
			auto advance = [&]() {
				cand = 0xffffffffu;
				while (left > 0 && seen < a.maxsearch) {
					left--;
					u32 rid = __ldg(&dv.ids[bstart + left]);
					if (!is_unclaimed(a.claim, rid)) continue; // removed from the bin in the reference (505-514)
					seen++;
					u64 rw[NW];
					load_read<NW>(a.reads, rid, rw);
					c_cmp++;
					if (hamming<NW>(s, rw, L, j, rev) <= a.thresh) { cand = rid; break; }
				}
			};
			if (hit) advance();
			while (true) {
				u32 bal = __ballot_sync(0xffffffffu, cand != 0xffffffffu);
				if (!bal) break;
				int win = __ffs(bal) - 1;
				int got = 0;
				if (lane == win) {
					got = try_claim(a.claim, cand);
					if (!got) c_fail++;
				}
				got = __shfl_sync(0xffffffffu, got, win);
				if (got) {
					found = true;
					k_rid = __shfl_sync(0xffffffffu, cand, win);
					k_j = jb + (win >> 2);
					k_rev = (win & 3) >= 2;
					break;
				}
				if (lane == win) advance();
			}
		}
		if (found) {
			// reorder.cpp:560-578 / 624-641
			current = k_rid;
			if (lane < NW) s.cur[lane] = __ldg(&a.reads[(size_t)current * NW + lane]);
			__syncwarp();
			update_ref<NW>(s, cnt, L, lane, false, k_rev, k_j, head);
			if (lane == 0) {
				if (prev_unmatched) emit(mkrec(prev, (u32)L, 0, 0, 0));
				emit(mkrec(current, (u32)k_j, (u32)k_rev, 1, 0));
			}
			prev_unmatched = false;
			continue;
		}
		// ---- no match: new chain head (reorder.cpp:650-688).  The reference takes the highest unclaimed index through a
		// private downward cursor per thread.  Here the reads are cut into one stripe per walker; a walker scans its own
		// stripe downward first and then the following stripes, so concurrent restarts do not fight over one bit.  With
		// one walker the stripe is the whole array and the choice is exactly the reference's.
		bool got_head = false;
		while (stripes_tried < a.walkers) {
			const long long slo = (long long)(((u64)stripe * a.n) / a.walkers);
			if (cursor < slo) {
				// stripe exhausted: everything in it is claimed for good
				if (lane == 0) a.stripe_done[stripe] = 1u;
				// move to the next stripe (cyclically) that is not known to be finished, 32 flags at a time
				bool found_stripe = false;
				while (stripes_tried + 1 < a.walkers) {
					const u32 span = min(32u, a.walkers - 1 - stripes_tried);
					u32 cand = stripe + 1 + lane;
					if (cand >= a.walkers) cand -= a.walkers;
					const bool open_ = (u32)lane < span && *((volatile u32 *)&a.stripe_done[cand]) == 0u;
					const u32 bal = __ballot_sync(0xffffffffu, open_);
					if (bal) {
						const u32 f = __ffs(bal) - 1;
						stripe = stripe + 1 + f;
						if (stripe >= a.walkers) stripe -= a.walkers;
						stripes_tried += f + 1;
						found_stripe = true;
						break;
					}
					stripe += span;
					if (stripe >= a.walkers) stripe -= a.walkers;
					stripes_tried += span;
				}
				if (!found_stripe) { stripes_tried = a.walkers; break; }
				cursor = (long long)((((u64)stripe + 1) * a.n) / a.walkers) - 1;
				continue;
			}
			const long long topw = cursor >> 5;
			const long long wi = topw - lane;
			u32 word = (wi >= 0 && wi >= (slo >> 5)) ? *((volatile u32 *)&a.claim[wi]) : 0u;
			if (lane == 0) { int bt = (int)(cursor & 31); if (bt != 31) word &= (2u << bt) - 1u; }
			if (wi == (slo >> 5)) word &= ~((1u << (slo & 31)) - 1u);
			u32 bal = __ballot_sync(0xffffffffu, word != 0u);
			if (!bal) { cursor = (topw - 31) * 32 - 1; continue; }
			int src = __ffs(bal) - 1;
			u32 wv = __shfl_sync(0xffffffffu, word, src);
			int bit = 31 - __clz(wv);
			u32 j = (u32)((topw - src) * 32 + bit);
			int got = 0;
			if (lane == 0) got = try_claim(a.claim, j);
			got = __shfl_sync(0xffffffffu, got, 0);
			cursor = (long long)j - 1; // j is claimed now, by this walker or by another one
			if (got) { current = j; got_head = true; break; }
		}
		if (lane == 0 && prev_unmatched) emit(mkrec(prev, 0, 0, 0, 1)); // previous head was a singleton (672-684)
		if (!got_head) break;
		c_restart++;
		if (lane < NW) s.cur[lane] = __ldg(&a.reads[(size_t)current * NW + lane]);
		__syncwarp();
		update_ref<NW>(s, cnt, L, lane, true, false, 0, head);
		prev_unmatched = true;
		prev = current;
	}
	if (lane == 0 && chunk != 0xffffffffu) a.chunk_fill[chunk] = fill;
	// counters: steps/probes/hits/restarts are warp-uniform, compares/fails per lane
	for (int o = 16; o > 0; o >>= 1) {
		c_cmp += __shfl_xor_sync(0xffffffffu, c_cmp, o);
		c_fail += __shfl_xor_sync(0xffffffffu, c_fail, o);
	}
	if (lane == 0) {
		atomicAdd(&a.counters[0], c_steps); atomicAdd(&a.counters[1], c_probes); atomicAdd(&a.counters[2], c_hits);
		atomicAdd(&a.counters[3], c_cmp); atomicAdd(&a.counters[4], c_fail); atomicAdd(&a.counters[5], c_restart);
	}
}

__global__ void __launch_bounds__(256) init_claim_kernel(u32 *claim, u32 n)
{
	u32 w = blockIdx.x * blockDim.x + threadIdx.x;
	u32 nw = (n + 31) / 32;
	if (w >= nw) return;
	u32 v = 0xffffffffu;
	if (w == nw - 1 && (n & 31)) v = (1u << (n & 31)) - 1u;
	claim[w] = v;
}

// ---- finalize: order the chunks by (walker, sequence) and split matched / singleton records ----------------
__global__ void __launch_bounds__(256) iota_kernel(u32 *v, u32 n)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) v[i] = i;
}
// one warp per chunk (in sorted order): count matched and singleton records
__global__ void __launch_bounds__(256) chunk_count_kernel(const u64 *__restrict__ recs, const u32 *__restrict__ sorted_chunk,
                                                          const u32 *__restrict__ chunk_fill, u32 nchunks,
                                                          u32 *__restrict__ cm, u32 *__restrict__ cs)
{
	u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (w >= nchunks) return;
	u32 ch = sorted_chunk[w], f = chunk_fill[ch];
	bool single = false, valid = lane < f;
	if (valid) single = (recs[(size_t)ch * CHUNK + lane] >> 42) & 1ull;
	u32 bs = __ballot_sync(0xffffffffu, valid && single);
	if (lane == 0) { cs[w] = __popc(bs); cm[w] = f - __popc(bs); }
}
__global__ void __launch_bounds__(256) chunk_gather_kernel(const u64 *__restrict__ recs, const u32 *__restrict__ sorted_chunk,
                                                           const u32 *__restrict__ chunk_fill, u32 nchunks,
                                                           const u32 *__restrict__ om, const u32 *__restrict__ os,
                                                           u32 *__restrict__ order, u8 *__restrict__ rev, u8 *__restrict__ flag,
                                                           u8 *__restrict__ pos, u32 *__restrict__ order_s)
{
	u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (w >= nchunks) return;
	u32 ch = sorted_chunk[w], f = chunk_fill[ch];
	bool valid = lane < f;
	u64 r = valid ? recs[(size_t)ch * CHUNK + lane] : 0ull;
	bool single = valid && ((r >> 42) & 1ull);
	u32 bs = __ballot_sync(0xffffffffu, single), bm = __ballot_sync(0xffffffffu, valid && !single);
	u32 below = (1u << lane) - 1u;
	if (single) order_s[os[w] + __popc(bs & below)] = (u32)r;
	else if (valid) {
		u32 dst = om[w] + __popc(bm & below);
		order[dst] = (u32)r;
		pos[dst] = (u8)(r >> 32);
		rev[dst] = ((r >> 40) & 1ull) ? 'r' : 'd';
		flag[dst] = ((r >> 41) & 1ull) ? '1' : '0';
	}
}
} // namespace

template <int NW>
static int launch_walk(harcgpu_ctx *c, const WalkArgs &a)
{
	int LP = (c->L + 31) & ~31;
	size_t per_warp = sizeof(WalkSmem<NW>) + (size_t)16 * LP;
	size_t smem = per_warp * WALK_WARPS;
	CK(cudaFuncSetAttribute(walk_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	walk_kernel<NW><<<KL + cdiv(a.walkers, WALK_WARPS), WALK_WARPS * 32, smem, c->st>>>(a);
	CK(cudaGetLastError());
	return 0;
}

// walkers that can be resident at once: a walker that is not resident only starts after the others have finished
template <int NW>
static int resident_warps(harcgpu_ctx *c, u32 *out)
{
	int LP = (c->L + 31) & ~31;
	size_t smem = (sizeof(WalkSmem<NW>) + (size_t)16 * LP) * WALK_WARPS;
	int nb = 0, sms = 0;
	CK(cudaFuncSetAttribute(walk_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, walk_kernel<NW>, WALK_WARPS * 32, smem));
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
	*out = (u32)nb * (u32)sms * WALK_WARPS;
	return 0;
}
static int walk_resident_warps(harcgpu_ctx *c, u32 *out)
{
	switch (c->NW) {
	case 1: return resident_warps<1>(c, out);
	case 2: return resident_warps<2>(c, out);
	case 3: return resident_warps<3>(c, out);
	case 4: return resident_warps<4>(c, out);
	case 5: return resident_warps<5>(c, out);
	case 6: return resident_warps<6>(c, out);
	case 7: return resident_warps<7>(c, out);
	case 8: return resident_warps<8>(c, out);
	}
	harcgpu_set_error("unsupported read length %d", c->L);
	return -1;
}

int s1_reorder(harcgpu_ctx *c)
{
	cudaStream_t st = c->st;
	const u32 n = c->n;
	c->reordered = false;
	c->release(c->order); c->release(c->order_s); c->release(c->rev); c->release(c->flag); c->release(c->pos);
	c->order = c->order_s = nullptr; c->rev = c->flag = c->pos = nullptr;
	c->n_matched = c->n_single = c->n_unmatched = 0;
	if (c->alloc(&c->order, n) || c->alloc(&c->order_s, n) || c->alloc(&c->rev, n) || c->alloc(&c->flag, n) || c->alloc(&c->pos, n))
		return -1;
	CK(cudaMemsetAsync(c->counters, 0, 8 * sizeof(u64), st));
	if (n == 0) { c->reordered = true; c->ms["walk"] = 0; c->ms["finalize"] = 0; return 0; }

	// walkers: the reference's num_thr.  Auto: one walker per 2048 reads (each extra walker costs ~4 chain
	// heads; >= 2000-4000 reads per walker keeps the size within budget), capped at 32 resident warps per SM.
	u32 resident = 0;
	if (walk_resident_warps(c, &resident)) return -1;
	u32 walkers = c->p.walkers > 0 ? (u32)c->p.walkers : (u32)std::min<u64>(resident, std::max<u64>(1, n / 2048));
	if (walkers > n) walkers = n;
	c->walkers_used = walkers;

	u32 max_chunks = n / CHUNK + walkers + 1;
	u64 *recs = nullptr, *chunk_key = nullptr, *key_sorted = nullptr, *scan_tmp = nullptr;
	u32 *chunk_fill = nullptr, *chunk_ctr = nullptr, *chunk_id = nullptr, *chunk_sorted = nullptr, *cm = nullptr, *cs = nullptr,
	    *om = nullptr, *os = nullptr, *totals = nullptr;
	if (c->alloc(&recs, (size_t)max_chunks * CHUNK) || c->alloc(&chunk_key, max_chunks) || c->alloc(&chunk_fill, max_chunks) ||
	    c->alloc(&chunk_ctr, 1))
		return -1;
	CK(cudaMemsetAsync(chunk_ctr, 0, 4, st));
	CK(cudaMemsetAsync(chunk_fill, 0, 4 * (size_t)max_chunks, st));
	init_claim_kernel<<<KL + cdiv((n + 31) / 32, 256), 256, 0, st>>>(c->claim, n);
	CK(cudaGetLastError());
	u32 *stripe_done = nullptr;
	if (c->alloc(&stripe_done, walkers)) return -1;
	CK(cudaMemsetAsync(stripe_done, 0, 4 * (size_t)walkers, st));

	WalkArgs a;
	a.reads = c->reads; a.n = n; a.L = c->L; a.maxmatch = c->p.maxmatch; a.thresh = c->p.thresh; a.maxsearch = c->p.maxsearch;
	a.numdict = c->p.numdict;
	for (int l = 0; l < 2; l++) {
		int ll = l < c->p.numdict ? l : 0;
		a.d[l].slots = c->d1[ll].slots; a.d[l].ids = c->d1[ll].ids; a.d[l].slot_mask = c->d1[ll].slot_mask;
		a.d[l].dstart = c->p.dict_start[ll]; a.d[l].dend = c->p.dict_end[ll];
		a.kbits[l] = c->d1[ll].nbits;
	}
	a.claim = c->claim; a.stripe_done = stripe_done; a.walkers = walkers;
	a.recs = recs; a.chunk_key = chunk_key; a.chunk_fill = chunk_fill; a.chunk_ctr = chunk_ctr; a.max_chunks = max_chunks;
	a.counters = c->counters;
	c->tic();
	int rc = -1;
	switch (c->NW) {
	case 1: rc = launch_walk<1>(c, a); break;
	case 2: rc = launch_walk<2>(c, a); break;
	case 3: rc = launch_walk<3>(c, a); break;
	case 4: rc = launch_walk<4>(c, a); break;
	case 5: rc = launch_walk<5>(c, a); break;
	case 6: rc = launch_walk<6>(c, a); break;
	case 7: rc = launch_walk<7>(c, a); break;
	case 8: rc = launch_walk<8>(c, a); break;
	default: harcgpu_set_error("unsupported read length %d", c->L); return -1;
	}
	if (rc) return rc;
	c->toc("walk");
	CK(cudaGetLastError());

	// ---- finalize
	c->tic();
	u32 nchunks = 0;
	CK(cudaMemcpyAsync(&nchunks, chunk_ctr, 4, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	if (nchunks > max_chunks) { harcgpu_set_error("record log overflow"); return -1; }
	if (c->alloc(&key_sorted, nchunks) || c->alloc(&chunk_id, nchunks) || c->alloc(&chunk_sorted, nchunks) || c->alloc(&cm, nchunks) ||
	    c->alloc(&cs, nchunks) || c->alloc(&om, nchunks) || c->alloc(&os, nchunks) || c->alloc(&totals, 2) ||
	    c->alloc(&scan_tmp, scan_tmp_elems(nchunks)))
		return -1;
	iota_kernel<<<KL + cdiv(nchunks, 256), 256, 0, st>>>(chunk_id, nchunks);
	CK(cudaGetLastError());
	size_t tb = 0;
	void *cub_tmp = nullptr;
	CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, chunk_key, key_sorted, chunk_id, chunk_sorted, (int64_t)nchunks, 0, 64, st));
	if (c->alloc((char **)&cub_tmp, tb)) return -1;
	CK(cub::DeviceRadixSort::SortPairs(cub_tmp, tb, chunk_key, key_sorted, chunk_id, chunk_sorted, (int64_t)nchunks, 0, 64, st));
	chunk_count_kernel<<<KL + cdiv((size_t)nchunks * 32, 256), 256, 0, st>>>(recs, chunk_sorted, chunk_fill, nchunks, cm, cs);
	CK(cudaGetLastError());
	if (exclusive_scan_u32(cm, om, nchunks, scan_tmp, totals, st)) return -1;
	if (exclusive_scan_u32(cs, os, nchunks, scan_tmp, totals + 1, st)) return -1;
	chunk_gather_kernel<<<KL + cdiv((size_t)nchunks * 32, 256), 256, 0, st>>>(recs, chunk_sorted, chunk_fill, nchunks, om, os, c->order,
	                                                                      c->rev, c->flag, c->pos, c->order_s);
	CK(cudaGetLastError());
	u32 tot[2];
	CK(cudaMemcpyAsync(tot, totals, 8, cudaMemcpyDeviceToHost, st));
	u64 cnt[8];
	CK(cudaMemcpyAsync(cnt, c->counters, 64, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	c->toc("finalize");
	c->n_matched = tot[0]; c->n_single = tot[1]; c->n_unmatched = (u32)cnt[5];
	if ((u64)tot[0] + tot[1] != n) { harcgpu_set_error("reorder lost reads: %u matched + %u singletons != %u", tot[0], tot[1], n); return -1; }
	c->release(recs); c->release(chunk_key); c->release(key_sorted); c->release(chunk_fill); c->release(chunk_ctr);
	c->release(chunk_id); c->release(chunk_sorted); c->release(cm); c->release(cs); c->release(om); c->release(os);
	c->release(totals); c->release(scan_tmp); c->release(cub_tmp); c->release(stripe_done);
	c->reordered = true;
	return 0;
}
