// Stage I chain walk of HARC (reference: reorder.cpp:434-703 reorder(), 863-915 updaterefcount) as one sm_100a kernel.
//
// A walker is the reference's OpenMP thread: it follows one chain of overlapping reads at a time.  Here a walker is a
// group of G = 8 lanes, four walkers per warp, thousands per GPU.  Per round the 8 lanes of a walker issue the
// reference's four probe kinds (forward dict 0, forward dict 1, reverse dict 0, reverse dict 1) for eight consecutive
// shifts (four independent probes per lane); the lowest lane with a claimable candidate wins, which is exactly the sequential order of
// reorder.cpp:517-649, so one walker reproduces the reference at num_thr=1 byte for byte.  Reads are claimed with one
// atomicAnd on a bitmap (instead of the reference's 2 x 2^24 striped locks and its in-place bin compaction); the
// consensus window of updaterefcount lives in shared memory as a circular array of vote counts.
//
// Not in the reference (switchable, params.extend): a new chain is first extended to the LEFT of its head by walking the
// reverse-complement strand, and that run is written in front of the head in reverse order.  The streams only encode
// order, orientation and shifts, so the reference's decoders read the result unchanged; it removes
// the orphaned left halves that every restart otherwise leaves behind, which is what lets thousands of concurrent
// walkers stay inside the 2 % size budget.
#include "ctx.h"
#include <algorithm>
#include <stdlib.h>
#include <cub/device/device_radix_sort.cuh>

namespace {
// A walker is a group of G lanes (G = 8, 16 or 32: template parameter), 32/G walkers per warp.  Per round a walker issues
// 32 probes: the four probe kinds of 8 consecutive shifts, i.e. U = 32/G independent probes per lane.
constexpr int WALK_WARPS = 4;         // warps per block
constexpr int CHUNK = 32;             // records per log chunk
constexpr u32 NONE = 0xffffffffu;
constexpr u32 FULL = 0xffffffffu;

enum { S_SEARCH = 0, S_CHAINEND, S_RESTART, S_NEWHEAD, S_DONE };

// record: rid | pos<<32 | rev<<40 | matched<<41 | singleton<<42
__device__ __forceinline__ u64 mkrec(u32 rid, u32 pos, u32 rev, u32 matched, u32 single)
{
	return (u64)rid | ((u64)pos << 32) | ((u64)rev << 40) | ((u64)matched << 41) | ((u64)single << 42);
}

struct WalkArgs {
	const u64 *reads;
	u32 n;
	int L, maxmatch, thresh, maxsearch, numdict, extend;
	DictView d[2];
	int kbits[2];
	// claim bitmap (1 = unclaimed).  One GPU: `claim` covers all n reads.  One job on several GPUs: the bitmap is cut into
	// `world` contiguous id ranges of `seg_per` reads, range r lives in the memory of GPU r (`seg[r]`, mapped through CUDA
	// IPC, read and claimed over NVLink); this GPU's walkers start and restart only inside its own range
	// [base, base + n_loc), whose words are `claim`.
	u32 *claim;
	u32 *hint; // sharded only: full-size local bitmap, bit cleared once this GPU knows the read is claimed (never authoritative)
	u32 *seg[8];
	u32 seg_per, base, n_loc;
	int world;
	u32 *stripe_done;
	u32 walkers;
	// forward record log: chunks of CHUNK records, ordered at finalize by (walker, sequence)
	u64 *recs;
	u64 *chunk_key;
	u32 *chunk_fill;
	u32 *chunk_ctr;
	u32 max_chunks;
	// scratch log of the left extension: chunks linked backwards
	u64 *lrecs;
	u32 *lprev;
	u32 *lchunk_ctr;
	u32 max_lchunks;
	u64 *counters;
};

template <int NW>
struct alignas(16) WalkerSmem {
	u32 key[4 * 32 * NW]; // vote keys (count << 2 | tie rank) per base code and window position, circular (see update_ref)
	u32 best[32 * NW];    // the largest of the four keys of a position
	u64 ref[NW];          // consensus of the current window, 2 bits/base (reorder.cpp:466)
	u64 rref[NW];         // its reverse complement
	u64 cur[NW];          // the read just appended
	long long cursor;     // restart: downward cursor inside the current stripe
	u32 stripe, stripes_tried;
	u32 chunk, fill, seq;                      // forward log
	u32 lchunk, lfill, lfirst, lcount, j1;     // left log
	u32 pend, pend_f;                          // left run: read found last, not yet written
	u32 pad_[3];
};

#ifdef WALK_PROF
// tuning aid (build with HARC_CUFLAGS=-DWALK_PROF): warp cycles per region of the walk, summed over all warps
__device__ unsigned long long g_walk_prof[16];
#define PROF_T0() long long pt_ = clock64()
#define PROF_ADD(i) do { long long t_ = clock64(); if (lane == 0) pacc[i] += (u64)(t_ - pt_); pt_ = t_; } while (0)
#define PROF_CNT(i, v) do { if (lane == 0) pacc[i] += (u64)(v); } while (0)
#else
#define PROF_T0()
#define PROF_ADD(i)
#define PROF_CNT(i, v)
#endif

__device__ __forceinline__ u32 ldvol(const u32 *p) { return *((const volatile u32 *)p); }

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

// bits [pos, pos+n) of a little-endian word array with zero fill outside [0, 64*words); pos may be negative, n <= 64
__device__ __forceinline__ u64 getbits_z(const u64 *w, int words, int pos, int n)
{
	const int q = pos >> 6, r = pos & 63; // arithmetic shift: floor division also for negative pos
	const u64 lo = (q >= 0 && q < words) ? w[q] : 0ull;
	const u64 hi = (q + 1 >= 0 && q + 1 < words) ? w[q + 1] : 0ull;
	u64 v = r ? (lo >> r) | (hi << (64 - r)) : lo;
	if (n < 64) v &= (1ull << n) - 1;
	return v;
}

// updaterefcount (reorder.cpp:863-915) for one walker, by its G lanes.
// Lane `sub` owns the B = 32*NW/G consecutive window positions [B*sub, B*sub+B).  Votes are kept as keys
// (count << 2 | tie rank), one array per base code, plus the largest key of every position: a read adds one vote per
// position, so only the voted key and the maximum change (one 4-byte read-modify-write each), and argmax with ties
// -> A < C < G < T (strict '>' from max = 0, reorder.cpp:893-899) is the tie rank in the low bits of the maximum.
// The window is circular (origin `head`) over LP = 32*NW slots so that a shift moves no data; slot p lives at index
// (p % B) * G + p / B, which makes the G lanes of a walker touch G consecutive words (no bank conflicts).
template <int NW, int G>
__device__ __forceinline__ void update_ref(WalkerSmem<NW> &s, int L, int sub, u32 gmask, bool reset, bool rev, int shift, int &head)
{
	constexpr int B = 32 * NW / G, LP = 32 * NW;
	if (reset) head = 0;
	else { head += shift; if (head >= LP) head -= LP; }
	// the 2B bits of the new read that fall on this lane's positions (reverse-complemented first if rev)
	u64 mine;
	if (!rev) mine = getbits_z(s.cur, NW, 2 * B * sub, 2 * B);
	else {
		u64 v = getbits_z(s.cur, NW, 2 * (L - B * sub - B), 2 * B);
		mine = (revpairs64_w(v) >> (64 - 2 * B)) ^ lowmask(2 * B);
	}
	int q = head / B, r = head % B;
	const int fresh = reset ? 0 : L - shift; // positions >= fresh enter the window with this read (reorder.cpp:903-908)
	u64 out = 0;
#pragma unroll
	for (int t = 0; t < B; t++) {
		const int i = B * sub + t;
		const int idx = r * G + ((sub + q) & (G - 1));
		const u32 cc = (u32)(mine >> (2 * t)) & 3u;
		// one vote: only the voted key and the maximum of the position change; a position that enters the window
		// (i >= fresh) starts from the bare tie ranks of A, G, C, T (bit-code order): 3, 1, 2, 0
		const bool inl = i < L, isfresh = i >= fresh;
		u32 k = s.key[cc * LP + idx], b = s.best[idx];
		if (isfresh) { k = (0x27u >> (2 * cc)) & 3u; b = 0u; }
		k += 4u;
		b = max(b, k);
		if (isfresh && inl) { s.key[idx] = 3u; s.key[LP + idx] = 1u; s.key[2 * LP + idx] = 2u; s.key[3 * LP + idx] = 0u; }
		if (inl) { s.key[cc * LP + idx] = k; s.best[idx] = b; }
		else b = 3u; // beyond the read: base A, the zero bits of the reference's bitset
		out |= (u64)((0x27u >> (2 * (b & 3u))) & 3u) << (2 * t); // tie rank 3,2,1,0 -> code A0 C2 G1 T3
		if (++r == B) { r = 0; q++; }
	}
	if ((2 * B) % 8 == 0) {
		unsigned char *rb = reinterpret_cast<unsigned char *>(s.ref) + (2 * B / 8) * sub;
		if (2 * B == 8) *rb = (unsigned char)out;
		else if (2 * B == 16) *reinterpret_cast<unsigned short *>(rb) = (unsigned short)out;
		else if (2 * B == 32) *reinterpret_cast<u32 *>(rb) = (u32)out;
		else if (2 * B == 64) *reinterpret_cast<u64 *>(rb) = out;
		else {
#pragma unroll
			for (int k = 0; k < 2 * B / 8; k++) rb[k] = (unsigned char)(out >> (8 * k));
		}
		__syncwarp(gmask);
	} else { // pieces that are not whole bytes (2B <= 28 bits here) are OR-ed into the zeroed words
		if (sub < NW) s.ref[sub] = 0ull;
		__syncwarp(gmask);
		u32 *r32 = reinterpret_cast<u32 *>(s.ref);
		const int o = 2 * B * sub, w = o >> 5, sh = o & 31;
		atomicOr(&r32[w], (u32)out << sh);
		if (sh + 2 * B > 32) atomicOr(&r32[w + 1], (u32)out >> (32 - sh));
		__syncwarp(gmask);
	}
	if (sub < NW) {
		const int sft = 2 * (32 * NW - L);
		const u64 t0 = revpairs64_w(s.ref[NW - 1 - sub]);
		const u64 t1 = sub + 1 < NW ? revpairs64_w(s.ref[NW - 2 - sub]) : 0ull;
		const u64 x = sft ? (t0 >> sft) | (t1 << (64 - sft)) : t0;
		s.rref[sub] = x ^ lowmask(2 * L - 64 * sub);
	}
	__syncwarp(gmask);
}

// popcount(ref ^ (read & mask[j])) with ref >>= 2j (forward, reorder.cpp:543) or
// popcount(revref ^ (read & revmask[j])) with revref <<= 2j (reverse, reorder.cpp:608).
// One code path for both (forward and reverse lanes sit in the same warp): the window is the consensus moved by
// off = +2j (ref) or -2j (rref) bits with zero fill, compared on the bits [mlo, mhi) that both reads cover.
template <int NW>
__device__ __forceinline__ int hamming(const WalkerSmem<NW> &s, const u64 (&rw)[NW], int L, int j, bool rev)
{
	const u64 *w = rev ? s.rref : s.ref;
	const int off = rev ? -2 * j : 2 * j;
	const int mlo = rev ? 2 * j : 0, mhi = rev ? 2 * L : 2 * (L - j);
	const int q = off >> 6, r = off & 63; // floor division also for negative off
	int d = 0;
	u64 lo = (q >= 0 && q < NW) ? w[q] : 0ull;
#pragma unroll
	for (int k = 0; k < NW; k++) {
		const u64 hi = (k + q + 1 >= 0 && k + q + 1 < NW) ? w[k + q + 1] : 0ull;
		const u64 x = r ? (lo >> r) | (hi << (64 - r)) : lo;
		const u64 m = lowmask(mhi - 64 * k) & ~lowmask(mlo - 64 * k);
		d += __popcll((x ^ rw[k]) & m);
		lo = hi;
	}
	return d;
}

__device__ __forceinline__ u32 *claim_word(const WalkArgs &a, u32 rid)
{
	if (a.world == 1) return a.claim + (rid >> 5);
	const u32 r = rid / a.seg_per;
	return a.seg[r] + ((rid - r * a.seg_per) >> 5);
}
// word to TEST before fetching a candidate: the authoritative word for the own range, the local hint for a peer's range
// (a stale hint only costs a failed remote claim, after which the hint is corrected)
__device__ __forceinline__ const u32 *peek_word(const WalkArgs &a, u32 rid)
{
	if (a.world == 1) return a.claim + (rid >> 5);
	return (rid - a.base < a.n_loc) ? a.claim + ((rid - a.base) >> 5) : a.hint + (rid >> 5);
}
__device__ __forceinline__ bool try_claim(const WalkArgs &a, u32 rid)
{
	u32 bit = 1u << (rid & 31);
	u32 old = atomicAnd(claim_word(a, rid), ~bit);
	if (a.world > 1 && rid - a.base >= a.n_loc) atomicAnd(a.hint + (rid >> 5), ~bit); // claimed now, by this GPU or by a peer
	return (old & bit) != 0;
}

template <int NW, int G>
__global__ void __launch_bounds__(WALK_WARPS * 32, (NW <= 4 ? 8 : 4) / (G == 8 ? 2 : 1)) walk_kernel(WalkArgs a)
{
	constexpr int WPW = 32 / G;  // walkers per warp
	constexpr int SPR = G / 4;   // shifts covered by the lanes of a walker at once
	constexpr int U = 32 / G;    // probes per lane per round: a round covers SPR * U = 8 shifts
	constexpr int UX = G == 32 ? 2 : U; // one walker per warp: after a first round without a match the rounds are twice as wide
	extern __shared__ uint4 smem_raw[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int sub = lane & (G - 1), gbase = lane & ~(G - 1);
	const u32 gmask = (G == 32 ? FULL : ((1u << (G & 31)) - 1u)) << gbase;
	const bool leader = sub == 0;
	const u32 wid = (blockIdx.x * WALK_WARPS + warp) * WPW + lane / G;
	WalkerSmem<NW> &s = reinterpret_cast<WalkerSmem<NW> *>(smem_raw)[warp * WPW + lane / G];
	const int L = a.L;

	u32 c_steps = 0, c_probes = 0, c_hits = 0, c_cmp = 0, c_fail = 0, c_restart = 0;
#ifdef WALK_PROF
	u64 pacc[16];
	for (int i = 0; i < 16; i++) pacc[i] = 0;
	const long long pstart_ = clock64();
#endif

	// forward log (leader only)
	auto emit = [&](u64 rec) {
		if (s.fill == CHUNK) {
			if (s.chunk != NONE) a.chunk_fill[s.chunk] = CHUNK;
			u32 c = atomicAdd(a.chunk_ctr, 1u);
			if (c >= a.max_chunks) c = a.max_chunks - 1; // cannot happen: max_chunks = n/CHUNK + walkers + 1
			a.chunk_key[c] = ((u64)wid << 32) | s.seq++;
			s.chunk = c;
			s.fill = 0;
		}
		a.recs[(size_t)s.chunk * CHUNK + s.fill++] = rec;
	};
	// left log (leader only)
	auto lemit = [&](u64 rec) {
		if (s.lchunk == NONE || s.lfill == CHUNK) {
			u32 c = atomicAdd(a.lchunk_ctr, 1u);
			if (c >= a.max_lchunks) c = a.max_lchunks - 1; // cannot happen: max_lchunks = n/CHUNK + walkers + 1
			a.lprev[c] = s.lchunk;
			if (s.lfirst == NONE) s.lfirst = c;
			s.lchunk = c;
			s.lfill = 0;
		}
		a.lrecs[(size_t)s.lchunk * CHUNK + s.lfill++] = rec;
	};

	int state = S_DONE, head = 0, jb = 0;
	bool left_mode = false, prev_unmatched = false;
	u32 current = 0, prev = 0;
	if (leader) {
		s.chunk = NONE; s.fill = CHUNK; s.seq = 0;
		s.lchunk = NONE; s.lfill = 0; s.lfirst = NONE; s.lcount = 0; s.j1 = 0; s.pend = 0; s.pend_f = 0;
		s.stripe = wid; s.stripes_tried = 0;
		s.cursor = wid < a.walkers ? (long long)((((u64)wid + 1) * a.n_loc) / a.walkers) - 1 : -1;
	}
	__syncwarp();
	if (wid < a.walkers) {
		// reorder.cpp:476-497: walker t starts at read t*(n/T).  (The reference thread gives up if that read is taken;
		// here the walker looks for another head instead.)
		const u32 start = a.base + (u32)((u64)wid * (a.n_loc / a.walkers));
		int ok = leader ? (int)try_claim(a, start) : 0;
		ok = __shfl_sync(gmask, ok, gbase);
		if (ok) { current = start; state = S_NEWHEAD; c_restart += leader; }
		else state = S_RESTART;
	}

	const int kind = sub & 3;   // 0: fwd dict0, 1: fwd dict1, 2: rev dict0, 3: rev dict1 (reorder.cpp:517-643 order)
	const bool rev = kind >= 2;
	const int l = kind & 1;
	const bool dict_on = l < a.numdict;
	const DictView dv = a.d[l];
	const int kb = a.kbits[l];

	while (true) {
		if (!__any_sync(FULL, state != S_DONE)) break;
		PROF_T0();

		// ---- new chain head (reorder.cpp:650-688).  The reference takes the highest unclaimed index through a private
		// downward cursor per thread.  Here the reads are cut into one stripe per walker; a walker scans its own stripe
		// downward first and then the following stripes, so concurrent restarts do not fight over one bit.  With one
		// walker the stripe is the whole array and the choice is exactly the reference's.
		if (__any_sync(FULL, state == S_RESTART)) {
			if (state == S_RESTART) {
				bool got_head = false;
				u32 stripe = s.stripe, tried = s.stripes_tried;
				long long cursor = s.cursor;
				while (tried < a.walkers) {
					const long long slo = (long long)(((u64)stripe * a.n_loc) / a.walkers);
					if (cursor < slo) {
						// stripe exhausted: everything in it is claimed for good
						if (leader) a.stripe_done[stripe] = 1u;
						// move to the next stripe (cyclically) that is not known to be finished, G flags at a time
						bool found_stripe = false;
						while (tried + 1 < a.walkers) {
							const u32 span = min((u32)G, a.walkers - 1 - tried);
							u32 cs = stripe + 1 + sub;
							if (cs >= a.walkers) cs -= a.walkers;
							const bool open_ = (u32)sub < span && ldvol(&a.stripe_done[cs]) == 0u;
							const u32 bal = __ballot_sync(gmask, open_) >> gbase;
							if (bal) {
								const u32 f = __ffs(bal) - 1;
								stripe = stripe + 1 + f;
								if (stripe >= a.walkers) stripe -= a.walkers;
								tried += f + 1;
								found_stripe = true;
								break;
							}
							stripe += span;
							if (stripe >= a.walkers) stripe -= a.walkers;
							tried += span;
						}
						if (!found_stripe) { tried = a.walkers; break; }
						cursor = (long long)((((u64)stripe + 1) * a.n_loc) / a.walkers) - 1;
						continue;
					}
					const long long topw = cursor >> 5;
					const long long wi = topw - sub;
					u32 word = (wi >= 0 && wi >= (slo >> 5)) ? ldvol(&a.claim[wi]) : 0u;
					if (sub == 0) { int bt = (int)(cursor & 31); if (bt != 31) word &= (2u << bt) - 1u; }
					if (wi == (slo >> 5)) word &= ~((1u << (slo & 31)) - 1u);
					const u32 bal = __ballot_sync(gmask, word != 0u) >> gbase;
					if (!bal) { cursor = (topw - (G - 1)) * 32 - 1; continue; }
					const int src = __ffs(bal) - 1;
					const u32 wv = __shfl_sync(gmask, word, gbase + src);
					const int bit = 31 - __clz(wv);
					const u32 j = (u32)((topw - src) * 32 + bit);
					int got = leader ? (int)try_claim(a, a.base + j) : 0;
					got = __shfl_sync(gmask, got, gbase);
					cursor = (long long)j - 1; // j is claimed now, by this walker or by another one
					if (got) { current = a.base + j; got_head = true; break; }
				}
				if (leader) { s.stripe = stripe; s.stripes_tried = tried; s.cursor = cursor; }
				if (got_head) { state = S_NEWHEAD; c_restart += leader; }
				else {
					state = S_DONE;
					if (leader && s.chunk != NONE) a.chunk_fill[s.chunk] = s.fill;
				}
			}
		}

		PROF_ADD(0);
		// ---- start a chain at `current`: the window is the read itself (reorder.cpp:875-883); with the left extension the
		// walk starts on the reverse-complement strand
		if (__any_sync(FULL, state == S_NEWHEAD)) {
			if (state == S_NEWHEAD) {
				__syncwarp(gmask);
				if (sub < NW) s.cur[sub] = __ldg(&a.reads[(size_t)current * NW + sub]);
				__syncwarp(gmask);
				left_mode = a.extend != 0;
				update_ref<NW, G>(s, L, sub, gmask, true, left_mode, 0, head);
				prev = current;
				prev_unmatched = true;
				jb = 0;
				state = S_SEARCH;
			}
		}

		PROF_ADD(1);
		// ---- search round: the four probe kinds of SPR * U consecutive shifts.  Every lane first issues its U slot loads
		// (independent, so their latencies overlap), then the hits are worked off in shift order.
		const bool searching = state == S_SEARCH;
		bool found = false;
		u32 k_rid = 0;
		int k_j = 0, k_rev = 0;
		{
			const int jq = sub >> 2;
			const int ucur = G == 32 ? (jb == 0 ? 1 : 2) : U; // warp-uniform
			u32 blo[UX], bsize[UX];
			{
				u64 keys[UX];
				u32 hs[UX];
				ulonglong2 sl0[UX], sl1[UX];
#pragma unroll
				for (int u = 0; u < UX; u++) {
					const int j = jb + u * SPR + jq;
					const bool valid = u < ucur && searching && dict_on && j < a.maxmatch && (rev ? dv.dstart > j : dv.dend + j < L);
					keys[u] = 0; hs[u] = 0; sl0[u] = sl1[u] = make_ulonglong2(0ull, 0ull);
					if (valid) {
						keys[u] = rev ? getbits(s.rref, NW, 2 * (dv.dstart - j), kb) : getbits(s.ref, NW, 2 * (dv.dstart + j), kb);
						hs[u] = slot_hash(keys[u]) & dv.slot_mask & ~1u;
						sl0[u] = __ldg(&dv.slots[hs[u]]);
						sl1[u] = __ldg(&dv.slots[hs[u] + 1]);
					}
					c_probes += valid;
				}
#pragma unroll
				for (int u = 0; u < UX; u++) { // finish the lookups (rarely more than the one bucket already loaded)
					blo[u] = 0; bsize[u] = 0;
					dict_resolve(dv, keys[u], hs[u], sl0[u], sl1[u], blo[u], bsize[u]);
					c_hits += bsize[u] != 0u;
				}
			}
			c_steps += leader && searching && jb == 0;
			PROF_ADD(2);
			PROF_CNT(8, searching);
			PROF_CNT(9, searching && jb == 0);
#pragma unroll
			for (int u = 0; u < UX; u++) {
				if (!__any_sync(FULL, bsize[u] != 0u && !found)) continue;
				const int j = jb + u * SPR + jq;
				const u32 lo = blo[u], size = bsize[u];
				// candidate scan: entries of the bin not looked at yet (from the tail, reorder.cpp:540), live entries seen so far
				u32 left = found ? 0u : size;
				int seen = 0;
				u32 cand = NONE;
				u64 rw[NW];
				auto advance = [&]() {
					cand = NONE;
					while (left > 0 && seen < a.maxsearch) {
						left--;
						const u32 rid = bin_entry(dv, lo, size, left);
						const u32 cw = ldvol(peek_word(a, rid)); // claim bit and read are fetched together
						load_read<NW>(a.reads, rid, rw);
						if (!((cw >> (rid & 31)) & 1u)) continue; // removed from the bin in the reference (505-514)
						seen++;
						c_cmp++;
						if (hamming<NW>(s, rw, L, j, rev) <= a.thresh) { cand = rid; break; }
					}
				};
				if (left) advance();
				while (true) {
					const u32 ball = __ballot_sync(FULL, cand != NONE);
					if (!ball) break;
					const u32 bal = (ball & gmask) >> gbase;
					if (bal) {
						const int win = __ffs(bal) - 1;
						int got = 0;
						if (sub == win) {
							got = try_claim(a, cand);
							if (!got) c_fail++;
						}
						got = __shfl_sync(gmask, got, gbase + win);
						if (got) {
							found = true;
							k_rid = __shfl_sync(gmask, cand, gbase + win);
							k_j = jb + u * SPR + (win >> 2);
							k_rev = (win & 3) >= 2;
							if (sub == win) {
#pragma unroll
								for (int k = 0; k < NW; k++) s.cur[k] = rw[k];
							}
							cand = NONE;
						} else if (sub == win) advance();
					}
				}
			}
		}

		PROF_ADD(3);
		PROF_CNT(10, found);
		// ---- a read was appended (reorder.cpp:560-578 / 624-641)
		if (__any_sync(FULL, found)) {
			if (found) {
				current = k_rid;
				__syncwarp(gmask);
				update_ref<NW, G>(s, L, sub, gmask, false, k_rev != 0, k_j, head);
				if (leader) {
					if (!left_mode) {
						if (prev_unmatched) emit(mkrec(prev, (u32)L, 0, 0, 0));
						emit(mkrec(current, (u32)k_j, (u32)k_rev, 1, 0));
					} else {
						// left run, found on the reverse-complement strand: in the final order this read precedes the one found
						// before it, which therefore gets this shift; orientations flip
						if (s.lcount == 0) s.j1 = (u32)k_j;
						else lemit(mkrec(s.pend, (u32)k_j, s.pend_f ^ 1u, 1, 0));
						s.pend = current;
						s.pend_f = (u32)k_rev;
						s.lcount++;
					}
				}
				if (!left_mode) prev_unmatched = false;
				jb = 0;
			}
		}
		if (searching && !found) {
			jb += SPR * (G == 32 ? (jb == 0 ? 1 : 2) : U);
			if (jb >= a.maxmatch) state = S_CHAINEND;
		}

		PROF_ADD(4);
		// ---- nothing matches the window any more
		if (__any_sync(FULL, state == S_CHAINEND)) {
			if (state == S_CHAINEND) {
				if (left_mode) {
					// the left run ends: write it in front of the head, last found first, then walk right from the head
					__syncwarp(gmask);
					const u32 k = s.lcount;
					if (k > 0) {
						if (leader) lemit(mkrec(s.pend, (u32)L, s.pend_f ^ 1u, 0, 0)); // leftmost read = head of the chain
						__syncwarp(gmask);
						u32 c = s.lchunk, f = s.lfill, remaining = k;
						while (remaining) {
							const u32 take = min(min(f, (u32)G), remaining);
							u64 r = 0;
							if ((u32)sub < take) r = __ldcg(&a.lrecs[(size_t)c * CHUNK + f - 1 - sub]);
							for (u32 t = 0; t < take; t++) {
								const u64 rr = __shfl_sync(gmask, r, gbase + t);
								if (leader) emit(rr);
							}
							f -= take;
							remaining -= take;
							if (f == 0 && remaining) { c = __ldcg(&a.lprev[c]); f = CHUNK; }
						}
						if (leader) {
							emit(mkrec(prev, s.j1, 0, 1, 0));
							s.lchunk = s.lfirst; s.lfill = 0; s.lcount = 0; // keep one chunk for the next left run
						}
						prev_unmatched = false;
					}
					left_mode = false;
					current = prev;
					__syncwarp(gmask);
					if (sub < NW) s.cur[sub] = __ldg(&a.reads[(size_t)current * NW + sub]);
					__syncwarp(gmask);
					update_ref<NW, G>(s, L, sub, gmask, true, false, 0, head);
					jb = 0;
					state = S_SEARCH;
				} else {
					if (leader && prev_unmatched) emit(mkrec(prev, 0, 0, 0, 1)); // the head stayed alone: singleton (672-684)
					state = S_RESTART;
				}
			}
		}
		PROF_ADD(5);
	}
#ifdef WALK_PROF
	if (lane == 0) {
		for (int i = 0; i < 11; i++) atomicAdd(&g_walk_prof[i], pacc[i]);
		const u64 life = (u64)(clock64() - pstart_);
		atomicAdd(&g_walk_prof[11], life);
		atomicMax(&g_walk_prof[12], life);
		atomicAdd(&g_walk_prof[13], 1ull);
	}
#endif
	// counters
	for (int o = 16; o > 0; o >>= 1) {
		c_steps += __shfl_xor_sync(FULL, c_steps, o);
		c_probes += __shfl_xor_sync(FULL, c_probes, o);
		c_hits += __shfl_xor_sync(FULL, c_hits, o);
		c_cmp += __shfl_xor_sync(FULL, c_cmp, o);
		c_fail += __shfl_xor_sync(FULL, c_fail, o);
		c_restart += __shfl_xor_sync(FULL, c_restart, o);
	}
	if (lane == 0) {
		atomicAdd(&a.counters[0], (u64)c_steps); atomicAdd(&a.counters[1], (u64)c_probes); atomicAdd(&a.counters[2], (u64)c_hits);
		atomicAdd(&a.counters[3], (u64)c_cmp); atomicAdd(&a.counters[4], (u64)c_fail); atomicAdd(&a.counters[5], (u64)c_restart);
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

template <int NW, int G>
int launch_walk(harcgpu_ctx *c, const WalkArgs &a)
{
	constexpr int WPB = WALK_WARPS * 32 / G; // walkers per block
	size_t smem = sizeof(WalkerSmem<NW>) * WPB;
	CK(cudaFuncSetAttribute(walk_kernel<NW, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	walk_kernel<NW, G><<<KL + cdiv(a.walkers, WPB), WALK_WARPS * 32, smem, c->st>>>(a);
	CK(cudaGetLastError());
	return 0;
}

// walkers that can be resident at once: a walker that is not resident only starts after the others have finished
template <int NW, int G>
int resident_walkers(harcgpu_ctx *c, u32 *out)
{
	constexpr int WPB = WALK_WARPS * 32 / G;
	size_t smem = sizeof(WalkerSmem<NW>) * WPB;
	int nb = 0, sms = 0;
	CK(cudaFuncSetAttribute(walk_kernel<NW, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, walk_kernel<NW, G>, WALK_WARPS * 32, smem));
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
	*out = (u32)nb * (u32)sms * WPB;
	return 0;
}
} // namespace

#define DISPATCH_NW_G(NWv, Gv, CALL)                                                                \
	switch ((NWv) * 100 + (Gv)) {                                                                   \
	case 108: { constexpr int NW = 1, G = 8; CALL; } break;                                          \
	case 208: { constexpr int NW = 2, G = 8; CALL; } break;                                          \
	case 308: { constexpr int NW = 3, G = 8; CALL; } break;                                          \
	case 408: { constexpr int NW = 4, G = 8; CALL; } break;                                          \
	case 508: { constexpr int NW = 5, G = 8; CALL; } break;                                          \
	case 608: { constexpr int NW = 6, G = 8; CALL; } break;                                          \
	case 708: { constexpr int NW = 7, G = 8; CALL; } break;                                          \
	case 808: { constexpr int NW = 8, G = 8; CALL; } break;                                          \
	case 116: { constexpr int NW = 1, G = 16; CALL; } break;                                         \
	case 216: { constexpr int NW = 2, G = 16; CALL; } break;                                         \
	case 316: { constexpr int NW = 3, G = 16; CALL; } break;                                         \
	case 416: { constexpr int NW = 4, G = 16; CALL; } break;                                         \
	case 516: { constexpr int NW = 5, G = 16; CALL; } break;                                         \
	case 616: { constexpr int NW = 6, G = 16; CALL; } break;                                         \
	case 716: { constexpr int NW = 7, G = 16; CALL; } break;                                         \
	case 816: { constexpr int NW = 8, G = 16; CALL; } break;                                         \
	case 132: { constexpr int NW = 1, G = 32; CALL; } break;                                         \
	case 232: { constexpr int NW = 2, G = 32; CALL; } break;                                         \
	case 332: { constexpr int NW = 3, G = 32; CALL; } break;                                         \
	case 432: { constexpr int NW = 4, G = 32; CALL; } break;                                         \
	case 532: { constexpr int NW = 5, G = 32; CALL; } break;                                         \
	case 632: { constexpr int NW = 6, G = 32; CALL; } break;                                         \
	case 732: { constexpr int NW = 7, G = 32; CALL; } break;                                         \
	case 832: { constexpr int NW = 8, G = 32; CALL; } break;                                         \
	default: harcgpu_set_error("unsupported read length %d / lanes per walker %d", c->L, (int)(Gv)); return -1; \
	}

int s1_init_claim(harcgpu_ctx *c, u32 *claim, u32 n)
{
	if (!n) return 0;
	init_claim_kernel<<<KL + cdiv((n + 31) / 32, 256), 256, 0, c->st>>>(claim, n);
	CK(cudaGetLastError());
	return 0;
}

int s1_reorder(harcgpu_ctx *c)
{
	cudaStream_t st = c->st;
	const u32 n = c->n;
	c->reordered = false;
	c->stream_set = false; c->pool_set = false; c->encoded = false; // stage II inputs derived from an earlier pass are stale now
	c->release(c->order); c->release(c->order_s); c->release(c->rev); c->release(c->flag); c->release(c->pos);
	c->order = c->order_s = nullptr; c->rev = c->flag = c->pos = nullptr;
	c->n_matched = c->n_single = c->n_unmatched = 0;
	if (c->alloc(&c->order, n) || c->alloc(&c->order_s, n) || c->alloc(&c->rev, n) || c->alloc(&c->flag, n) || c->alloc(&c->pos, n))
		return -1;
	CK(cudaMemsetAsync(c->counters, 0, 8 * sizeof(u64), st));
	if (n == 0) { c->reordered = true; c->ms["walk"] = 0; c->ms["finalize"] = 0; return 0; }

	// walkers: the reference's num_thr.  Auto: one walker per reads_per_walker reads (every extra walker costs chain
	// heads), capped at what is resident at once.
	u32 resident = 0;
	int rc = -1;
	int lanes = c->p.lanes_per_walker;
	if (const char *e = getenv("HARCGPU_LANES")) lanes = atoi(e); // tuning aid
	if (lanes == 0) lanes = 32;
	if (lanes != 8 && lanes != 16 && lanes != 32) { harcgpu_set_error("lanes_per_walker must be 8, 16 or 32"); return -1; }
	DISPATCH_NW_G(c->NW, lanes, (rc = resident_walkers<NW, G>(c, &resident)));
	if (rc) return rc;
	// one job on several GPUs: this GPU's walkers own the id range [base, base + n_loc) for starts and restarts
	const bool sharded = c->shard_world > 1;
	if (sharded && (c->shard_n != n || !c->shard_ready)) {
		harcgpu_set_error("sharded reorder: call harcgpu_shard_init/connect for these %u reads and harcgpu_shard_reset before every pass", n);
		return -1;
	}
	c->shard_ready = false; // a sharded pass consumes the reset
	const u32 base = sharded ? std::min<u64>((u64)c->shard_rank * c->seg_per, n) : 0u;
	const u32 n_loc = sharded ? (u32)(std::min<u64>(((u64)c->shard_rank + 1) * c->seg_per, n) - base) : n;
	const u32 per = c->p.reads_per_walker > 0 ? (u32)c->p.reads_per_walker : 4096u;
	u32 walkers = c->p.walkers > 0 ? (u32)c->p.walkers : (u32)std::min<u64>(resident, std::max<u64>(1, n_loc / per));
	if (walkers > n_loc) walkers = n_loc;
	if (walkers == 0) { c->reordered = true; c->ms["walk"] = 0; c->ms["finalize"] = 0; return 0; } // no read in this GPU's range
	c->walkers_used = walkers;
	// left extension: off for a single walker unless asked for (one walker without it = the reference at num_thr=1)
	const int extend = c->p.extend > 0 ? 1 : (c->p.extend < 0 ? 0 : (walkers > 1 ? 1 : 0));

	u32 max_chunks = n / CHUNK + walkers + 1;
	u64 *recs = nullptr, *chunk_key = nullptr, *key_sorted = nullptr, *scan_tmp = nullptr, *lrecs = nullptr;
	u32 *chunk_fill = nullptr, *ctrs = nullptr, *chunk_id = nullptr, *chunk_sorted = nullptr, *cm = nullptr, *cs = nullptr,
	    *om = nullptr, *os = nullptr, *totals = nullptr, *lprev = nullptr;
	if (c->alloc(&recs, (size_t)max_chunks * CHUNK) || c->alloc(&chunk_key, max_chunks) || c->alloc(&chunk_fill, max_chunks) ||
	    c->alloc(&ctrs, 2))
		return -1;
	if (extend && (c->alloc(&lrecs, (size_t)max_chunks * CHUNK) || c->alloc(&lprev, max_chunks))) return -1;
	CK(cudaMemsetAsync(ctrs, 0, 8, st));
	CK(cudaMemsetAsync(chunk_fill, 0, 4 * (size_t)max_chunks, st));
	// one GPU: the bitmap.  Sharded: the authoritative ranges are armed by harcgpu_shard_reset before the barrier that
	// precedes the walk; c->claim serves as this GPU's hint bitmap for the peers' ranges.
	init_claim_kernel<<<KL + cdiv((n + 31) / 32, 256), 256, 0, st>>>(c->claim, n);
	CK(cudaGetLastError());
	u32 *stripe_done = nullptr;
	if (c->alloc(&stripe_done, walkers)) return -1;
	CK(cudaMemsetAsync(stripe_done, 0, 4 * (size_t)walkers, st));

	WalkArgs a;
	a.reads = c->reads; a.n = n; a.L = c->L; a.maxmatch = c->p.maxmatch; a.thresh = c->p.thresh; a.maxsearch = c->p.maxsearch;
	a.numdict = c->p.numdict; a.extend = extend;
	for (int l = 0; l < 2; l++) {
		int ll = l < c->p.numdict ? l : 0;
		a.d[l].slots = c->d1[ll].slots; a.d[l].ids = c->d1[ll].ids; a.d[l].slot_mask = c->d1[ll].slot_mask;
		a.d[l].dstart = c->p.dict_start[ll]; a.d[l].dend = c->p.dict_end[ll];
		a.kbits[l] = c->d1[ll].nbits;
	}
	a.claim = sharded ? c->seg[c->shard_rank] : c->claim; a.hint = c->claim; a.stripe_done = stripe_done; a.walkers = walkers;
	for (int r = 0; r < 8; r++) a.seg[r] = sharded && r < c->shard_world ? c->seg[r] : nullptr;
	a.seg_per = sharded ? c->seg_per : 0u; a.base = base; a.n_loc = n_loc; a.world = sharded ? c->shard_world : 1;
	a.recs = recs; a.chunk_key = chunk_key; a.chunk_fill = chunk_fill; a.chunk_ctr = ctrs; a.max_chunks = max_chunks;
	a.lrecs = lrecs; a.lprev = lprev; a.lchunk_ctr = ctrs + 1; a.max_lchunks = max_chunks;
	a.counters = c->counters;
	c->tic();
	DISPATCH_NW_G(c->NW, lanes, (rc = launch_walk<NW, G>(c, a)));
	if (rc) return rc;
	c->toc("walk");
	CK(cudaGetLastError());
#ifdef WALK_PROF
	{
		unsigned long long h[16], z[16] = { 0 };
		CK(cudaStreamSynchronize(st));
		CK(cudaMemcpyFromSymbol(h, g_walk_prof, sizeof h));
		CK(cudaMemcpyToSymbol(g_walk_prof, z, sizeof z));
		fprintf(stderr, "WALK_PROF walkers %u cycles: restart %llu newhead %llu probe %llu cand %llu append %llu chainend %llu | rounds %llu steps %llu found %llu | warp life avg %llu max %llu warps %llu\n",
		        walkers, h[0], h[1], h[2], h[3], h[4], h[5], h[8], h[9], h[10], h[13] ? h[11] / h[13] : 0ull, h[12], h[13]);
	}
#endif

	// ---- finalize
	c->tic();
	u32 nchunks = 0;
	CK(cudaMemcpyAsync(&nchunks, ctrs, 4, cudaMemcpyDeviceToHost, st));
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
	void *tmp[] = { recs, chunk_key, key_sorted, chunk_fill, ctrs, chunk_id, chunk_sorted, cm, cs, om, os, totals, scan_tmp, cub_tmp,
	                stripe_done, lrecs, lprev };
	for (void *q : tmp) c->release(q);
	if (!sharded && (u64)tot[0] + tot[1] != n) { harcgpu_set_error("reorder lost reads: %u matched + %u singletons != %u", tot[0], tot[1], n); return -1; }
	c->reordered = true;
	return 0;
}
