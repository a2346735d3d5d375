// Stage I chain walk of HARC (reference: reorder.cpp:434-703 reorder(), 863-915 updaterefcount) as one sm_100a kernel.
//
// A walker is the reference's OpenMP thread: it follows one chain of overlapping reads at a time.  Here a walker is one
// warp, thousands per GPU.  Per round the 32 lanes issue the reference's four probe kinds (forward dict 0, forward
// dict 1, reverse dict 0, reverse dict 1) for eight consecutive shifts, lane = 4 * shift + kind; the lowest lane with a
// claimable candidate wins, which is exactly the sequential order of reorder.cpp:517-649, so one walker reproduces the
// reference at num_thr=1 byte for byte.  Reads are claimed with one atomicAnd on a bitmap (instead of the reference's
// 2 x 2^24 striped locks and its in-place bin compaction); the consensus window of updaterefcount lives in shared
// memory as a circular array of vote keys.  Everything on the hot path works on 32-bit words (funnel shifts over
// zero-padded shared-memory arrays, compare masks from a per-block table), because the kernel is bound by instruction
// issue and dependent latency, not by bandwidth (profiles/).  A blocked Bloom filter over the keys of both dictionaries,
// small enough to stay in L2, is asked before the key tables (three of four probes are for keys in no dictionary), and
// bins of many reads (repeats) are scanned through a cursor cache instead of from their used-up tail.
//
// One job on several GPUs (job.cu): the same kernel; a probe goes to the key's owner over NVLink after the local filter
// said "maybe", a claim to the owner of the read's range of the bitmap, the candidate read comes from the local replica.
//
// Not in the reference (switchable, params.extend): a new chain is first extended to the LEFT of its head by walking the
// reverse-complement strand, and that run is written in front of the head in reverse order.  The streams only encode
// order, orientation and shifts, so the reference's decoders read the result unchanged; it removes
// the orphaned left halves that every restart otherwise leaves behind, which is what lets thousands of concurrent
// walkers stay inside the 2 % size budget.
#include "ctx.h"
#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace {
constexpr int WALK_WARPS = 4;         // warps (= walkers) per block
constexpr int CHUNK = 32;             // records per log chunk
constexpr u32 NONE = 0xffffffffu;
constexpr u32 FULL = 0xffffffffu;
constexpr int DRY_HEADS = 4;          // left extension is skipped after this many heads in a row that stayed alone
constexpr int STREAK_HEADS = 16;      // after this many, a head is searched over STREAK_SHIFTS shifts only
constexpr int STREAK_SHIFTS = 8;
constexpr u32 BIG_BIN = 64;           // bins of at least this many reads go through the cursor cache (advance())
constexpr u32 TAILC_ENTRIES = 1u << 16;

enum { S_SEARCH = 0, S_CHAINEND, S_RESTART, S_NEWHEAD, S_DONE };
enum { P_DEAD = 0, P_PENDING, P_BIN, P_CAND };

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
	// cursor cache of the big bins: entry = bin tag | index below which the bin still has unclaimed reads (see advance())
	unsigned long long *tailc;
	u32 tailc_mask;
	const u32 *bloom;      // sharded dictionaries: this GPU's copy of the Bloom filter over all shards (null: no filter)
	u32 bloom_seg_words;
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

// Per-walker shared memory.  ref / rref / cur are arrays of 32-bit words with zero words around them, so that a window
// moved by any shift is read without bounds checks: forward windows read up to NW words above ref, reverse windows up
// to NW words below rref (the same zeros), key extraction up to two words above either.
template <int NW>
struct alignas(16) WalkerSmem {
	uint4 key[32 * NW];   // vote keys (count << 2 | tie rank) of the four base codes per window position, circular (update_ref)
	long long cursor;     // restart: downward cursor inside the current stripe
	u32 stripe, stripes_tried;
	u32 chunk, fill, seq;                      // forward log
	u32 lchunk, lfill, lfirst, lcount, j1;     // left log
	u32 pend, pend_f;                          // left run: read found last, not yet written
	u32 cnt[4];                                // counters that change rarely live here, not in registers: steps, restarts, harvested, lost claims
	u32 ref[2 * NW];      // consensus of the current window, 2 bits/base (reorder.cpp:466)
	u32 zpad[NW];
	u32 rref[2 * NW];     // its reverse complement
	u32 zpad2[2];
	u32 cur[2 * NW];      // the read just appended
	u32 zpad3[2];
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
__device__ __forceinline__ void load_read(const u64 *__restrict__ reads, u32 rid, u32 (&rw)[2 * NW])
{
	const u64 *r = reads + (size_t)rid * NW;
	if (NW % 2 == 0) {
		const uint4 *r4 = reinterpret_cast<const uint4 *>(r);
#pragma unroll
		for (int k = 0; k < NW / 2; k++) { uint4 v = __ldg(&r4[k]); rw[4 * k] = v.x; rw[4 * k + 1] = v.y; rw[4 * k + 2] = v.z; rw[4 * k + 3] = v.w; }
	} else {
		const uint2 *r2 = reinterpret_cast<const uint2 *>(r);
#pragma unroll
		for (int k = 0; k < NW; k++) { uint2 v = __ldg(&r2[k]); rw[2 * k] = v.x; rw[2 * k + 1] = v.y; }
	}
}

__device__ __forceinline__ u32 revpairs32(u32 x) // reverse the order of the 16 base pairs of a word
{
	const u32 y = __brev(x);
	return ((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u);
}
__device__ __forceinline__ u32 lowmask32(int n) { return n >= 32 ? ~0u : (n <= 0 ? 0u : ((1u << n) - 1u)); }

// updaterefcount (reorder.cpp:863-915) for one walker, by the G lanes of its group.
// Lane `lane` owns the B = 32*NW/G consecutive window positions [B*lane, B*lane+B).  The votes of a position are four keys
// (count << 2 | tie rank), one per base code (bit-code order A, G, C, T), in one uint4: a read adds one vote per
// position, and argmax with ties -> A < C < G < T (strict '>' from max = 0, reorder.cpp:893-899) is the tie rank in
// the low bits of the largest key.  The window is circular (origin `head`) over LP = 32*NW slots so that a shift
// moves no data; slot p lives at index (p % B) * G + p / B, which makes the lanes touch G consecutive uint4.
template <int NW, int G>
__device__ __forceinline__ void update_ref(WalkerSmem<NW> &s, int L, int lane, u32 gmask, bool reset, bool rev, int shift, int &head,
                                           bool final = true)
{
	constexpr int B = 32 * NW / G, LP = 32 * NW, W2 = 2 * NW; // `lane` = lane inside the walker's group of G
	if (reset) head = 0;
	else { head += shift; if (head >= LP) head -= LP; }
	// the 2B bits of the new read that fall on this lane's positions (reverse-complemented first if rev)
	u32 mine;
	{
		const int pos = rev ? 2 * (L - B * lane - B) : 2 * B * lane; // >= -64: the words below cur are padding
		const int q = pos >> 5, r = pos & 31;
		const u32 v = __funnelshift_r(s.cur[q], s.cur[q + 1], r);
		mine = rev ? (revpairs32(v << (32 - 2 * B)) ^ lowmask32(2 * B)) : v;
	}
	int q = head / B, r = head % B;
	const int fresh = reset ? 0 : L - shift; // positions >= fresh enter the window with this read (reorder.cpp:903-908)
	u32 out = 0;
#pragma unroll
	for (int t = 0; t < B; t++) {
		const int i = B * lane + t;
		const int idx = r * G + ((lane + q) & (G - 1));
		const u32 cc = (mine >> (2 * t)) & 3u;
		uint4 v = s.key[idx];
		// a position that enters the window starts from the bare tie ranks of A, G, C, T (bit-code order): 3, 1, 2, 0
		if (i >= fresh) v = make_uint4(3u, 1u, 2u, 0u);
		v.x += cc == 0u ? 4u : 0u;
		v.y += cc == 1u ? 4u : 0u;
		v.z += cc == 2u ? 4u : 0u;
		v.w += cc == 3u ? 4u : 0u;
		s.key[idx] = v;
		if (++r == B) { r = 0; q++; }
		if (!final) continue; // more reads of this round follow: only the votes, the consensus is taken with the last one
		u32 b = max(max(v.x, v.y), max(v.z, v.w));
		// positions beyond the read (i >= L) live in slots that no valid position uses (LP >= L) and that are re-initialised
		// when they enter the window, so the store needs no guard; their base is A, the zero bits of the reference's bitset
		if (i >= L) b = 3u;
		out |= ((0x27u >> (2 * (b & 3u))) & 3u) << (2 * t); // tie rank 3,2,1,0 -> code A0 C2 G1 T3
	}
	if (!final) { __syncwarp(gmask); return; }
	if ((2 * B) % 8 == 0) {
		unsigned char *rb = reinterpret_cast<unsigned char *>(s.ref) + (2 * B / 8) * lane;
		if (2 * B == 8) *rb = (unsigned char)out;
		else if (2 * B == 16) *reinterpret_cast<unsigned short *>(rb) = (unsigned short)out;
		else if (2 * B == 32) *reinterpret_cast<u32 *>(rb) = out;
		else {
#pragma unroll
			for (int k = 0; k < 2 * B / 8; k++) rb[k] = (unsigned char)(out >> (8 * k));
		}
		__syncwarp(gmask);
	} else { // pieces that are not whole bytes (2B <= 28 bits here) are OR-ed into the zeroed words
		if (lane < W2) s.ref[lane] = 0u;
		__syncwarp(gmask);
		const int o = 2 * B * lane, w = o >> 5, sh = o & 31;
		atomicOr(&s.ref[w], out << sh);
		if (sh + 2 * B > 32) atomicOr(&s.ref[w + 1], out >> (32 - sh));
		__syncwarp(gmask);
	}
	if (lane < W2) {
		// rref = (ref with its 32*NW base pairs in reverse order) >> (64*NW - 2L), valid bits complemented
		const int sft = 64 * NW - 2 * L, so = sft >> 5, sr = sft & 31;
		const int i0 = W2 - 1 - (lane + so), i1 = i0 - 1;
		const u32 t0 = i0 >= 0 ? revpairs32(s.ref[i0]) : 0u;
		const u32 t1 = i1 >= 0 ? revpairs32(s.ref[i1]) : 0u;
		s.rref[lane] = __funnelshift_r(t0, t1, sr) ^ lowmask32(2 * L - 32 * lane);
	}
	__syncwarp(gmask);
}

// popcount(ref ^ (read & mask[j])) with ref >>= 2j (forward, reorder.cpp:543) or
// popcount(revref ^ (read & revmask[j])) with revref <<= 2j (reverse, reorder.cpp:608).
// One code path for both: `w` points at the word of ref (rref) that holds bit +2j (-2j, in the zero padding), r is
// that offset modulo 32, and `m` is the row of the block's mask table with the bits both reads cover.
template <int NW>
__device__ __forceinline__ int hamming(const u32 *w, int r, const u32 *m, const u32 (&rw)[2 * NW])
{
	u32 mk[2 * NW];
	if (NW % 2 == 0) {
#pragma unroll
		for (int k = 0; k < NW / 2; k++) {
			const uint4 v = reinterpret_cast<const uint4 *>(m)[k];
			mk[4 * k] = v.x; mk[4 * k + 1] = v.y; mk[4 * k + 2] = v.z; mk[4 * k + 3] = v.w;
		}
	} else {
#pragma unroll
		for (int k = 0; k < NW; k++) { const uint2 v = reinterpret_cast<const uint2 *>(m)[k]; mk[2 * k] = v.x; mk[2 * k + 1] = v.y; }
	}
	int d = 0;
	u32 lo = w[0];
#pragma unroll
	for (int k = 0; k < 2 * NW; k++) {
		const u32 hi = w[k + 1];
		d += __popc((__funnelshift_r(lo, hi, r) ^ rw[k]) & mk[k]);
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

// One dictionary probe in flight: the key and the bucket of two slots being looked at.
struct Probe {
	u64 key;
	u32 h;
	ulonglong2 s0, s1;
	bool pend, home;
};

// G = lanes per walker: 32 (one walker per warp, 8 shifts per round) or 16 (two walkers per warp, 4 shifts per round;
// the two walkers share every instruction of a round while they are in the same phase).  `sub` is the lane inside the
// walker's group; every warp-level primitive below is restricted to the group (gmask).
#ifndef WALK_MB
#define WALK_MB 8 // blocks per SM the registers are bounded for when NW <= 4 (8 -> 64 registers)
#endif
template <int NW, int G>
__global__ void __launch_bounds__(WALK_WARPS * 32, NW <= 4 ? WALK_MB : 4) walk_kernel(WalkArgs a)
{
	constexpr int W2 = 2 * NW;
	constexpr int WPW = 32 / G;          // walkers per warp
	constexpr int SPR = G / 4;           // shifts probed per round
	constexpr int SPR_LOG = G == 32 ? 3 : 2;
	extern __shared__ uint4 smem_raw[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int sub = lane & (G - 1), gbase = lane & ~(G - 1);
	const u32 gmask = G == 32 ? FULL : (0xffffu << gbase);
	const bool leader = sub == 0;
	const u32 wid = (blockIdx.x * WALK_WARPS + warp) * WPW + lane / G;
	WalkerSmem<NW> &s = reinterpret_cast<WalkerSmem<NW> *>(smem_raw)[warp * WPW + lane / G];
	// compare masks, one row of W2 words per (direction, shift): forward bits [0, 2(L-j)), reverse bits [2j, 2L)
	u32 *mtab = reinterpret_cast<u32 *>(smem_raw) + WALK_WARPS * WPW * (sizeof(WalkerSmem<NW>) / 4);
	const int L = a.L;
	for (int i = threadIdx.x; i < 2 * a.maxmatch * W2; i += WALK_WARPS * 32) {
		const int row = i / W2, k = i % W2, rv = row >= a.maxmatch, j = rv ? row - a.maxmatch : row;
		const int mlo = rv ? 2 * j : 0, mhi = rv ? 2 * L : 2 * (L - j);
		mtab[i] = lowmask32(mhi - 32 * k) & ~lowmask32(mlo - 32 * k);
	}
	if (sub < NW) s.zpad[sub] = 0u;
	if (sub < 2) { s.zpad2[sub] = 0u; s.zpad3[sub] = 0u; }
	__syncthreads();

	u32 c_probes = 0, c_hits = 0, c_cmp = 0;
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
			if (c >= a.max_chunks) c = a.max_chunks - 1; // overflow (the host sees the counter and fails): keep the stores in bounds
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
			if (c >= a.max_lchunks) c = a.max_lchunks - 1; // overflow (the host sees the counter and fails): keep the stores in bounds
			a.lprev[c] = s.lchunk;
			if (s.lfirst == NONE) s.lfirst = c;
			s.lchunk = c;
			s.lfill = 0;
		}
		a.lrecs[(size_t)s.lchunk * CHUNK + s.lfill++] = rec;
	};

	// walker state: identical in all lanes of the walker's group (every update comes from a broadcast value)
	int state = S_DONE, head = 0, jb = 0, dry = 0;
	bool left_mode = false, prev_unmatched = false;
	u32 current = 0, prev = 0;
	if (leader) {
		s.chunk = NONE; s.fill = CHUNK; s.seq = 0;
		s.lchunk = NONE; s.lfill = 0; s.lfirst = NONE; s.lcount = 0; s.j1 = 0; s.pend = 0; s.pend_f = 0;
		s.stripe = wid; s.stripes_tried = 0;
		s.cnt[0] = s.cnt[1] = s.cnt[2] = s.cnt[3] = 0u;
		s.cursor = wid < a.walkers ? (long long)((((u64)wid + 1) * a.n_loc) / a.walkers) - 1 : -1;
	}
	__syncwarp(gmask);
	if (wid < a.walkers) {
		// reorder.cpp:476-497: walker t starts at read t*(n/T).  (The reference thread gives up if that read is taken;
		// here the walker looks for another head instead.)
		const u32 start = a.base + (u32)((u64)wid * (a.n_loc / a.walkers));
		int ok = leader ? (int)try_claim(a, start) : 0;
		ok = __shfl_sync(gmask, ok, gbase);
		if (ok) { current = start; state = S_NEWHEAD; if (leader) s.cnt[1]++; }
		else state = S_RESTART;
	}

	// lane constants: probe kind and shift inside a round
	const int kind = sub & 3;    // 0: fwd dict0, 1: fwd dict1, 2: rev dict0, 3: rev dict1 (reorder.cpp:517-643 order)
	const int jq = sub >> 2;
	const bool rev = kind >= 2;
	const int l = kind & 1;
	const DictView dv = a.d[l];
	const bool full64 = a.kbits[0] == 64 && a.kbits[a.numdict - 1] == 64; // warp-uniform: keys need no masking
	const int kb = a.kbits[l];
	const int step2 = rev ? -2 : 2;       // a shift by one base moves this lane's windows by step2 bits
	const u32 *wbase = rev ? s.rref : s.ref;
	const u32 *mbase = mtab + (rev ? a.maxmatch * W2 : 0);
	// bit r of vmask: this lane's probe exists in the round that starts at shift SPR*r (reorder.cpp:520-523, 585-588)
	u32 vmask = 0;
	for (int r = 0; SPR * r < a.maxmatch; r++) {
		const int j = SPR * r + jq;
		if (l < a.numdict && j < a.maxmatch && (rev ? dv.dstart > j : dv.dend + j < L)) vmask |= 1u << r;
	}

	// issue the probe of this lane for shift j: key = the dictionary window of the consensus moved by j
	auto probe_key = [&](int j, u64 &key, u32 &h) {
		const int koff = 2 * dv.dstart + step2 * j;
		const u32 *w = wbase + (koff >> 5);
		const int r = koff & 31;
		const u32 w0 = w[0], w1 = w[1], w2 = w[2];
		u32 klo = __funnelshift_r(w0, w1, r), khi = __funnelshift_r(w1, w2, r);
		if (!full64) { klo &= lowmask32(kb); khi &= lowmask32(kb - 32); }
		key = key_mix(((u64)khi << 32) | klo); // from here on the mixed key stands for the key (common.cuh)
		h = slot_home(key, dv.slot_shift, dv.world);
	};
	// sharded dictionaries (one job on several GPUs): table and id lists of the shard that owns the key, in the owner's
	// HBM and read over NVLink; dv.world == 0 (warp-uniform) otherwise
	// (the per-shard pointers are indexed in the kernel parameters, not in a local copy)
	auto slots_of = [&](u64 t) -> const ulonglong2 * { return dv.world ? a.d[l].sslots[mix_shard(t, dv.world)] : dv.slots; };
	auto ids_of = [&](u64 t) -> const u32 * { return dv.world ? a.d[l].sids[mix_shard(t, dv.world)] : dv.ids; };
	auto issue = [&](int j, Probe &p) {
		p.pend = (vmask >> (j >> SPR_LOG)) & 1u;
		p.home = true;
		if (p.pend) {
			probe_key(j, p.key, p.h);
			c_probes++;
			if (a.bloom) {
				// most window keys are in no dictionary: ask the filter first (one job on several GPUs: the table of the key lives
				// on another GPU world - 1 times out of world, and then nothing crosses NVLink; one GPU: the filter sits in L2
				// and the tables in HBM)
				u32 bw, bb;
				job_bloom_pos(p.key, l, dv.world ? dv.world : 1, a.bloom_seg_words, bw, bb);
				if ((__ldg(&a.bloom[bw]) & bb) != bb) { p.pend = false; return; }
			}
			const ulonglong2 *sl = slots_of(p.key);
			p.s0 = __ldg(&sl[p.h]);
			p.s1 = __ldg(&sl[p.h + 1]);
		}
	};

	// L2 prefetch of the bucket that the probe for shift j will read (used one round ahead in a fruitless search)
	auto prefetch = [&](int j) {
		if (!dv.world && !a.bloom && ((vmask >> (j >> SPR_LOG)) & 1u)) {
			u64 key;
			u32 h;
			probe_key(j, key, h);
			asm volatile("prefetch.global.L2 [%0];" ::"l"(&dv.slots[h]));
		}
	};
	Probe pc;
	pc.pend = false;

	while (__any_sync(FULL, state != S_DONE)) {
		PROF_T0();
		// ---- new chain head (reorder.cpp:650-688).  The reference takes the highest unclaimed index through a private
		// downward cursor per thread.  Here the reads are cut into one stripe per walker; a walker scans its own stripe
		// downward first and then the following stripes, so concurrent restarts do not fight over one bit.  With one
		// walker the stripe is the whole array and the choice is exactly the reference's.
		if (state == S_RESTART) {
			bool got_head = false;
			__syncwarp(gmask); // the leader's stores of the last restart are ordered before these loads (racecheck)
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
			if (got_head) { state = S_NEWHEAD; if (leader) s.cnt[1]++; }
			else {
				state = S_DONE;
				if (leader && s.chunk != NONE) a.chunk_fill[s.chunk] = s.fill;
			}
		}
		PROF_ADD(0);

		// ---- start a chain at `current`: the window is the read itself (reorder.cpp:875-883); with the left extension the
		// walk starts on the reverse-complement strand.  A walker whose last heads all stayed alone (the tail of the job,
		// where the unclaimed reads are the ones no dictionary window finds) goes right only.
		if (state == S_NEWHEAD) {
			__syncwarp(gmask);
			if (sub < W2) s.cur[sub] = __ldg(reinterpret_cast<const u32 *>(a.reads + (size_t)current * NW) + sub);
			__syncwarp(gmask);
			left_mode = a.extend != 0 && dry < DRY_HEADS;
			update_ref<NW, G>(s, L, sub, gmask, true, left_mode, 0, head);
			prev = current;
			prev_unmatched = true;
			jb = 0;
			state = S_SEARCH;
		}
		PROF_ADD(1);

		// ---- search round: the four probe kinds of SPR consecutive shifts from jb on, one probe per lane.  After a round
		// without a match the buckets of the following round are prefetched into L2 while this round is worked off.
		bool found = false;
		u32 k_rid = 0;
		int k_j = 0, k_rev = 0;
		if (state == S_SEARCH) {
			const int j = jb + jq;
			issue(j, pc);
			if (jb > 0) prefetch(j + SPR);
			if (leader && jb == 0) s.cnt[0]++;
			PROF_CNT(8, 1);
			PROF_CNT(9, jb == 0);

			// The lanes are worked off in lane order = the sequential order of the reference.  A lane is PENDING while its
			// bucket chain is not resolved, then owns a BIN whose entries it tests from the tail (reorder.cpp:540), and is a
			// CAND once an entry passed the Hamming test.  The lowest CAND may claim as soon as no PENDING lane precedes it;
			// lanes behind it are never waited for.
			int ps = pc.pend ? P_PENDING : P_DEAD;
			u32 lo = 0, size = 0, left = 0, cand = NONE;
			int seen = 0;
			u32 rw[W2];
			const int off = step2 * j;
			const u32 *hw = wbase + (off >> 5);
			const int hr = off & 31;
			const u32 *hm = mbase + j * W2;
			// next entry of this lane's bin that is unclaimed and within the Hamming threshold (scanned from the tail)
			auto advance = [&](const u32 *hw, int hr, const u32 *hm) { // window word, bit offset and mask row of the lane's shift
				ps = P_DEAD;
				if (size >= BIG_BIN) {
					// A bin of many reads (repeats: poly-A, tandem repeats, transposons).  The reference compacts a bin when a read
					// is removed (reorder.cpp:409-432), so its scans only meet live ids; here claimed reads stay in the list and
					// pile up at the tail, where every scan starts.  A small direct-mapped cache keeps, per big bin, an index above
					// which every entry is known to be claimed (claims are never undone during a walk, so any stored value stays
					// true; a lost or overwritten entry only costs time).  The claim bit is tested before the read is fetched.
					const u32 shard = dv.world ? mix_shard(pc.key, dv.world) : 0u;
					const u64 tag = ((u64)lo << 32) | ((u64)l << 31) | ((u64)shard << 28);
					unsigned long long *ce = a.tailc + ((lo * 2u + (u32)l + shard * 0x9E3779B1u) & a.tailc_mask);
					const bool cacheable = size < (1u << 28);
					if (cacheable && left == size) { // first look at this bin in this round
						const u64 e = *((volatile unsigned long long *)ce);
						if ((e & ~0x0fffffffull) == tag) left = min(left, (u32)(e & 0x0fffffffull));
					}
					// Many walkers: thousands of them may be inside one repeat at once, all scanning one bin from the same end and
					// losing the same claims (62 M lost claims for 2.3 M reads on a genome with a long poly-A run).  All reads of
					// a bin carry the same key, so most probes start a walker- and lane-specific distance (< 512 entries) below the cursor;
					// one in eight starts at the cursor and keeps it moving.  (One walker: the reference's scan, from the tail.)
					bool at_cursor = true;
					if (a.extend != 0 && left > 256u) {
						const u32 hsh = (wid * 0x9E3779B1u) ^ ((u32)lane * 0x85EBCA6Bu) ^ (c_probes * 0xC2B2AE35u);
						if ((hsh >> 29) != 0u) { left -= (hsh & 0x00ffffffu) % min(left >> 1, 512u); at_cursor = false; }
					}
					const u32 top = left;
					bool tail_claimed = at_cursor; // every entry from `top` down to here was claimed (and `top` is the cursor)
					// many walkers: a probe gives up after 8 x maxsearch entries (one walker: the reference's scan, to the end)
					u32 budget = a.extend != 0 ? 8u * (u32)a.maxsearch : 0xffffffffu;
					while (left > 0 && seen < a.maxsearch && budget-- > 0u) {
						left--;
						const u32 rid = __ldg(&ids_of(pc.key)[lo + left]);
						const u32 cw = ldvol(peek_word(a, rid));
						if (!((cw >> (rid & 31)) & 1u)) continue; // removed from the bin in the reference (505-514)
						if (tail_claimed) {
							tail_claimed = false;
							if (cacheable && left + 1 < top) *((volatile unsigned long long *)ce) = tag | (u64)(left + 1);
						}
						load_read<NW>(a.reads, rid, rw);
						seen++;
						c_cmp++;
						if (hamming<NW>(hw, hr, hm, rw) <= a.thresh) { cand = rid; ps = P_CAND; break; }
					}
					if (tail_claimed && cacheable && left < top) *((volatile unsigned long long *)ce) = tag | (u64)left;
					return;
				}
				while (left > 0 && seen < a.maxsearch) {
					left--;
					const u32 rid = size == 1 ? lo : __ldg(&ids_of(pc.key)[lo + left]);
					const u32 cw = ldvol(peek_word(a, rid)); // claim bit and read are fetched together
					load_read<NW>(a.reads, rid, rw);
					if (!((cw >> (rid & 31)) & 1u)) continue; // removed from the bin in the reference (505-514)
					seen++;
					c_cmp++;
					if (hamming<NW>(hw, hr, hm, rw) <= a.thresh) { cand = rid; ps = P_CAND; break; }
				}
			};
			int k_win = 0;
			u32 bh = 0; // extend: lanes behind the first read of the round whose candidates were claimed in the same wave
			// one step of a lane's bucket chain
			auto resolve = [&]() {
				const int r = bucket_step(pc.key, pc.s0, pc.s1, pc.home, lo, size);
				if (r == 1) { ps = P_BIN; left = size; c_hits++; }
				else if (r == 0) ps = P_DEAD;
				else {
					pc.home = false;
					pc.h += 2; // the table ends with two empty slots and never wraps
					const ulonglong2 *sl = slots_of(pc.key);
					pc.s0 = __ldg(&sl[pc.h]);
					pc.s1 = __ldg(&sl[pc.h + 1]);
				}
			};
			// `win` = lane of the first candidate in sequential order once it is known (G before that)
			int win = G;
			while (true) {
				const bool mine_to_work = sub > win || win == G; // with a winner known, only the lanes behind it go on (harvest)
				if (mine_to_work && ps == P_PENDING) resolve();
				if (mine_to_work && ps == P_BIN) advance(hw, hr, hm);
				if (win == G) {
					const u32 bc = __ballot_sync(gmask, ps == P_CAND) >> gbase, bp = __ballot_sync(gmask, ps == P_PENDING) >> gbase;
					if (!(bc | bp)) break; // nothing matches in these shifts
					if (!(bc && (!bp || (bc & (0u - bc)) < (bp & (0u - bp))))) continue;
					win = __ffs(bc) - 1;
					if (a.extend == 0) {
						// the reference's walk (reorder.cpp:545-556): the first candidate in sequential order is claimed
						int got = 0;
						if (sub == win) {
							got = try_claim(a, cand);
							if (!got) { atomicAdd(&s.cnt[3], 1u); ps = P_BIN; }
						}
						got = __shfl_sync(gmask, got, gbase + win);
						if (!got) { win = G; continue; }
						found = true;
						k_win = win;
						break;
					}
				}
				// Harvest (not in the reference; with `extend`, i.e. never with one walker): the lanes behind the winner have
				// already fetched their buckets of this round -- reads that start a few positions further on, which the next
				// rounds would find again one by one.  They first finish their open bucket chains and test their candidates
				// against this window; then the winner and the lanes behind it, up to the first lane that is not settled -- a
				// bin with entries left (reads that start at the same position: the next round finds them at shift 0, with all
				// lanes at work) -- claim their candidates in ONE wave of atomics (one claim per distinct read).  The claimed
				// reads are appended in lane order = in order of their start positions, each with the difference of the
				// shifts.  The probe round, the candidate fetches and the claim round trip (over NVLink when the read's bitmap
				// range lives on another GPU) are paid once for all of them.
				// (Draining the multi-entry bins inside the harvest was tried: it finds more reads per round, 1.55 against
				// 1.0 extra, but its one-lane dependent loads make the walk slower, 31.7 ms against 24.5.)
				if (__any_sync(gmask, sub > win && (ps == P_PENDING || ps == P_BIN))) continue;
				const bool unsettled = sub >= win && ps == P_CAND && left > 0 && seen < a.maxsearch;
				const u32 bu = __ballot_sync(gmask, unsettled) >> gbase;
				// a lane whose bin has entries left ends the harvest behind its own shift: the lanes of the same shift (the
				// other three probe kinds) may still give what they hold, the window then stops at that position
				const int limit = bu ? ((__ffs(bu) - 1) | 3) : G - 1;
				const bool mine = sub >= win && sub <= limit && ps == P_CAND;
				const u32 grp = __match_any_sync(gmask, mine ? (u64)cand : ((1ull << 32) | (u64)lane)); // lanes that hold the same read
				const bool first_of_grp = mine && (__ffs(grp) - 1) == lane;
				int got = 0;
				if (first_of_grp) { got = try_claim(a, cand); if (!got) atomicAdd(&s.cnt[3], 1u); }
				bh = __ballot_sync(gmask, got != 0) >> gbase;
				if (mine && !got) ps = P_BIN; // lost (or the same read as a lane in front): on with the bin if the round goes on
				if (!bh) { win = G; continue; } // every claim of the wave was lost to other walkers
				found = true;
				k_win = __ffs(bh) - 1;
				bh &= bh - 1;
				break;
			}
			if (found) {
				k_rid = __shfl_sync(gmask, cand, gbase + k_win);
				k_j = jb + (k_win >> 2);
				k_rev = (k_win & 3) >= 2;
				if (sub == k_win) {
#pragma unroll
					for (int k = 0; k < W2; k++) s.cur[k] = rw[k];
				}
			}
#ifdef WALK_PROF
			if (dry >= DRY_HEADS) { PROF_ADD(3); PROF_CNT(6, 1); } else
#endif
			PROF_ADD(2);

			// ---- a read is appended (reorder.cpp:560-578 / 624-641); the lane that holds it has put it into s.cur
			auto append = [&](u32 rid, int shift, int rv, bool last) { // last: no other read of this round follows
				current = rid;
				__syncwarp(gmask);
				update_ref<NW, G>(s, L, sub, gmask, false, rv != 0, shift, head, last);
				if (leader) {
					if (!left_mode) {
						if (prev_unmatched) emit(mkrec(prev, (u32)L, 0, 0, 0));
						emit(mkrec(current, (u32)shift, (u32)rv, 1, 0));
					} else {
						// left run, found on the reverse-complement strand: in the final order this read precedes the one found
						// before it, which therefore gets this shift; orientations flip
						if (s.lcount == 0) s.j1 = (u32)shift;
						else lemit(mkrec(s.pend, (u32)shift, s.pend_f ^ 1u, 1, 0));
						s.pend = current;
						s.pend_f = (u32)rv;
						s.lcount++;
					}
				}
				if (!left_mode) prev_unmatched = false;
			};
			if (found) {
				append(k_rid, k_j, k_rev, bh == 0);
				int last_rel = k_win >> 2;
				while (bh) {
					const int w2 = __ffs(bh) - 1;
					bh &= bh - 1;
					const u32 rid2 = __shfl_sync(gmask, cand, gbase + w2);
					__syncwarp(gmask);
					if (sub == w2) {
#pragma unroll
						for (int k = 0; k < W2; k++) s.cur[k] = rw[k];
					}
					append(rid2, (w2 >> 2) - last_rel, (w2 & 3) >= 2, bh == 0);
					last_rel = w2 >> 2;
					if (leader) s.cnt[2]++;
				}
				dry = 0;
				jb = 0;
					PROF_CNT(10, 1);
			} else {
				jb += SPR;
				// Tail of the job: a walker whose last heads all stayed alone is working off the reads that no dictionary
				// window finds (errors in both windows); their neighbours are claimed, and such a head is searched over the
				// first round of shifts only.  It is declared a singleton like any other and goes to stage II's pool.  (Not
				// with one walker / extend off, where the walk is the reference's, shift for shift.)
				const int jmax = (a.extend != 0 && prev_unmatched && dry >= STREAK_HEADS) ? min(a.maxmatch, STREAK_SHIFTS) : a.maxmatch;
				if (jb >= jmax) state = S_CHAINEND;
			}
			PROF_ADD(4);
		}

		// ---- nothing matches the window any more
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
				if (sub < W2) s.cur[sub] = __ldg(reinterpret_cast<const u32 *>(a.reads + (size_t)current * NW) + sub);
				__syncwarp(gmask);
				update_ref<NW, G>(s, L, sub, gmask, true, false, 0, head);
				jb = 0;
				state = S_SEARCH;
			} else {
				if (prev_unmatched) {
					if (leader) emit(mkrec(prev, 0, 0, 0, 1)); // the head stayed alone: singleton (672-684)
					dry++;
				}
				state = S_RESTART;
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
		c_probes += __shfl_xor_sync(FULL, c_probes, o);
		c_hits += __shfl_xor_sync(FULL, c_hits, o);
		c_cmp += __shfl_xor_sync(FULL, c_cmp, o);
	}
	__syncwarp();
	if (lane == 0) { atomicAdd(&a.counters[1], (u64)c_probes); atomicAdd(&a.counters[2], (u64)c_hits); atomicAdd(&a.counters[3], (u64)c_cmp); }
	if (leader) {
		atomicAdd(&a.counters[0], (u64)s.cnt[0]); atomicAdd(&a.counters[5], (u64)s.cnt[1]); atomicAdd(&a.counters[6], (u64)s.cnt[2]);
		atomicAdd(&a.counters[4], (u64)s.cnt[3]);
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
size_t walk_smem(const WalkArgs &a)
{
	return sizeof(WalkerSmem<NW>) * WALK_WARPS * (32 / G) + (size_t)2 * a.maxmatch * 2 * NW * sizeof(u32);
}
template <int NW, int G>
int launch_walk(harcgpu_ctx *c, const WalkArgs &a)
{
	constexpr int WPB = WALK_WARPS * 32 / G; // walkers per block
	const size_t smem = walk_smem<NW, G>(a);
	CK(cudaFuncSetAttribute(walk_kernel<NW, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	walk_kernel<NW, G><<<KL + cdiv(a.walkers, WPB), WALK_WARPS * 32, smem, c->st>>>(a);
	CK(cudaGetLastError());
	return 0;
}

// walkers that can be resident at once: a walker that is not resident only starts after the others have finished
template <int NW, int G>
int resident_walkers(harcgpu_ctx *c, const WalkArgs &a, u32 *out)
{
	constexpr int WPB = WALK_WARPS * 32 / G;
	const size_t smem = walk_smem<NW, G>(a);
	int nb = 0, sms = 0;
	CK(cudaFuncSetAttribute(walk_kernel<NW, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, walk_kernel<NW, G>, WALK_WARPS * 32, smem));
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
	*out = (u32)nb * (u32)sms * WPB;
	return 0;
}
} // namespace

#define DISPATCH_NW_G(NWv, Gv, CALL)                                                         \
	switch ((NWv) * 100 + (Gv)) {                                                            \
	case 116: { constexpr int NW = 1, G = 16; CALL; } break;                                 \
	case 216: { constexpr int NW = 2, G = 16; CALL; } break;                                 \
	case 316: { constexpr int NW = 3, G = 16; CALL; } break;                                 \
	case 416: { constexpr int NW = 4, G = 16; CALL; } break;                                 \
	case 516: { constexpr int NW = 5, G = 16; CALL; } break;                                 \
	case 616: { constexpr int NW = 6, G = 16; CALL; } break;                                 \
	case 716: { constexpr int NW = 7, G = 16; CALL; } break;                                 \
	case 816: { constexpr int NW = 8, G = 16; CALL; } break;                                 \
	case 132: { constexpr int NW = 1, G = 32; CALL; } break;                                 \
	case 232: { constexpr int NW = 2, G = 32; CALL; } break;                                 \
	case 332: { constexpr int NW = 3, G = 32; CALL; } break;                                 \
	case 432: { constexpr int NW = 4, G = 32; CALL; } break;                                 \
	case 532: { constexpr int NW = 5, G = 32; CALL; } break;                                 \
	case 632: { constexpr int NW = 6, G = 32; CALL; } break;                                 \
	case 732: { constexpr int NW = 7, G = 32; CALL; } break;                                 \
	case 832: { constexpr int NW = 8, G = 32; CALL; } break;                                 \
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
	CK(cudaMemsetAsync(c->counters, 0, 8 * sizeof(u64), st));
	auto no_reads = [&]() { // empty streams (the getters and stage II take null arrays of zero entries, but keep them valid)
		if (c->alloc(&c->order, 1) || c->alloc(&c->order_s, 1) || c->alloc(&c->rev, 1) || c->alloc(&c->flag, 1) || c->alloc(&c->pos, 1)) return -1;
		c->reordered = true; c->ms["walk"] = 0; c->ms["finalize"] = 0;
		return 0;
	};
	if (n == 0) return no_reads();

	// walkers: the reference's num_thr.  Auto: one walker per reads_per_walker reads (every extra walker costs chain
	// heads), capped at what is resident at once.
	u32 resident = 0;
	int rc = -1;
	int lanes = c->p.lanes_per_walker;
	if (const char *e = getenv("HARCGPU_LANES")) lanes = atoi(e); // tuning aid
	if (lanes == 0) lanes = 32;
	if (lanes != 16 && lanes != 32) { harcgpu_set_error("lanes_per_walker must be 16 or 32"); return -1; }
	if (c->p.maxmatch < 1 || c->p.maxmatch > 16 * c->NW) { harcgpu_set_error("maxmatch %d out of range for read length %d", c->p.maxmatch, c->L); return -1; }
	WalkArgs a;
	a.maxmatch = c->p.maxmatch;
	DISPATCH_NW_G(c->NW, lanes, (rc = resident_walkers<NW, G>(c, a, &resident)));
	if (rc) return rc;
	// one job on several GPUs: this GPU's walkers own the id range [base, base + n_loc) for starts and restarts
	const bool sharded = c->shard_world > 1;
	if (sharded && (c->shard_n != n || !c->shard_ready)) {
		harcgpu_set_error("one job on several GPUs: call harcgpu_job_reorder (after harcgpu_job_init/connect/load_reads for these %u reads)", n);
		return -1;
	}
	c->shard_ready = false; // a sharded pass consumes the reset
	const u32 base = sharded ? std::min<u64>((u64)c->shard_rank * c->seg_per, n) : 0u;
	const u32 n_loc = sharded ? (u32)(std::min<u64>(((u64)c->shard_rank + 1) * c->seg_per, n) - base) : n;
	const u32 per = c->p.reads_per_walker > 0 ? (u32)c->p.reads_per_walker : 4096u;
	u32 walkers = c->p.walkers > 0 ? (u32)c->p.walkers : (u32)std::min<u64>(resident, std::max<u64>(1, n_loc / per));
	if (walkers > n_loc) walkers = n_loc;
	if (walkers == 0) { // no read in this GPU's range: it still meets the others where they start their walks
		if (sharded && job_barrier(c)) return -1;
		return no_reads();
	}
	c->walkers_used = walkers;
	// left extension: off for a single walker unless asked for (one walker without it = the reference at num_thr=1)
	const int extend = c->p.extend > 0 ? 1 : (c->p.extend < 0 ? 0 : (walkers > 1 ? 1 : 0));

	// Room of the record logs: every read once.  One job on several GPUs: a GPU's walkers may claim any read, but the GPUs
	// advance side by side, so twice its share (+ slack) is reserved instead of the whole job; running over is detected.
	const u64 cap_reads = sharded ? std::min<u64>(n, 2 * (((u64)n + c->shard_world - 1) / c->shard_world) + (1u << 20)) : n;
	u32 max_chunks = (u32)(cap_reads / CHUNK) + walkers + 1;
	u64 *recs = nullptr, *chunk_key = nullptr, *key_sorted = nullptr, *scan_tmp = nullptr, *lrecs = nullptr;
	u32 *chunk_fill = nullptr, *ctrs = nullptr, *chunk_id = nullptr, *chunk_sorted = nullptr, *cm = nullptr, *cs = nullptr,
	    *om = nullptr, *os = nullptr, *totals = nullptr, *lprev = nullptr;
	if (c->alloc(&recs, (size_t)max_chunks * CHUNK) || c->alloc(&chunk_key, max_chunks) || c->alloc(&chunk_fill, max_chunks) ||
	    c->alloc(&ctrs, 2))
		return -1;
	if (extend && (c->alloc(&lrecs, (size_t)max_chunks * CHUNK) || c->alloc(&lprev, max_chunks))) return -1;
	CK(cudaMemsetAsync(ctrs, 0, 8, st));
	CK(cudaMemsetAsync(chunk_fill, 0, 4 * (size_t)max_chunks, st));
	// one GPU: the bitmap.  Several GPUs: the authoritative ranges are armed by harcgpu_job_reorder before the barrier that
	// precedes the walk; c->claim serves as this GPU's hint bitmap for the peers' ranges.
	init_claim_kernel<<<KL + cdiv((n + 31) / 32, 256), 256, 0, st>>>(c->claim, n);
	CK(cudaGetLastError());
	u32 *stripe_done = nullptr;
	if (!c->tailc && c->alloc(&c->tailc, TAILC_ENTRIES)) return -1; // kept for the life of the context
	unsigned long long *tailc = c->tailc;
	if (c->alloc(&stripe_done, walkers)) return -1;
	CK(cudaMemsetAsync(stripe_done, 0, 4 * (size_t)walkers, st));
	CK(cudaMemsetAsync(tailc, 0xff, 8 * (size_t)TAILC_ENTRIES, st)); // no bin has this tag

	a.reads = c->reads; a.n = n; a.L = c->L; a.maxmatch = c->p.maxmatch; a.thresh = c->p.thresh; a.maxsearch = c->p.maxsearch;
	a.numdict = c->p.numdict; a.extend = extend;
	for (int l = 0; l < 2; l++) {
		int ll = l < c->p.numdict ? l : 0;
		a.d[l].slots = c->d1[ll].slots; a.d[l].ids = c->d1[ll].ids; a.d[l].slot_shift = c->d1[ll].slot_shift;
		a.d[l].dstart = c->p.dict_start[ll]; a.d[l].dend = c->p.dict_end[ll];
		a.d[l].world = c->dicts_sharded ? c->shard_world : 0;
		for (int r = 0; r < 8; r++) {
			const bool on = c->dicts_sharded && r < c->shard_world && c->arena[r];
			a.d[l].sslots[r] = on ? (const ulonglong2 *)(c->arena[r] + c->arena_slots_off[ll]) : nullptr;
			a.d[l].sids[r] = on ? (const u32 *)(c->arena[r] + c->arena_ids_off[ll]) : nullptr;
		}
		a.kbits[l] = c->d1[ll].nbits;
	}
	a.bloom = c->dicts_sharded && c->job_bloom ? (const u32 *)(c->arena[c->shard_rank] + c->arena_bloom_off) : nullptr;
	a.bloom_seg_words = c->bloom_seg_words;
	if (!c->dicts_sharded && !sharded && c->bloom1) { a.bloom = c->bloom1; a.bloom_seg_words = c->bloom1_words; }
	a.claim = sharded ? c->seg[c->shard_rank] : c->claim; a.hint = c->claim; a.stripe_done = stripe_done; a.walkers = walkers;
	for (int r = 0; r < 8; r++) a.seg[r] = sharded && r < c->shard_world ? c->seg[r] : nullptr;
	a.seg_per = sharded ? c->seg_per : 0u; a.base = base; a.n_loc = n_loc; a.world = sharded ? c->shard_world : 1;
	a.recs = recs; a.chunk_key = chunk_key; a.chunk_fill = chunk_fill; a.chunk_ctr = ctrs; a.max_chunks = max_chunks;
	a.lrecs = lrecs; a.lprev = lprev; a.lchunk_ctr = ctrs + 1; a.max_lchunks = max_chunks;
	a.counters = c->counters;
	a.tailc = tailc; a.tailc_mask = TAILC_ENTRIES - 1;
	// tuning aid: keep the claim bitmap (1 bit per read, hit by every candidate test and claim) in the persisting part of L2
	bool l2win = false;
	if (const char *e = getenv("HARCGPU_L2PERSIST")) {
		const bool pin_bloom = c->bloom1 && a.bloom == c->bloom1; // the filter rather than the claim bitmap if there is one
		const size_t want = pin_bloom ? (size_t)c->bloom1_words * 4 : ((size_t)n + 31) / 32 * 4;
		int maxwin = 0;
		CK(cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, c->device));
		if (atoi(e) > 0 && !sharded && want <= (size_t)maxwin) {
			CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(want * 2, (size_t)atoi(e) << 20)));
			cudaStreamAttrValue av;
			memset(&av, 0, sizeof av);
			av.accessPolicyWindow.base_ptr = pin_bloom ? (void *)c->bloom1 : (void *)c->claim;
			av.accessPolicyWindow.num_bytes = want;
			av.accessPolicyWindow.hitRatio = 1.0f;
			av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
			av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
			CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
			l2win = true;
		}
	}
	// One job on several GPUs: the GPUs meet HERE, with every allocation and memset of this call behind them, so that all
	// walk kernels start within microseconds of one another.  (A GPU that starts late -- one cudaMalloc is ~20 ms when the
	// block has to be mapped for the peers -- finds most reads claimed, ends up with a small share, and the sizes of
	// everything downstream change from pass to pass.)  Every range of the bitmap was armed before this call.
	if (sharded && job_barrier(c)) return -1;
	c->tic();
	DISPATCH_NW_G(c->NW, lanes, (rc = launch_walk<NW, G>(c, a)));
	if (rc) return rc;
	c->toc("walk");
	if (l2win) {
		cudaStreamAttrValue av;
		memset(&av, 0, sizeof av);
		CK(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av));
		CK(cudaCtxResetPersistingL2Cache());
	}
	CK(cudaGetLastError());
#ifdef WALK_PROF
	{
		unsigned long long h[16], z[16] = { 0 };
		CK(cudaStreamSynchronize(st));
		CK(cudaMemcpyFromSymbol(h, g_walk_prof, sizeof h));
		CK(cudaMemcpyToSymbol(g_walk_prof, z, sizeof z));
		fprintf(stderr, "WALK_PROF walkers %u cycles: restart %llu newhead %llu search %llu (+ in singleton streaks %llu) append %llu chainend %llu | rounds %llu (%llu in streaks) steps %llu found %llu | warp life avg %llu max %llu warps %llu\n",
		        walkers, h[0], h[1], h[2], h[3], h[4], h[5], h[8], h[6], h[9], h[10], h[13] ? h[11] / h[13] : 0ull, h[12], h[13]);
	}
#endif

	// ---- finalize
	c->tic();
	u32 hctr[2] = { 0, 0 };
	CK(cudaMemcpyAsync(hctr, ctrs, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	const u32 nchunks = hctr[0], nlchunks = hctr[1];
	if (nchunks > max_chunks || nlchunks > max_chunks) {
		harcgpu_set_error("record log overflow: this GPU's walkers claimed more than %llu reads (its share of the job twice over)", (unsigned long long)cap_reads);
		return -1;
	}
	if (c->alloc(&key_sorted, nchunks) || c->alloc(&chunk_id, nchunks) || c->alloc(&chunk_sorted, nchunks) || c->alloc(&cm, nchunks) ||
	    c->alloc(&cs, nchunks) || c->alloc(&om, nchunks) || c->alloc(&os, nchunks) || c->alloc(&totals, 2) ||
	    c->alloc(&scan_tmp, scan_tmp_elems(nchunks)))
		return -1;
	iota_kernel<<<KL + cdiv(nchunks, 256), 256, 0, st>>>(chunk_id, nchunks);
	CK(cudaGetLastError());
	{
		// key = walker << 32 | sequence number: two stable sorts over the bits that can be set (sequence < nchunks, walker < walkers)
		int sb = 1, wb = 1;
		while (sb < 32 && (nchunks >> sb) != 0) sb++;
		while (wb < 32 && (walkers >> wb) != 0) wb++;
		if (radix_sort_pairs(c, &chunk_key, &key_sorted, &chunk_id, &chunk_sorted, nchunks, 0, sb)) return -1;
		if (radix_sort_pairs(c, &chunk_key, &key_sorted, &chunk_id, &chunk_sorted, nchunks, 32, 32 + wb)) return -1;
		std::swap(chunk_id, chunk_sorted); // chunk_sorted = chunk ids in (walker, sequence) order
	}
	chunk_count_kernel<<<KL + cdiv((size_t)nchunks * 32, 256), 256, 0, st>>>(recs, chunk_sorted, chunk_fill, nchunks, cm, cs);
	CK(cudaGetLastError());
	if (exclusive_scan_u32(cm, om, nchunks, scan_tmp, totals, st)) return -1;
	if (exclusive_scan_u32(cs, os, nchunks, scan_tmp, totals + 1, st)) return -1;
	u32 tot[2];
	CK(cudaMemcpyAsync(tot, totals, 8, cudaMemcpyDeviceToHost, st));
	u64 cnt[8];
	CK(cudaMemcpyAsync(cnt, c->counters, 64, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	// the five streams, at their exact sizes
	if (c->alloc(&c->order, tot[0]) || c->alloc(&c->order_s, tot[1]) || c->alloc(&c->rev, tot[0]) || c->alloc(&c->flag, tot[0]) || c->alloc(&c->pos, tot[0]))
		return -1;
	chunk_gather_kernel<<<KL + cdiv((size_t)nchunks * 32, 256), 256, 0, st>>>(recs, chunk_sorted, chunk_fill, nchunks, om, os, c->order,
	                                                                      c->rev, c->flag, c->pos, c->order_s);
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(st));
	c->toc("finalize");
	c->n_matched = tot[0]; c->n_single = tot[1]; c->n_unmatched = (u32)cnt[5];
	void *tmp[] = { recs, chunk_key, key_sorted, chunk_fill, ctrs, chunk_id, chunk_sorted, cm, cs, om, os, totals, scan_tmp,
	                stripe_done, lrecs, lprev };
	for (void *q : tmp) c->release(q);
	if (!sharded && (u64)tot[0] + tot[1] != n) { harcgpu_set_error("reorder lost reads: %u matched + %u singletons != %u", tot[0], tot[1], n); return -1; }
	c->reordered = true;
	return 0;
}
