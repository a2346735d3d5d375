// Stage I of HARC -- hash-based read reordering (reference: src/reorder.cpp) -- as sm_100a kernels.
//
//   K1 pack_kernel       ASCII lines -> 2 bits/base (one TMA bulk copy per block)   reorder.cpp:203-209, 240-263
//   K2 keys_kernel       (read & mask) >> 2*dict_start, mixed by one multiply        reorder.cpp:284-302
//   K3 radix sort (sort.cu) + heads/scan/bins + prefix-maximum table placement       reorder.cpp:305-390 (sort, dedup, MPHF, CSR fill)
//   K4 walk_kernel       warp-per-chain greedy walk (walk.cu)                        reorder.cpp:434-703, 863-915
//   K5 finalize (chunk sort, partition; walk.cu) + unpack_kernel                     reorder.cpp:722-830
//
// Data layout in HBM: reads[n][NW] u64 (NW = ceil(2L/64)); per dictionary the CSR (mixed keys, start, ids) in mixed-key
// order and a 16-byte-slot open-addressing table ordered the same way, load factor <= 0.5; a claim bitmap (1 bit/read);
// a chunked record log.
#include "ctx.h"
#include <utility>
#include <algorithm>
#include <limits.h>

// ------------------------------------------------------------------------------------------------ K1 pack
namespace {
constexpr int PACK_RPB = 64; // reads per block: 64*(L+1) bytes is a multiple of 16 for every L

// 4 ASCII bases in a u32 -> their four 2-bit codes in 8 bits (base 0 in bits 0-1).  (ch >> 1) & 3 is A0 C1 T2 G3; the
// reference's code (reorder.cpp:188-195) is A0 G1 C2 T3 = ((b0 ^ b1) << 1) | b1 of those two bits.
__device__ __forceinline__ u32 squeeze4(u32 v) // one 2-bit field per byte -> 8 contiguous bits
{
	return (v | (v >> 6) | (v >> 12) | (v >> 18)) & 0xffu;
}
__device__ __forceinline__ u32 codes4(u32 w)
{
	const u32 b0 = (w >> 1) & 0x01010101u, b1 = (w >> 2) & 0x01010101u;
	return ((b0 ^ b1) << 1) | b1;
}
// One block stages 64 lines in shared memory with 16-byte loads, then one thread per output word packs its 32 bases
// four at a time from aligned 32-bit shared-memory words (funnel shift for the line's byte offset).
// WITH_N (stage II pool, encoder.cpp:731-745): 'N' is stored as code 0 and flagged in a second word array at the
// low bit of the base's pair; the reference's 3-bit code is then 2*code2 + nflag.
template <bool WITH_N>
__global__ void __launch_bounds__(256) pack_kernel(const char *__restrict__ ascii, u32 n, int L, int NWo, u64 *__restrict__ out,
                                                   u64 *__restrict__ outN)
{
	extern __shared__ uint4 stage4[];
	char *stage = reinterpret_cast<char *>(stage4);
	const u32 *stage32 = reinterpret_cast<const u32 *>(stage4);
	const size_t line = (size_t)L + 1;
	const size_t r0 = (size_t)blockIdx.x * PACK_RPB;
	const u32 nr = (u32)min((size_t)PACK_RPB, (size_t)n - r0);
	const size_t bytes = (size_t)nr * line;
	const char *src = ascii + r0 * line;
	if (bytes % 16 == 0) {
		// a full block: its 64 lines are one 16-byte-aligned span, fetched by ONE bulk copy of the TMA unit
		// (cp.async.bulk, global -> shared, completion counted in bytes on an mbarrier) issued by one thread
		__shared__ __align__(8) unsigned long long bar;
		const u32 bar_a = (u32)__cvta_generic_to_shared(&bar), dst_a = (u32)__cvta_generic_to_shared(stage);
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((u32)bytes) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a), "l"(src),
			             "r"((u32)bytes), "r"(bar_a)
			             : "memory");
		}
		__syncthreads(); // the barrier is initialised before anybody polls it
		u32 done = 0;
		while (!done)
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a) : "memory");
	} else {
		const size_t nvec = bytes / 16;
		const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
		for (size_t v = threadIdx.x; v < nvec; v += blockDim.x) stage4[v] = __ldg(&src4[v]);
		for (size_t b = nvec * 16 + threadIdx.x; b < bytes; b += blockDim.x) stage[b] = src[b];
		__syncthreads();
	}
	for (u32 w = threadIdx.x; w < nr * (u32)NWo; w += blockDim.x) {
		const u32 r = w / NWo, k = w % NWo;
		const u32 o = r * (u32)line + 32 * k; // byte offset of the word's first base
		const u32 q = o >> 2, sh = (o & 3u) * 8;
		const int nb = min(32, L - 32 * (int)k); // bases in this word
		u64 v = 0, vn = 0;
		u32 lo = stage32[q];
#pragma unroll
		for (int i = 0; i < 8; i++) {
			const u32 hi = stage32[q + i + 1]; // the staging area has 32 bytes of slack
			const u32 ch = __funnelshift_r(lo, hi, sh);
			lo = hi;
			u32 c = codes4(ch);
			if (WITH_N) {
				const u32 isn = __vcmpeq4(ch, 0x4E4E4E4Eu); // 0xff where the base is 'N'
				c &= ~isn;
				vn |= (u64)squeeze4(isn & 0x01010101u) << (8 * i);
			}
			v |= (u64)squeeze4(c) << (8 * i);
		}
		const u64 m = lowmask(2 * nb);
		out[(r0 + r) * NWo + k] = v & m;
		if (WITH_N) outN[(r0 + r) * NWo + k] = vn & m;
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
__global__ void __launch_bounds__(256) keys_kernel(const u64 *__restrict__ reads, u32 n, int words, int bitpos, int nbits, u32 id0,
                                                   u64 *__restrict__ keys, u32 *__restrict__ ids)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const u64 *r = reads + (size_t)i * words;
	int q = bitpos >> 6, sh = bitpos & 63;
	u64 v = __ldg(&r[q]) >> sh;
	if (sh && q + 1 < words) v |= __ldg(&r[q + 1]) << (64 - sh);
	if (nbits < 64) v &= (1ull << nbits) - 1;
	keys[i] = key_mix(v); // the dictionary is built, stored and probed in mixed-key order (common.cuh)
	ids[i] = id0 + i;
}

// stage II keys (encoder.cpp:893-911): the reference's 3-bit code of base b is 2*code2 + nflag.  The window's 2-bit
// codes and N flags (bit 2b of the flag words) are cut out with two shifts each and moved apart by spread2to3.
__global__ void __launch_bounds__(256) keys3_kernel(const u64 *__restrict__ r2, const u64 *__restrict__ rN, u32 n, int words, int ds,
                                                    int de, u64 *__restrict__ keys, u32 *__restrict__ ids)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int nb = de - ds + 1, bitpos = 2 * ds, q = bitpos >> 6, sh = bitpos & 63;
	const u64 *a = r2 + (size_t)i * words, *b = rN + (size_t)i * words;
	u64 c2 = __ldg(&a[q]) >> sh, nf = __ldg(&b[q]) >> sh;
	if (sh && q + 1 < words) { c2 |= __ldg(&a[q + 1]) << (64 - sh); nf |= __ldg(&b[q + 1]) << (64 - sh); }
	// spread2to3 puts 2*c at bits 3t: for the flags (c = 0 or 1 in the low bit of the pair) that is flag << 1
	const u64 key = spread2to3(c2, nb) | (spread2to3(nf & 0x5555555555555555ull, nb) >> 1);
	keys[i] = key_mix(key);
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

// ---- placement of the bins in the key table.  The bins come in mixed-key order, so their home buckets never decrease
// and sequential linear probing puts bin b at pos_b = max(home_b, pos_{b-1} + 1) = b + max_{j <= b}(home_j - j): an
// inclusive prefix maximum instead of one atomic compare-and-swap per bin, and the slots are written in order.
constexpr int PM_TILE = 1024; // bins per block of the prefix maximum: 256 threads x 4
__global__ void __launch_bounds__(256) pm_local_kernel(const u64 *__restrict__ mixed, u32 nk, int shift, int world, long long *__restrict__ pm,
                                                       long long *__restrict__ block_max)
{
	__shared__ long long wmax[8];
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const size_t b0 = (size_t)blockIdx.x * PM_TILE + 4 * threadIdx.x;
	long long v[4], run = LLONG_MIN;
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const size_t b = b0 + k;
		v[k] = b < nk ? (long long)slot_home(mixed[b], shift, world) - (long long)b : LLONG_MIN;
		run = max(run, v[k]);
		v[k] = run; // inclusive inside the thread
	}
	long long incl = run;
	for (int o = 1; o < 32; o <<= 1) {
		const long long t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl = max(incl, t);
	}
	if (lane == 31) wmax[w] = incl;
	__syncthreads();
	long long before = __shfl_up_sync(0xffffffffu, incl, 1);
	if (lane == 0) before = LLONG_MIN;
	for (int k = 0; k < w; k++) before = max(before, wmax[k]);
#pragma unroll
	for (int k = 0; k < 4; k++)
		if (b0 + k < nk) pm[b0 + k] = max(v[k], before);
	if (threadIdx.x == 255) block_max[blockIdx.x] = max(incl, before);
}
// exclusive prefix maximum of the block maxima, by one block
__global__ void __launch_bounds__(1024) pm_blocks_kernel(long long *block_max, u32 nblocks)
{
	__shared__ long long wmax[32];
	__shared__ long long carry_s;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry_s = LLONG_MIN;
	__syncthreads();
	for (u32 base = 0; base < nblocks; base += 1024) {
		const u32 i = base + threadIdx.x;
		const long long v = i < nblocks ? block_max[i] : LLONG_MIN;
		long long incl = v;
		for (int o = 1; o < 32; o <<= 1) {
			const long long t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= o) incl = max(incl, t);
		}
		if (lane == 31) wmax[w] = incl;
		__syncthreads();
		long long before = __shfl_up_sync(0xffffffffu, incl, 1);
		if (lane == 0) before = LLONG_MIN;
		for (int k = 0; k < w; k++) before = max(before, wmax[k]);
		const long long carry = carry_s;
		before = max(before, carry);
		if (i < nblocks) block_max[i] = before; // exclusive
		__syncthreads();
		if (threadIdx.x == 1023) carry_s = max(before, v);
		__syncthreads();
	}
}
// slot of bin b = b + prefix maximum; last[0] = slot of the last bin (the table must reach two slots beyond it)
__global__ void __launch_bounds__(256) pm_finish_kernel(long long *__restrict__ pm, const long long *__restrict__ block_excl, u32 nk, u32 *__restrict__ pos,
                                                        unsigned long long *__restrict__ last)
{
	const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nk) return;
	const long long m = max(pm[b], block_excl[b / PM_TILE]);
	const long long p = (long long)b + m;
	pos[b] = (u32)p;
	if (b == nk - 1) last[0] = (unsigned long long)p;
}
__global__ void __launch_bounds__(256) place_kernel(const u64 *__restrict__ mixed, const u32 *__restrict__ start, const u32 *__restrict__ ids,
                                                    const u32 *__restrict__ pos, u32 nk, ulonglong2 *__restrict__ slots)
{
	const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nk) return;
	const u32 s = start[b], sz = start[b + 1] - s;
	slots[pos[b]] = make_ulonglong2(mixed[b], (u64)(sz == 1 ? ids[s] : s) | ((u64)sz << 32));
}
// a bin outside its home bucket raises the home bucket's overflow bit (common.cuh: bucket_step); runs after place_kernel
__global__ void __launch_bounds__(256) overflow_kernel(const u64 *__restrict__ mixed, const u32 *__restrict__ pos, u32 nk, int shift, int world,
                                                       ulonglong2 *slots)
{
	const u32 b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nk) return;
	const u32 h = slot_home(mixed[b], shift, world);
	if (pos[b] > h + 1) atomicOr(&slots[h].y, SLOT_OVERFLOW);
}
} // namespace

int s1_pack_reads(harcgpu_ctx *c, const void *d_ascii, u32 n)
{
	if (n == 0) return 0;
	size_t smem = (size_t)PACK_RPB * (c->L + 1);
	smem = (smem + 15) / 16 * 16 + 48; // slack: the last word of the last line reads up to 35 bytes past its start
	return s1_pack_reads_to(c, d_ascii, n, c->reads);
}
int s1_pack_reads_to(harcgpu_ctx *c, const void *d_ascii, u32 n, u64 *out)
{
	if (n == 0) return 0;
	size_t smem = (size_t)PACK_RPB * (c->L + 1);
	smem = (smem + 15) / 16 * 16 + 48;
	pack_kernel<false><<<KL + cdiv(n, PACK_RPB), 256, smem, c->st>>>((const char *)d_ascii, n, c->L, c->NW, out, nullptr);
	CK(cudaGetLastError());
	return 0;
}
int s1_keys(harcgpu_ctx *c, const u64 *reads, u32 n, int words, int bitpos, int nbits, u32 id0, u64 *keys, u32 *ids)
{
	if (n == 0) return 0;
	keys_kernel<<<KL + cdiv(n, 256), 256, 0, c->st>>>(reads, n, words, bitpos, nbits, id0, keys, ids);
	CK(cudaGetLastError());
	return 0;
}
// d_ascii must be 16-byte aligned.  out2/outN: [n][NW]
int s1_packN(harcgpu_ctx *c, const void *d_ascii, u32 n, u64 *out2, u64 *outN)
{
	if (n == 0) return 0;
	size_t smem = (size_t)PACK_RPB * (c->L + 1);
	smem = (smem + 15) / 16 * 16 + 48; // slack: the last word of the last line reads up to 35 bytes past its start
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
	c->release(d.keys); c->release(d.start);
	if (!d.external) { c->release(d.ids); c->release(d.slots); }
	d = DictDev();
}

namespace {
// one job on several GPUs, sharded dictionaries: keep the (key, id) pairs whose key hashes to this GPU's shard
__global__ void __launch_bounds__(256) shard_flag_kernel(const u64 *__restrict__ keys, u32 n, int rank, int world, u32 *__restrict__ flag)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = mix_shard(keys[i], world) == (u32)rank ? 1u : 0u;
}
__global__ void __launch_bounds__(256) shard_compact_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ ids, const u32 *__restrict__ flag,
                                                            const u32 *__restrict__ ex, u32 n, u64 *__restrict__ k_out, u32 *__restrict__ id_out)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && flag[i]) { k_out[ex[i]] = keys[i]; id_out[ex[i]] = ids[i]; }
}
} // namespace

// bits == 2: stage I key = bits [2*ds, 2*(de+1)) of the packed read; bits == 3: stage II key from (reads, readsN)
// shard != nullptr (stage I of one job on several GPUs): only the keys of this GPU's shard are kept; the table and
// the id lists are built in place in the arena the peers have mapped.
int build_dict(harcgpu_ctx *c, DictDev &d, const u64 *reads, const u64 *readsN, u32 n, int words, int ds, int de, int bits,
               const DictShard *shard)
{
	cudaStream_t st = c->st;
	free_dict(c, d);
	if (shard) {
		d.external = true;
		d.ids = shard->ids;
		d.slots = shard->slots;
	}
	const int bitpos = bits * ds, nbits = bits * (de - ds + 1);
	d.bitpos = bitpos; d.nbits = nbits;
	c->lap(nullptr);
	if (shard && shard->pair_keys) n = shard->npairs;
	u64 *k_in = nullptr, *k_out = nullptr, *scan_tmp = nullptr;
	u32 *id_in = nullptr, *head = nullptr, *binidx = nullptr, *d_total = nullptr;
	u32 *id_alt = nullptr;
	if (n == 0) {
		if (!shard && c->alloc(&d.ids, 1)) return -1;
		d.numkeys = 0;
		if (c->alloc(&d.keys, 1) || c->alloc(&d.start, 1)) return -1;
		CK(cudaMemsetAsync(d.start, 0, 4, st));
		d.slot_shift = 60;
		if (shard) {
			int kb = 0;
			while ((1ull << kb) < shard->cap) kb++;
			d.slot_shift = 64 - kb;
			d.nslots = shard->nslots;
			CK(cudaMemsetAsync(d.slots, 0, (size_t)shard->nslots * sizeof(ulonglong2), st));
			CK(cudaStreamSynchronize(st));
			return 0;
		}
		d.nslots = 18;
		if (c->alloc(&d.slots, 18)) return -1;
		CK(cudaMemsetAsync(d.slots, 0, 18 * sizeof(ulonglong2), st));
		return 0;
	}
	if (c->alloc(&k_in, n) || c->alloc(&k_out, n) || c->alloc(&id_in, n) || c->alloc(&head, n) || c->alloc(&binidx, n) ||
	    c->alloc(&scan_tmp, scan_tmp_elems(n)) || c->alloc(&d_total, 1))
		return -1;
	if (shard && shard->pair_keys) {
		// the pairs of this shard came through the key exchange (job.cu): sources in rank order = ids ascending among equal keys
		CK(cudaMemcpyAsync(k_in, shard->pair_keys, 8 * (size_t)n, cudaMemcpyDeviceToDevice, st));
		CK(cudaMemcpyAsync(id_in, shard->pair_ids, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st));
	} else if (bits == 2) keys_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(reads, n, words, bitpos, nbits, 0u, k_in, id_in);
	else keys3_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(reads, readsN, n, words, ds, de, k_in, id_in);
	CK(cudaGetLastError());
	if (shard && !shard->pair_keys) {
		// keep this GPU's keys, in id order (the scan keeps the order, so ids stay ascending inside a bin)
		u32 *flag = head, *ex = binidx, kept = 0; // both arrays are free until the sort is done
		u64 *k_f = k_out;
		u32 *id_f = nullptr;
		if (c->alloc(&id_f, n)) return -1;
		shard_flag_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(k_in, n, shard->rank, shard->world, flag); // k_in holds mixed keys
		CK(cudaGetLastError());
		if (exclusive_scan_u32(flag, ex, n, scan_tmp, d_total, st)) return -1;
		shard_compact_kernel<<<KL + cdiv(n, 256), 256, 0, st>>>(k_in, id_in, flag, ex, n, k_f, id_f);
		CK(cudaGetLastError());
		CK(cudaMemcpyAsync(&kept, d_total, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		// from here on the dictionary is built over the kept pairs only
		std::swap(k_in, k_out); // k_in = compacted keys, k_out = sort output
		c->release(id_in);
		id_in = id_f;
		n = kept;
		if (n == 0) {
			d.numkeys = 0;
			if (c->alloc(&d.keys, 1) || c->alloc(&d.start, 1)) return -1;
			CK(cudaMemsetAsync(d.start, 0, 4, st));
			int kb = 0;
			while ((1ull << kb) < shard->cap) kb++;
			d.slot_shift = 64 - kb;
			d.nslots = shard->nslots;
			CK(cudaMemsetAsync(d.slots, 0, (size_t)shard->nslots * sizeof(ulonglong2), st));
			CK(cudaStreamSynchronize(st));
			c->release(k_in); c->release(k_out); c->release(id_in); c->release(head); c->release(binidx);
			c->release(scan_tmp); c->release(d_total);
			return 0;
		}
	}
	// LSD radix sort is stable and ids start ascending, so ids stay ascending inside a bin (reorder.cpp:371-384)
	if (c->alloc(&id_alt, n)) return -1;
	c->lap(bits == 2 ? "bd1_keys" : "bd2_keys", true);
	if (radix_sort_mixed(c, &k_in, &k_out, &id_in, &id_alt, n)) return -1; // mixed keys use all 64 bits
	c->lap(bits == 2 ? "bd1_sort" : "bd2_sort", true);
	std::swap(k_in, k_out); // k_out = sorted keys from here on
	if (shard) CK(cudaMemcpyAsync(d.ids, id_in, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st)); // into the arena the peers have mapped
	else { d.ids = id_in; id_in = nullptr; }
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
	c->lap(bits == 2 ? "bd1_csr" : "bd2_csr", true);
	// table: nominal 2^k >= 2 numkeys slots (k fixes the home buckets), placed in mixed-key order
	u64 cap = 16;
	int kbits = 4;
	if (shard) {
		cap = shard->cap;
		kbits = 0;
		while ((1ull << kbits) < cap) kbits++;
		if ((u64)nk * 10 > cap * 9) { harcgpu_set_error("dictionary shard overflow: %u keys for %llu slots", nk, cap); return -1; }
	} else {
		while (cap < 2ull * nk) { cap <<= 1; kbits++; }
	}
	d.slot_shift = 64 - kbits;
	const int world = shard ? shard->world : 0;
	long long *pm = nullptr, *bmax = nullptr;
	u32 *pos = nullptr;
	unsigned long long *d_last = nullptr, h_last = 0;
	const u32 pmb = cdiv(nk, PM_TILE);
	if (c->alloc(&pm, nk) || c->alloc(&bmax, pmb) || c->alloc(&pos, nk) || c->alloc(&d_last, 1)) return -1;
	pm_local_kernel<<<KL + pmb, 256, 0, st>>>(d.keys, nk, d.slot_shift, world, pm, bmax);
	pm_blocks_kernel<<<KL + 1, 1024, 0, st>>>(bmax, pmb);
	pm_finish_kernel<<<KL + cdiv(nk, 256), 256, 0, st>>>(pm, bmax, nk, pos, d_last);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(&h_last, d_last, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	const u64 need = std::max<u64>(cap, h_last + 1) + 2; // two empty slots end every probe sequence
	if (shard) {
		if (need > shard->nslots) { harcgpu_set_error("dictionary shard overflow: the table needs %llu slots, the arena has %llu", need, shard->nslots); return -1; }
		d.nslots = shard->nslots;
	} else {
		d.nslots = need;
		if (c->alloc(&d.slots, need)) return -1;
	}
	CK(cudaMemsetAsync(d.slots, 0, d.nslots * sizeof(ulonglong2), st));
	place_kernel<<<KL + cdiv(nk, 256), 256, 0, st>>>(d.keys, d.start, d.ids, pos, nk, d.slots);
	overflow_kernel<<<KL + cdiv(nk, 256), 256, 0, st>>>(d.keys, pos, nk, d.slot_shift, world, d.slots);
	CK(cudaGetLastError());
	CK(cudaStreamSynchronize(st));
	c->lap(bits == 2 ? "bd1_place" : "bd2_place", true);
	c->release(pm); c->release(bmax); c->release(pos); c->release(d_last);
	c->release(k_in); c->release(k_out); c->release(id_in); c->release(head); c->release(binidx);
	c->release(scan_tmp); c->release(d_total); c->release(id_alt);
	return 0;
}
