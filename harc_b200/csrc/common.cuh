// Shared declarations of libharcgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include "../../include/harcgpu.h"

typedef unsigned long long u64;
typedef uint32_t u32;
typedef uint8_t u8;

void harcgpu_set_error(const char *fmt, ...);

#define CK(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess) {                                                                   \
			harcgpu_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			return -1;                                                                             \
		}                                                                                          \
	} while (0)

// every kernel launch of this library bumps the counter (bench.py reports it as gpu_launches)
extern unsigned long long g_harcgpu_launches;
#define KL (g_harcgpu_launches++, 0u)

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// ---- bit helpers shared by the kernels -------------------------------------------------------------------
// bits [pos, pos+n) of a little-endian word array (bitset & mask >> start, to_ullong), n in 1..64
__device__ __forceinline__ u64 getbits(const u64 *w, int words, int pos, int n)
{
	int q = pos >> 6, r = pos & 63;
	u64 v = w[q] >> r;
	if (r && q + 1 < words) v |= w[q + 1] << (64 - r);
	if (n < 64) v &= (1ull << n) - 1;
	return v;
}
// mask with the low `n` bits set, n clamped to [0,64]
__device__ __forceinline__ u64 lowmask(int n) { return n >= 64 ? ~0ull : (n <= 0 ? 0ull : ((1ull << n) - 1)); }

// Slot hash of the key-storing table that stands in for BooPHF (BooPHF.h:970-1008): one 64-bit multiply, the high
// half (the well-mixed one) folded with the low half.
__device__ __forceinline__ u32 slot_hash(u64 x)
{
	x *= 0x9E3779B97F4A7C15ull;
	return (u32)(x >> 32) ^ (u32)x;
}
struct DictView;

// One dictionary: canonical CSR (keys ascending, ids ascending inside a bin: reorder.cpp:344-391) plus an
// open-addressing table key -> bin.  A slot is 16 bytes {key, val}: val == 0 = empty, else size = bits 32..62 and the low
// half is the bin's first index into ids[] -- or, for a bin of one read (most bins), the read id itself, so that the
// common probe needs no second dependent load.
struct DictDev {
	bool external = false;    // slots and ids live in the shard arena (one job on several GPUs): not owned
	u64 *keys = nullptr;      // [numkeys] ascending
	u32 *start = nullptr;     // [numkeys+1]
	u32 *ids = nullptr;       // [n]
	ulonglong2 *slots = nullptr;
	u32 numkeys = 0;
	u32 slot_mask = 0;
	int bitpos = 0, nbits = 0; // key = bits [bitpos, bitpos+nbits) of the packed read
};
struct DictView {
	const ulonglong2 *slots;
	const u32 *ids;
	u32 slot_mask;
	int dstart, dend; // in bases
	// one job on several GPUs with sharded dictionaries: shard s of the table and of the id lists (peer memory), every
	// shard with slot_mask + 1 slots; world == 0 otherwise
	const ulonglong2 *sslots[8];
	const u32 *sids[8];
	int world;
};
// 64-bit product behind slot_hash; the shard of a key comes from bits the slot index does not use
__device__ __host__ __forceinline__ u64 key_mix(u64 x) { return x * 0x9E3779B97F4A7C15ull; }
__device__ __host__ __forceinline__ u32 mix_slot(u64 m) { return (u32)(m >> 32) ^ (u32)m; }
__device__ __host__ __forceinline__ u32 mix_shard(u64 m, int world) { return (u32)((((m >> 40) & 0xffffffull) * (u64)world) >> 24); }

// Slots are probed in buckets of two (one 32-byte sector): a key hashes to an even slot (its home bucket) and is inserted
// into the first empty slot from there on.  A key that did not fit into its home bucket sets the OVERFLOW bit (bit 63 of
// val) of the home bucket's first slot, so a lookup that misses in the home bucket only goes on when that bit is set:
// ~95 % of all lookups, present or absent, end after one 32-byte load.
// val = overflow << 63 | size << 32 | lo; size == 0 <=> empty slot.
constexpr u64 SLOT_OVERFLOW = 1ull << 63;
__device__ __forceinline__ u32 slot_size(u64 val) { return (u32)(val >> 32) & 0x7fffffffu; }

// One step of a lookup over the bucket (s0, s1); `home` = this is the key's home bucket.
// returns 1: found (lo, size set), 0: absent, 2: go on with the next bucket
__device__ __forceinline__ int bucket_step(u64 key, ulonglong2 s0, ulonglong2 s1, bool home, u32 &lo, u32 &size)
{
	const u32 z0 = slot_size(s0.y), z1 = slot_size(s1.y);
	if (z0 != 0u && s0.x == key) { lo = (u32)s0.y; size = z0; return 1; }
	if (z0 != 0u && z1 != 0u && s1.x == key) { lo = (u32)s1.y; size = z1; return 1; }
	if (home) return (s0.y & SLOT_OVERFLOW) ? 2 : 0;
	return (z0 != 0u && z1 != 0u) ? 2 : 0;
}
// true if the key is present: size = reads in the bin, lo = first index into ids[] (size > 1) or the read id (size == 1)
__device__ __forceinline__ bool dict_resolve(const DictView &d, u64 key, u32 h, ulonglong2 s0, ulonglong2 s1, u32 &lo, u32 &size)
{
	bool home = true;
	while (true) {
		const int r = bucket_step(key, s0, s1, home, lo, size);
		if (r != 2) return r == 1;
		home = false;
		h = (h + 2) & d.slot_mask;
		s0 = __ldg(&d.slots[h]);
		s1 = __ldg(&d.slots[h + 1]);
	}
}
__device__ __forceinline__ bool dict_lookup(const DictView &d, u64 key, u32 &lo, u32 &size)
{
	const u32 h = slot_hash(key) & d.slot_mask & ~1u;
	return dict_resolve(d, key, h, __ldg(&d.slots[h]), __ldg(&d.slots[h + 1]), lo, size);
}
// entry t (0 = lowest id) of a bin returned by dict_lookup
__device__ __forceinline__ u32 bin_entry(const DictView &d, u32 lo, u32 size, u32 t)
{
	return size == 1 ? lo : __ldg(&d.ids[lo + t]);
}

// ---- hand-written device-wide exclusive scan (scan.cu) ------------------------------------------------------
// out[i] = sum_{k<i} in[k]; *total (device pointer, may be null) = sum of all.  tmp must hold scan_tmp_elems(n) u64.
size_t scan_tmp_elems(size_t n);
int exclusive_scan_u32(const u32 *in, u32 *out, size_t n, u64 *tmp, u32 *total, cudaStream_t st);
int exclusive_scan_u64(const u64 *in, u64 *out, size_t n, u64 *tmp, u64 *total, cudaStream_t st);
