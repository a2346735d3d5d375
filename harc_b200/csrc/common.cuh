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

// base t: 2-bit code c -> 3-bit code 2c at bits 3t (nb <= 21 bases): the groups are moved apart in five doubling steps
__device__ __forceinline__ u64 spread2to3(u64 k2, int nb)
{
	u64 x = nb < 32 ? k2 & ((1ull << (2 * nb)) - 1) : k2;
	// (masks generated and checked against the per-base loop for every nb <= 21)
	x = (x & 0x00000000ffffffffull) | ((x & 0x000003ff00000000ull) << 16); // bases 16..20 move by 16
	x = (x & 0x03ff00000000ffffull) | ((x & 0x00000000ffff0000ull) << 8);  // bases with bit 3 set move by 8
	x = (x & 0x00ff0000ff0000ffull) | ((x & 0x030000ff0000ff00ull) << 4);  // bit 2: by 4
	x = (x & 0x300f00f00f00f00full) | ((x & 0x00f00f00f00f00f0ull) << 2);  // bit 1: by 2
	x = (x & 0x30c30c30c30c30c3ull) | ((x & 0x030c30c30c30c30cull) << 1);  // bit 0: by 1
	return x << 1;
}

// ---- the key table that stands in for BooPHF (BooPHF.h:970-1008) ------------------------------------------------
// A key is mixed by one 64-bit multiply (a bijection, so equal keys <=> equal mixed keys); the table is ORDERED by the
// mixed key: the home bucket of a key is the top bits of its mixed value, and the bins are placed in mixed-key order
// by linear probing -- which, for sorted input, is a prefix maximum (stage1.cu: place_kernel) instead of one atomic
// per bin.  The mixed key is also what the slots store and what lookups compare.
constexpr u64 KEY_MIX = 0x9E3779B97F4A7C15ull;   // 2^64 / golden ratio, odd
constexpr u64 KEY_UNMIX = 0xF1DE83E19937733Dull; // its inverse modulo 2^64
__device__ __host__ __forceinline__ u64 key_mix(u64 key) { return key * KEY_MIX; }
__device__ __host__ __forceinline__ u64 key_unmix(u64 t) { return t * KEY_UNMIX; }
// One job on several GPUs with sharded dictionaries: shard = floor(t * world / 2^64); inside a shard t * world (mod 2^64)
// runs over the whole 64-bit range again, monotonically, and takes the place of t for the home bucket.
__device__ __forceinline__ u32 mix_shard(u64 t, int world) { return (u32)__umul64hi(t, (u64)world); }
// home bucket (even slot) of mixed key t in a table of 2^(64 - shift) nominal slots
__device__ __forceinline__ u32 slot_home(u64 t, int shift, int world) { return (u32)((world > 1 ? t * (u64)world : t) >> shift) & ~1u; }

// One job on several GPUs: blocked Bloom filter over the keys of all shards of both dictionaries, one segment per shard
// (built by the shard's owner, copied to every GPU): word and the two bits of mixed key t of dictionary l.
__device__ __forceinline__ void job_bloom_pos(u64 t, int l, int world, u32 seg_words, u32 &word, u32 &bits)
{
	u64 h = (t ^ (l ? 0x9E3779B97F4A7C15ull : 0ull)) * 0xD6E8FEB86659FD93ull;
	h ^= h >> 32;
	word = mix_shard(t, world) * seg_words + ((u32)(h >> 10) & (seg_words - 1u));
	bits = (1u << ((u32)h & 31u)) | (1u << (((u32)h >> 5) & 31u));
}

// One dictionary: CSR over the bins in mixed-key order (ids ascending inside a bin: reorder.cpp:344-391) plus the
// open-addressing table mixed key -> bin.  A slot is 16 bytes {mixed key, val}: val == 0 = empty, else size = bits 32..62
// and the low half is the bin's first index into ids[] -- or, for a bin of one read (most bins), the read id itself, so
// that the common probe needs no second dependent load.
struct DictDev {
	bool external = false;    // slots and ids live in the shard arena (one job on several GPUs): not owned
	u64 *keys = nullptr;      // [numkeys] mixed keys, ascending
	u32 *start = nullptr;     // [numkeys+1]
	u32 *ids = nullptr;       // [n]
	ulonglong2 *slots = nullptr;
	u32 numkeys = 0;
	int slot_shift = 60;      // home bucket = (mixed key >> slot_shift) & ~1
	u64 nslots = 0;           // allocated slots (nominal 2^(64 - slot_shift) + spill + 2 empty ones at the end)
	int bitpos = 0, nbits = 0; // key = bits [bitpos, bitpos+nbits) of the packed read
};
struct DictView {
	const ulonglong2 *slots;
	const u32 *ids;
	int slot_shift;
	int dstart, dend; // in bases
	// one job on several GPUs with sharded dictionaries: shard s of the table and of the id lists (peer memory), all
	// shards with the same slot_shift; world == 0 otherwise
	const ulonglong2 *sslots[8];
	const u32 *sids[8];
	int world;
};

// Slots are probed in buckets of two (one 32-byte sector).  A bin that did not fit into its home bucket sets the
// OVERFLOW bit (bit 63 of val) of the home bucket's first slot, so a lookup that misses in the home bucket only goes on
// when that bit is set: ~95 % of all lookups, present or absent, end after one 32-byte load.  A lookup that goes on
// walks towards higher slots until it meets an empty one; the table ends with two empty slots and never wraps.
// val = overflow << 63 | size << 32 | lo; size == 0 <=> empty slot.
constexpr u64 SLOT_OVERFLOW = 1ull << 63;
__device__ __forceinline__ u32 slot_size(u64 val) { return (u32)(val >> 32) & 0x7fffffffu; }

// One step of a lookup of mixed key t over the bucket (s0, s1); `home` = this is the key's home bucket.
// returns 1: found (lo, size set), 0: absent, 2: go on with the next bucket
__device__ __forceinline__ int bucket_step(u64 t, ulonglong2 s0, ulonglong2 s1, bool home, u32 &lo, u32 &size)
{
	const u32 z0 = slot_size(s0.y), z1 = slot_size(s1.y);
	if (z0 != 0u && s0.x == t) { lo = (u32)s0.y; size = z0; return 1; }
	if (z0 != 0u && z1 != 0u && s1.x == t) { lo = (u32)s1.y; size = z1; return 1; }
	if (home) return (s0.y & SLOT_OVERFLOW) ? 2 : 0;
	return (z0 != 0u && z1 != 0u) ? 2 : 0;
}
// true if the key is present: size = reads in the bin, lo = first index into ids[] (size > 1) or the read id (size == 1)
// (unsharded tables only: stage II)
__device__ __forceinline__ bool dict_lookup(const DictView &d, u64 key, u32 &lo, u32 &size)
{
	const u64 t = key_mix(key);
	u32 h = slot_home(t, d.slot_shift, 0);
	bool home = true;
	while (true) {
		const int r = bucket_step(t, __ldg(&d.slots[h]), __ldg(&d.slots[h + 1]), home, lo, size);
		if (r != 2) return r == 1;
		home = false;
		h += 2;
	}
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
