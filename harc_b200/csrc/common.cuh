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

// splitmix64 finaliser: slot hash of the key-storing table that stands in for BooPHF (BooPHF.h:970-1008)
__device__ __forceinline__ u64 mix64(u64 x)
{
	x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
	x ^= x >> 27; x *= 0x94d049bb133111ebull;
	x ^= x >> 31;
	return x;
}

// One dictionary: canonical CSR (keys ascending, ids ascending inside a bin: reorder.cpp:344-391) plus an
// open-addressing table key -> (bin start, bin size).  A slot is 16 bytes {key, start | size<<32}; size==0 = empty.
struct DictDev {
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
};

__device__ __forceinline__ bool dict_lookup(const DictView &d, u64 key, u32 &start, u32 &size)
{
	u32 h = (u32)mix64(key) & d.slot_mask;
	while (true) {
		ulonglong2 s = __ldg(&d.slots[h]);
		u32 sz = (u32)(s.y >> 32);
		if (sz == 0) return false;
		if (s.x == key) { start = (u32)s.y; size = sz; return true; }
		h = (h + 1) & d.slot_mask;
	}
}

// ---- hand-written device-wide exclusive scan (scan.cu) ------------------------------------------------------
// out[i] = sum_{k<i} in[k]; *total (device pointer, may be null) = sum of all.  tmp must hold scan_tmp_elems(n) u64.
size_t scan_tmp_elems(size_t n);
int exclusive_scan_u32(const u32 *in, u32 *out, size_t n, u64 *tmp, u32 *total, cudaStream_t st);
int exclusive_scan_u64(const u64 *in, u64 *out, size_t n, u64 *tmp, u64 *total, cudaStream_t st);
