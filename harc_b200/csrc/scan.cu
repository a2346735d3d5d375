// Device-wide exclusive prefix sum: warp-shuffle scans inside a 2048-element tile, tile sums scanned recursively.
// Used for bucket offsets (reorder.cpp:366-367 cumulative startpos), stream compaction and every variable-length
// output of stage II.  Streaming and HBM-bound: reads n, writes n (+ n/2048 tile sums).
#include "common.cuh"

namespace {
constexpr int TPB = 256, IPT = 8, TILE = TPB * IPT;

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v, int lane)
{
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		T o = __shfl_up_sync(0xffffffffu, v, d);
		if (lane >= d) v += o;
	}
	return v;
}

// Exclusive scan of one value per thread across the block; the block total comes back through `total`.
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T &total)
{
	__shared__ T wsum[TPB / 32 + 1];
	int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	T inc = warp_incl_scan(v, lane);
	if (lane == 31) wsum[w] = inc;
	__syncthreads();
	if (w == 0) {
		T s = lane < TPB / 32 ? wsum[lane] : 0;
		T si = warp_incl_scan(s, lane);
		if (lane < TPB / 32) wsum[lane] = si - s;
		if (lane == 31) wsum[TPB / 32] = si;
	}
	__syncthreads();
	total = wsum[TPB / 32];
	T r = wsum[w] + inc - v;
	__syncthreads();
	return r;
}

template <typename T>
__global__ void __launch_bounds__(TPB) tile_sums(const T *__restrict__ in, T *__restrict__ sums, size_t n)
{
	size_t base = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * IPT;
	T s = 0;
#pragma unroll
	for (int k = 0; k < IPT; k++) if (base + k < n) s += in[base + k];
	T total = 0;
	block_excl_scan(s, total);
	if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename T>
__global__ void __launch_bounds__(TPB) tile_scan(const T *__restrict__ in, T *__restrict__ out, const T *__restrict__ offs,
                                                 size_t n, T *__restrict__ grand)
{
	size_t base = (size_t)blockIdx.x * TILE + (size_t)threadIdx.x * IPT;
	T v[IPT];
	T s = 0;
#pragma unroll
	for (int k = 0; k < IPT; k++) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
	T total = 0;
	T ex = block_excl_scan(s, total) + (offs ? offs[blockIdx.x] : 0);
#pragma unroll
	for (int k = 0; k < IPT; k++) { if (base + k < n) out[base + k] = ex; ex += v[k]; }
	if (grand && blockIdx.x == gridDim.x - 1 && threadIdx.x == TPB - 1) *grand = ex;
}

template <typename T>
int scan_rec(const T *in, T *out, size_t n, T *tmp, T *total, cudaStream_t st)
{
	if (n == 0) {
		if (total) CK(cudaMemsetAsync(total, 0, sizeof(T), st));
		return 0;
	}
	size_t nt = (n + TILE - 1) / TILE;
	if (nt == 1) {
		tile_scan<T><<<KL + 1, TPB, 0, st>>>(in, out, nullptr, n, total);
		CK(cudaGetLastError());
		return 0;
	}
	T *sums = tmp, *rest = tmp + nt;
	tile_sums<T><<<KL + (unsigned)nt, TPB, 0, st>>>(in, sums, n);
	CK(cudaGetLastError());
	if (scan_rec<T>(sums, sums, nt, rest, nullptr, st)) return -1;
	tile_scan<T><<<KL + (unsigned)nt, TPB, 0, st>>>(in, out, sums, n, total);
	CK(cudaGetLastError());
	return 0;
}
} // namespace

size_t scan_tmp_elems(size_t n)
{
	size_t t = 0;
	while (n > (size_t)TILE) { n = (n + TILE - 1) / TILE; t += n; }
	return t + 8;
}
int exclusive_scan_u32(const u32 *in, u32 *out, size_t n, u64 *tmp, u32 *total, cudaStream_t st)
{
	return scan_rec<u32>(in, out, n, reinterpret_cast<u32 *>(tmp), total, st);
}
int exclusive_scan_u64(const u64 *in, u64 *out, size_t n, u64 *tmp, u64 *total, cudaStream_t st)
{
	return scan_rec<u64>(in, out, n, tmp, total, st);
}
