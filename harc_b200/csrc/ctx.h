// Internal context of libharcgpu: everything behind the opaque harcgpu_ctx of include/harcgpu.h.
#pragma once
#include "common.cuh"
#include <map>
#include <vector>

struct DevBuf {
	void *p = nullptr;
	size_t bytes = 0;
};

struct SetOut { // one read_*.txt.<k> file set, device resident
	u8 *seq = nullptr; u64 seq_bytes = 0; char seq_tail[4]; u32 seq_ntail = 0;
	u8 *pos = nullptr; u64 pos_bytes = 0;
	char *noise = nullptr; u64 noise_bytes = 0;
	u8 *noisepos = nullptr; u64 noisepos_bytes = 0;
	u8 *rev = nullptr; u64 rev_bytes = 0; char rev_tail[8]; u32 rev_ntail = 0;
};

struct harcgpu_ctx {
	int device = 0;
	harcgpu_params p;
	cudaStream_t st = nullptr;
	int L = 0, NW = 0, NW3 = 0;
	std::vector<void *> allocs;
	std::map<std::string, double> ms;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;

	// ---- stage I
	u32 n = 0;
	u64 *reads = nullptr;   // [n][NW] 2-bit packed (reorder.cpp:188-195)
	DictDev d1[2];
	bool dicts_built = false;
	u32 *claim = nullptr;   // bitmap, 1 = still unclaimed (remainingreads, reorder.cpp:449)
	long long *gpos = nullptr;
	u64 *counters = nullptr; // 8 x u64
	u32 walkers_used = 0;
	// finalized stage I streams (device)
	bool reordered = false;
	u32 n_matched = 0, n_single = 0, n_unmatched = 0;
	u32 *order = nullptr, *order_s = nullptr;
	u8 *rev = nullptr, *flag = nullptr, *pos = nullptr;

	// ---- stage II inputs
	bool stream_set = false;
	u32 m = 0;              // reads in the reordered stream
	u64 *sreads = nullptr;  // [m][NW] stream reads, already reverse-complemented where flagged
	u32 *s_order = nullptr;
	u8 *s_rev = nullptr, *s_flag = nullptr, *s_pos = nullptr;
	bool pool_set = false;
	u32 n_s = 0, n_N = 0;   // pool = n_s singletons ++ n_N reads with N
	u64 *pool = nullptr;    // [n_s+n_N][NW] 2-bit codes with N stored as 0
	u64 *poolN = nullptr;   // [n_s+n_N][NW] bit 2i set where base i is N; 3-bit code of encoder.cpp:731-745 = 2*code2 + nflag
	u32 *pool_order = nullptr; // order_s (encoder.cpp:865-870)
	DictDev d2[2];
	// ---- stage II outputs
	bool encoded = false;
	std::vector<SetOut> sets;
	std::vector<void *> s2_keep; // global stage II streams the per-set views point into
	harcgpu_encode_sizes esz;
	u32 *o_order = nullptr, *o_order_N = nullptr;
	u8 *o_single = nullptr; char single_tail[4];
	char *o_inputN = nullptr;

	// Stream-ordered allocations from the device's default pool (release threshold raised in harcgpu_create), so the
	// many short-lived work buffers of a pass cost no cudaMalloc/cudaFree round trips after the first pass.
	template <typename T> int alloc(T **out, size_t count)
	{
		void *q = nullptr;
		size_t bytes = (count ? count : 1) * sizeof(T);
		cudaError_t e = cudaMallocAsync(&q, bytes, st);
		if (e != cudaSuccess) {
			harcgpu_set_error("cudaMallocAsync(%zu bytes): %s", bytes, cudaGetErrorString(e));
			*out = nullptr;
			return -1;
		}
		allocs.push_back(q);
		*out = (T *)q;
		return 0;
	}
	void release(void *q)
	{
		if (!q) return;
		for (size_t i = 0; i < allocs.size(); i++)
			if (allocs[i] == q) { allocs[i] = allocs.back(); allocs.pop_back(); break; }
		cudaFreeAsync(q, st);
	}
	void tic() { cudaEventRecord(ev0, st); }
	void toc(const char *phase)
	{
		cudaEventRecord(ev1, st);
		cudaEventSynchronize(ev1);
		float f = 0;
		cudaEventElapsedTime(&f, ev0, ev1);
		ms[phase] = f;
	}
};

// stage1.cu
int s1_pack_reads(harcgpu_ctx *c, const void *d_ascii, u32 n);
int s1_packN(harcgpu_ctx *c, const void *d_ascii, u32 n, u64 *out2, u64 *outN);
int build_dict(harcgpu_ctx *c, DictDev &d, const u64 *reads, const u64 *readsN, u32 n, int words, int ds, int de, int bits);
void free_dict(harcgpu_ctx *c, DictDev &d);
int s1_reorder(harcgpu_ctx *c);
int s1_unpack_reads(harcgpu_ctx *c, const u64 *reads, const u32 *order, const u8 *rev, u32 cnt, char *d_out);
// stage2.cu
int s2_set_stream_from_stage1(harcgpu_ctx *c);
int s2_set_stream_host(harcgpu_ctx *c, const char *dna, const char *flag, const u8 *pos, const u32 *order, const char *rev, u32 n);
int s2_load_pool(harcgpu_ctx *c, const char *s_ascii, const u32 *order_s, u32 n_s, const char *N_ascii, u32 n_N);
int s2_load_pool_dev(harcgpu_ctx *c, const void *d_N_ascii, u32 n_N);
int s2_encode(harcgpu_ctx *c);
