// C ABI of libharcgpu.so (include/harcgpu.h).  Host side only: argument checks, device memory, copies, timing.
#include "ctx.h"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <sys/stat.h>
#include <string>
#include <vector>

unsigned long long g_harcgpu_launches = 0;
static thread_local char g_err[1024] = "";
void harcgpu_set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
}

extern "C" {

const char *harcgpu_last_error(void) { return g_err; }
uint64_t harcgpu_launch_count(void) { return g_harcgpu_launches; }

int harcgpu_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

// harc:52-63
int harcgpu_default_params(int L, harcgpu_params *p)
{
	if (!p || L < 1 || L > 255) { harcgpu_set_error("readlen must be in 1..255"); return -1; }
	memset(p, 0, sizeof *p);
	p->readlen = L;
	p->maxmatch = L / 2;
	p->thresh = 4;
	p->thresh_s = 24;
	p->numdict = 2;
	p->maxsearch = 1000;
	p->dict_start[0] = L > 100 ? L / 2 - 32 : L / 2 - L * 32 / 100;
	p->dict_end[0] = L / 2 - 1;
	p->dict_start[1] = L / 2;
	p->dict_end[1] = L > 100 ? L / 2 - 1 + 32 : L / 2 - 1 + L * 32 / 100;
	p->walkers = 0;
	p->file_sets = 1;
	return 0;
}

int harcgpu_create(int device, const harcgpu_params *p, harcgpu_ctx **out)
{
	if (!p || !out) { harcgpu_set_error("null argument"); return -1; }
	if (p->readlen < 1 || p->readlen > 255 || p->numdict < 1 || p->numdict > 2 || p->maxmatch < 0 || p->maxmatch > p->readlen) {
		harcgpu_set_error("bad parameters (readlen 1..255, numdict 1..2)");
		return -1;
	}
	for (int l = 0; l < p->numdict; l++)
		if (p->dict_start[l] < 0 || p->dict_end[l] < p->dict_start[l] || p->dict_end[l] >= p->readlen ||
		    p->dict_end[l] - p->dict_start[l] + 1 > 32) {
			harcgpu_set_error("dictionary %d window [%d,%d] invalid (at most 32 bases inside the read)", l, p->dict_start[l], p->dict_end[l]);
			return -1;
		}
	int ndev = harcgpu_device_count();
	if (ndev <= 0) { harcgpu_set_error("no CUDA device: libharcgpu has no CPU fallback"); return -1; }
	if (device < 0 || device >= ndev) { harcgpu_set_error("device %d out of range (%d devices)", device, ndev); return -1; }
	CK(cudaSetDevice(device));
	// tuning aids (profiles/): L2 fetch granularity for the random 32-byte bucket / read fetches of the walk
	if (const char *e = getenv("HARCGPU_L2FETCH")) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)));
	harcgpu_ctx *c = new harcgpu_ctx();
	c->device = device;
	c->p = *p;
	if (c->p.file_sets <= 0) c->p.file_sets = 1;
	c->L = p->readlen;
	c->NW = (2 * c->L + 63) / 64;
	c->NW3 = (3 * c->L + 63) / 64;
	memset(&c->esz, 0, sizeof c->esz);
	// a context that cannot be completed is taken down again (stream, events, blocks), not leaked
	auto fail = [&](cudaError_t e, const char *what) {
		if (e != cudaSuccess) harcgpu_set_error("harcgpu_create: %s: %s", what, cudaGetErrorString(e));
		harcgpu_destroy(c);
		return -1;
	};
	cudaError_t e;
	if ((e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking)) != cudaSuccess) { c->st = nullptr; return fail(e, "cudaStreamCreate"); }
	if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) { c->ev0 = nullptr; return fail(e, "cudaEventCreate"); }
	if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) { c->ev1 = nullptr; return fail(e, "cudaEventCreate"); }
	if (c->alloc(&c->counters, 8) || c->alloc(&c->gpos, 1)) return fail(cudaSuccess, "alloc");
	if ((e = cudaMemsetAsync(c->counters, 0, 64, c->st)) != cudaSuccess) return fail(e, "cudaMemsetAsync");
	*out = c;
	return 0;
}

void harcgpu_destroy(harcgpu_ctx *c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->st) cudaStreamSynchronize(c->st);
	job_close(c);
	if (c->st_bcast) {
		cudaStreamSynchronize(c->st_bcast); cudaStreamDestroy(c->st_bcast); cudaEventDestroy(c->ev_packed); cudaEventDestroy(c->ev_bcast);
		for (int r = 0; r < 8; r++) { cudaStreamSynchronize(c->st_peer[r]); cudaStreamDestroy(c->st_peer[r]); cudaEventDestroy(c->ev_peer[r]); }
	}
	if (c->st_copy) { cudaStreamSynchronize(c->st_copy); cudaStreamDestroy(c->st_copy); cudaEventDestroy(c->ev_staged); cudaEventDestroy(c->ev_order); }
	for (auto &b : c->live) cudaFree(b.p);
	c->trim();
	if (c->ev0) cudaEventDestroy(c->ev0);
	if (c->ev1) cudaEventDestroy(c->ev1);
	if (c->evl0) { cudaEventDestroy(c->evl0); cudaEventDestroy(c->evl1); }
	if (c->st) cudaStreamDestroy(c->st);
	delete c;
}

void *harcgpu_stream(harcgpu_ctx *c) { return c ? (void *)c->st : nullptr; }

int harcgpu_trim(harcgpu_ctx *c)
{
	if (!c) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	CK(cudaStreamSynchronize(c->st));
	c->trim();
	return 0;
}

double harcgpu_last_ms(harcgpu_ctx *c, const char *phase)
{
	if (!c || !phase) return -1;
	if (!strcmp(phase, "cudaMalloc_calls")) return (double)c->n_cuda_malloc;
	if (!strcmp(phase, "cached_MB")) return (double)c->cached_bytes / 1048576.0;
	if (!strcmp(phase, "peak_MB")) return (double)c->peak_bytes / 1048576.0;
	if (!strcmp(phase, "job_bloom_bytes")) return c->dicts_sharded ? (double)(c->shard_world - 1) * c->bloom_seg_words * 4.0 : 0.0;
	if (!strcmp(phase, "walkers")) return (double)c->walkers_used;
	auto it = c->ms.find(phase);
	return it == c->ms.end() ? -1.0 : it->second;
}

static int reset_stage1(harcgpu_ctx *c, u32 n)
{
	if (c->arena[c->shard_rank]) job_close(c); // the context leaves a job on several GPUs: its reads lived in the arena
	c->release(c->reads); c->release(c->claim); c->release(c->bloom1);
	c->reads = nullptr; c->claim = nullptr; c->bloom1 = nullptr; c->bloom1_words = 0;
	for (int l = 0; l < 2; l++) free_dict(c, c->d1[l]);
	// everything derived from the reads of before is stale now, stage II inputs and outputs included
	c->dicts_built = false; c->reordered = false; c->stream_set = false; c->pool_set = false; c->encoded = false;
	c->n = n;
	if (c->alloc(&c->reads, (size_t)n * c->NW) || c->alloc(&c->claim, ((size_t)n + 31) / 32)) return -1;
	return 0;
}

int harcgpu_load_reads_device(harcgpu_ctx *c, const void *d_ascii, u32 n)
{
	if (!c || (!d_ascii && n)) { harcgpu_set_error("null argument"); return -1; }
	if ((uintptr_t)d_ascii & 15) { harcgpu_set_error("harcgpu_load_reads_device: the buffer must be 16-byte aligned"); return -1; }
	CK(cudaSetDevice(c->device));
	if (reset_stage1(c, n)) return -1;
	c->tic();
	if (s1_pack_reads(c, d_ascii, n)) return -1;
	c->toc("pack");
	return 0;
}

static int ensure_copy_stream(harcgpu_ctx *c)
{
	if (!c->st_copy) {
		CK(cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking));
		CK(cudaEventCreateWithFlags(&c->ev_staged, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&c->ev_order, cudaEventDisableTiming));
	}
	return 0;
}

// The upload is cut into chunks on the copy stream and every chunk is packed as soon as it has arrived, so the pack runs
// under the copy (and the copies of several contexts that share the link interleave chunk by chunk).
int harcgpu_load_reads(harcgpu_ctx *c, const char *ascii, u32 n)
{
	if (!c || (!ascii && n)) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	if (ensure_copy_stream(c)) return -1;
	char *d = nullptr;
	const size_t line = (size_t)c->L + 1, bytes = (size_t)n * line;
	if (c->alloc(&d, bytes + 16)) return -1;
	if (reset_stage1(c, n)) return -1;
	// the staging block may still be in use by work queued on the compute stream: order the copies behind it
	CK(cudaEventRecord(c->ev_order, c->st));
	CK(cudaStreamWaitEvent(c->st_copy, c->ev_order, 0));
	c->tic();
	const u32 chunk = 1u << 18; // reads per chunk: a multiple of 64, so every chunk starts on a 16-byte boundary
	for (u32 r0 = 0; r0 < n; r0 += chunk) {
		const u32 nr = std::min<u32>(chunk, n - r0);
		CK(cudaMemcpyAsync(d + (size_t)r0 * line, ascii + (size_t)r0 * line, (size_t)nr * line, cudaMemcpyHostToDevice, c->st_copy));
		CK(cudaEventRecord(c->ev_order, c->st_copy));
		CK(cudaStreamWaitEvent(c->st, c->ev_order, 0));
		if (s1_pack_reads_to(c, d + (size_t)r0 * line, nr, c->reads + (size_t)r0 * c->NW)) return -1;
	}
	c->toc("pack");
	c->release(d);
	return 0;
}

// ---- fused ingest: preprocess.cpp:49-138 + reorder.cpp:240-263 ------------------------------------------------
int harcgpu_fastq_readlen(const char *fastq, uint64_t nbytes)
{
	// harc:44  readlen=$(head -2 $filename | tail -1 | wc -L)
	if (!fastq) return -1;
	const char *end = fastq + nbytes;
	const char *l1 = (const char *)memchr(fastq, '\n', nbytes);
	if (!l1) return -1;
	l1++;
	const char *l2 = (const char *)memchr(l1, '\n', (size_t)(end - l1));
	return (int)((l2 ? l2 : end) - l1);
}

int harcgpu_ingest_fastq_device(harcgpu_ctx *c, const void *d_fastq, uint64_t nbytes, harcgpu_ingest_info *info)
{
	if (!c || (!d_fastq && nbytes)) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	if (c->arena[c->shard_rank]) job_close(c);
	for (int l = 0; l < 2; l++) free_dict(c, c->d1[l]);
	c->dicts_built = false; c->reordered = false; c->stream_set = false; c->pool_set = false; c->encoded = false;
	u64 total = 0;
	u32 nc = 0, nn = 0;
	if (ing_ingest(c, (const char *)d_fastq, nbytes, &total, &nc, &nn)) return -1;
	if (info) { info->readlen = (uint32_t)c->L; info->total_reads = total; info->n_clean = nc; info->n_N = nn; }
	return 0;
}

int harcgpu_ingest_fastq(harcgpu_ctx *c, const char *fastq, uint64_t nbytes, harcgpu_ingest_info *info)
{
	if (!c || (!fastq && nbytes)) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	char *d = nullptr;
	if (c->alloc(&d, nbytes + 16)) return -1;
	CK(cudaMemcpyAsync(d, fastq, nbytes, cudaMemcpyHostToDevice, c->st));
	int rc = harcgpu_ingest_fastq_device(c, d, nbytes, info);
	CK(cudaStreamSynchronize(c->st));
	c->release(d);
	return rc;
}

int harcgpu_get_ingest(harcgpu_ctx *c, char *input_clean, char *input_N, uint32_t *order_N)
{
	if (!c || !c->reads) { harcgpu_set_error("nothing ingested"); return -1; }
	CK(cudaSetDevice(c->device));
	const size_t line = (size_t)c->L + 1;
	if (input_clean && c->n) {
		char *d = nullptr;
		if (c->alloc(&d, (size_t)c->n * line)) return -1;
		if (ing_unpack_clean(c, d)) return -1;
		CK(cudaMemcpyAsync(input_clean, d, (size_t)c->n * line, cudaMemcpyDeviceToHost, c->st));
		CK(cudaStreamSynchronize(c->st));
		c->release(d);
	}
	if (input_N && c->ing_nN) CK(cudaMemcpyAsync(input_N, c->ing_N, (size_t)c->ing_nN * line, cudaMemcpyDeviceToHost, c->st));
	if (order_N && c->ing_nN) CK(cudaMemcpyAsync(order_N, c->ing_orderN, 4 * (size_t)c->ing_nN, cudaMemcpyDeviceToHost, c->st));
	CK(cudaStreamSynchronize(c->st));
	return 0;
}

int harcgpu_load_pool_ingested(harcgpu_ctx *c)
{
	if (!c) { harcgpu_set_error("null argument"); return -1; }
	if (!c->reordered) { harcgpu_set_error("harcgpu_load_pool_ingested needs the singletons of harcgpu_reorder on this context"); return -1; }
	if (c->ing_nN && !c->ing_N) { harcgpu_set_error("nothing ingested"); return -1; }
	CK(cudaSetDevice(c->device));
	return s2_load_pool_dev(c, c->ing_N, c->ing_nN);
}

int harcgpu_build_dicts(harcgpu_ctx *c)
{
	if (!c || !c->reads) { harcgpu_set_error("load reads first"); return -1; }
	CK(cudaSetDevice(c->device));
	if (c->shard_world > 1) return harcgpu_job_build_dicts(c);
	for (const char *k : { "lap:bd1_keys", "lap:bd1_sort", "lap:bd1_csr", "lap:bd1_place" }) c->ms.erase(k);
	c->tic();
	for (int l = 0; l < c->p.numdict; l++)
		if (build_dict(c, c->d1[l], c->reads, nullptr, c->n, c->NW, c->p.dict_start[l], c->p.dict_end[l], 2, nullptr)) return -1;
	// Most probes of the walk ask for keys that are in no dictionary.  A blocked Bloom filter over both dictionaries, small
	// enough to live in L2, answers those without a trip to the tables in HBM (configs[1]: 40.5 -> 13.8 GB of DRAM traffic
	// per walk, 21.9 -> 20.9 ms; profiles/r2_walk_experiments.md).
	c->release(c->bloom1);
	c->bloom1 = nullptr; c->bloom1_words = 0;
	{
		// 8 bits per key if that fits, never more than 32 MiB (what stays in the 126 MB L2 next to the streams of the walk),
		// and no filter at all below HARCGPU_BLOOM1_BITS (default 3) bits per key: a filter in HBM would only add a trip
		int min_bits = 3;
		if (const char *e = getenv("HARCGPU_BLOOM1_BITS")) min_bits = atoi(e);
		const u64 nk = (u64)c->d1[0].numkeys + (c->p.numdict > 1 ? c->d1[1].numkeys : 0);
		u64 words = 1024;
		while (words * 32 < 8 * nk && words < (1ull << 23)) words <<= 1;
		if (min_bits > 0 && nk > 0 && words * 32 >= (u64)min_bits * nk) {
			if (c->alloc(&c->bloom1, words)) return -1;
			c->bloom1_words = (u32)words;
			CK(cudaMemsetAsync(c->bloom1, 0, words * 4, c->st));
			for (int l = 0; l < c->p.numdict; l++)
				if (job_bloom_insert(c, c->d1[l].keys, c->d1[l].numkeys, l, 1, c->bloom1_words, c->bloom1)) return -1;
		}
	}
	c->toc("dict");
	c->dicts_built = true;
	return 0;
}

int harcgpu_dump_dict(harcgpu_ctx *c, int stage, int l, uint64_t *keys, uint32_t *counts, uint32_t *ids, uint32_t *numkeys, uint32_t *nids)
{
	if (!c || l < 0 || l > 1 || (stage != 1 && stage != 2)) { harcgpu_set_error("bad argument"); return -1; }
	CK(cudaSetDevice(c->device));
	DictDev &d = stage == 1 ? c->d1[l] : c->d2[l];
	u32 n = stage == 1 ? c->n : c->n_s + c->n_N;
	if (!d.ids) { harcgpu_set_error("dictionary not built"); return -1; }
	if (numkeys) *numkeys = d.numkeys;
	if (nids) *nids = n;
	if (!keys && !ids && !counts) return 0;
	// The dictionary lives in mixed-key order (common.cuh); the canonical view is by key: un-mix and sort on the host
	// (a test hook, not on the hot path).
	const size_t nk = d.numkeys;
	std::vector<u64> mk(nk);
	std::vector<u32> st(nk + 1), hid(n);
	CK(cudaMemcpy(mk.data(), d.keys, 8 * nk, cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(st.data(), d.start, 4 * (nk + 1), cudaMemcpyDeviceToHost));
	CK(cudaMemcpy(hid.data(), d.ids, 4 * (size_t)(nk ? st[nk] : 0), cudaMemcpyDeviceToHost));
	std::vector<u32> perm(nk);
	for (size_t i = 0; i < nk; i++) { perm[i] = (u32)i; mk[i] = key_unmix(mk[i]); }
	std::sort(perm.begin(), perm.end(), [&](u32 x, u32 y) { return mk[x] < mk[y]; });
	size_t w = 0;
	for (size_t i = 0; i < nk; i++) {
		const u32 b = perm[i];
		if (keys) keys[i] = mk[b];
		if (counts) counts[i] = st[b + 1] - st[b];
		if (ids) for (u32 k = st[b]; k < st[b + 1]; k++) ids[w++] = hid[k];
	}
	return 0;
}

int harcgpu_debug_sort(harcgpu_ctx *c, uint64_t *keys, uint32_t *vals, uint64_t n, int mode, int begin_bit, int end_bit)
{
	if (!c || ((!keys || !vals) && n)) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	u64 *k = nullptr, *k2 = nullptr;
	u32 *v = nullptr, *v2 = nullptr;
	if (c->alloc(&k, n) || c->alloc(&k2, n) || c->alloc(&v, n) || c->alloc(&v2, n)) return -1;
	CK(cudaMemcpyAsync(k, keys, 8 * n, cudaMemcpyHostToDevice, c->st));
	CK(cudaMemcpyAsync(v, vals, 4 * n, cudaMemcpyHostToDevice, c->st));
	int rc = mode == 1 ? radix_sort_mixed(c, (u64 **)&k, (u64 **)&k2, &v, &v2, n)
	       : mode == 2 ? radix_sort_pairs(c, (u64 **)&k, (u64 **)&k2, &v, &v2, n, begin_bit, end_bit)
	                   : radix_sort_pairs(c, (u64 **)&k, (u64 **)&k2, &v, &v2, n, 0, 64);
	if (!rc) {
		CK(cudaMemcpyAsync(keys, k, 8 * n, cudaMemcpyDeviceToHost, c->st));
		CK(cudaMemcpyAsync(vals, v, 4 * n, cudaMemcpyDeviceToHost, c->st));
		CK(cudaStreamSynchronize(c->st));
	}
	c->release(k); c->release(k2); c->release(v); c->release(v2);
	return rc;
}

int harcgpu_reorder(harcgpu_ctx *c)
{
	if (!c || !c->reads) { harcgpu_set_error("load reads first"); return -1; }
	CK(cudaSetDevice(c->device));
	if (!c->dicts_built && harcgpu_build_dicts(c)) return -1;
	if (c->shard_world > 1) return harcgpu_job_reorder(c);
	return s1_reorder(c);
}

int harcgpu_reorder_counts(harcgpu_ctx *c, uint32_t *nm, uint32_t *ns, uint32_t *nu)
{
	if (!c || !c->reordered) { harcgpu_set_error("reorder first"); return -1; }
	if (nm) *nm = c->n_matched;
	if (ns) *ns = c->n_single;
	if (nu) *nu = c->n_unmatched;
	return 0;
}

int harcgpu_get_reorder(harcgpu_ctx *c, uint32_t *order, char *rev, char *flag, uint8_t *pos, uint32_t *order_s)
{
	if (!c || !c->reordered) { harcgpu_set_error("reorder first"); return -1; }
	CK(cudaSetDevice(c->device));
	size_t m = c->n_matched;
	if (order) CK(cudaMemcpyAsync(order, c->order, 4 * m, cudaMemcpyDeviceToHost, c->st));
	if (rev) CK(cudaMemcpyAsync(rev, c->rev, m, cudaMemcpyDeviceToHost, c->st));
	if (flag) CK(cudaMemcpyAsync(flag, c->flag, m, cudaMemcpyDeviceToHost, c->st));
	if (pos) CK(cudaMemcpyAsync(pos, c->pos, m, cudaMemcpyDeviceToHost, c->st));
	if (order_s) CK(cudaMemcpyAsync(order_s, c->order_s, 4 * (size_t)c->n_single, cudaMemcpyDeviceToHost, c->st));
	CK(cudaStreamSynchronize(c->st));
	return 0;
}

int harcgpu_get_reordered_reads(harcgpu_ctx *c, char *temp_dna, char *temp_dna_singleton)
{
	if (!c || !c->reordered) { harcgpu_set_error("reorder first"); return -1; }
	CK(cudaSetDevice(c->device));
	size_t line = (size_t)c->L + 1;
	if (temp_dna && c->n_matched) {
		char *d = nullptr;
		if (c->alloc(&d, line * c->n_matched)) return -1;
		if (s1_unpack_reads(c, c->reads, c->order, c->rev, c->n_matched, d)) return -1;
		CK(cudaMemcpyAsync(temp_dna, d, line * c->n_matched, cudaMemcpyDeviceToHost, c->st));
		CK(cudaStreamSynchronize(c->st));
		c->release(d);
	}
	if (temp_dna_singleton && c->n_single) {
		char *d = nullptr;
		if (c->alloc(&d, line * c->n_single)) return -1;
		if (s1_unpack_reads(c, c->reads, c->order_s, nullptr, c->n_single, d)) return -1;
		CK(cudaMemcpyAsync(temp_dna_singleton, d, line * c->n_single, cudaMemcpyDeviceToHost, c->st));
		CK(cudaStreamSynchronize(c->st));
		c->release(d);
	}
	return 0;
}

int harcgpu_device_result(harcgpu_ctx *c, const char *name, const void **ptr, uint64_t *count)
{
	if (!c || !name || !ptr || !count) { harcgpu_set_error("null argument"); return -1; }
	if (!strcmp(name, "singleton_ids")) {
		if (!c->reordered) { harcgpu_set_error("reorder first"); return -1; }
		*ptr = c->order_s; *count = c->n_single;
	} else if (!strcmp(name, "order")) {
		if (!c->reordered) { harcgpu_set_error("reorder first"); return -1; }
		*ptr = c->order; *count = c->n_matched;
	} else if (!strcmp(name, "out_order")) {
		if (!c->encoded) { harcgpu_set_error("encode first"); return -1; }
		*ptr = c->o_order; *count = c->esz.n_order;
	} else { harcgpu_set_error("unknown device result '%s'", name); return -1; }
	return 0;
}

int harcgpu_get_counters(harcgpu_ctx *c, harcgpu_counters *o)
{
	if (!c || !o) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	u64 v[8];
	CK(cudaMemcpy(v, c->counters, 64, cudaMemcpyDeviceToHost));
	o->steps = v[0]; o->probes = v[1]; o->key_hits = v[2]; o->compares = v[3]; o->claim_fails = v[4]; o->restarts = v[5]; o->harvested = v[6];
	return 0;
}

// ---- stage II ---------------------------------------------------------------------------------------------
int harcgpu_set_stream(harcgpu_ctx *c, const char *dna, const char *flag, const uint8_t *pos, const uint32_t *order, const char *rev, uint32_t n)
{
	if (!c || (n && (!dna || !flag || !pos || !order || !rev))) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	return s2_set_stream_host(c, dna, flag, pos, order, rev, n);
}

int harcgpu_load_pool(harcgpu_ctx *c, const char *s_ascii, const uint32_t *order_s, uint32_t n_s, const char *N_ascii, uint32_t n_N)
{
	if (!c || (n_N && !N_ascii)) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	return s2_load_pool(c, s_ascii, order_s, n_s, N_ascii, n_N);
}

int harcgpu_load_pool_ids(harcgpu_ctx *c, const uint32_t *singleton_ids, uint32_t n_s, const char *N_ascii, uint32_t n_N)
{
	if (!c || (n_N && !N_ascii) || (n_s && !singleton_ids)) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	return s2_load_pool_ids(c, singleton_ids, n_s, N_ascii, n_N);
}

int harcgpu_stage_nreads(harcgpu_ctx *c, const char *N_ascii, uint32_t n_N)
{
	if (!c || (n_N && !N_ascii)) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	if (ensure_copy_stream(c)) return -1;
	c->release(c->staged_N);
	c->staged_N = nullptr; c->staged_host = nullptr; c->staged_n = 0;
	if (!n_N) return 0;
	const size_t bytes = (size_t)n_N * (c->L + 1);
	if (c->alloc(&c->staged_N, bytes + 16)) return -1;
	// the block may still be in use by work queued on the compute stream: order the copy behind it
	CK(cudaEventRecord(c->ev_order, c->st));
	CK(cudaStreamWaitEvent(c->st_copy, c->ev_order, 0));
	CK(cudaMemcpyAsync(c->staged_N, N_ascii, bytes, cudaMemcpyHostToDevice, c->st_copy));
	CK(cudaEventRecord(c->ev_staged, c->st_copy));
	c->staged_host = N_ascii; c->staged_n = n_N;
	return 0;
}

int harcgpu_load_pool_device(harcgpu_ctx *c, const void *d_N_ascii, uint32_t n_N)
{
	if (!c || (n_N && !d_N_ascii)) { harcgpu_set_error("null argument"); return -1; }
	if (!c->reordered) { harcgpu_set_error("harcgpu_load_pool_device needs the singletons of harcgpu_reorder on this context"); return -1; }
	CK(cudaSetDevice(c->device));
	return s2_load_pool_dev(c, d_N_ascii, n_N);
}

int harcgpu_encode(harcgpu_ctx *c)
{
	if (!c) { harcgpu_set_error("null argument"); return -1; }
	CK(cudaSetDevice(c->device));
	if (!c->stream_set) {
		if (!c->reordered) { harcgpu_set_error("no reordered stream: call harcgpu_reorder or harcgpu_set_stream first"); return -1; }
		if (s2_set_stream_from_stage1(c)) return -1;
	}
	if (!c->pool_set && s2_load_pool(c, nullptr, nullptr, 0, nullptr, 0)) return -1;
	return s2_encode(c);
}

int harcgpu_get_encode_sizes(harcgpu_ctx *c, harcgpu_encode_sizes *s)
{
	if (!c || !s || !c->encoded) { harcgpu_set_error("encode first"); return -1; }
	*s = c->esz;
	return 0;
}

int harcgpu_get_set_sizes(harcgpu_ctx *c, int k, harcgpu_set_sizes *s)
{
	if (!c || !s || !c->encoded || k < 0 || k >= (int)c->sets.size()) { harcgpu_set_error("bad file set"); return -1; }
	const SetOut &o = c->sets[k];
	s->seq_bytes = o.seq_bytes; s->seq_tail = o.seq_ntail; s->pos_bytes = o.pos_bytes; s->noise_bytes = o.noise_bytes;
	s->noisepos_bytes = o.noisepos_bytes; s->rev_bytes = o.rev_bytes; s->rev_tail = o.rev_ntail;
	return 0;
}

int harcgpu_get_set(harcgpu_ctx *c, int k, uint8_t *seq, char *seq_tail, uint8_t *pos, char *noise, uint8_t *noisepos, uint8_t *rev, char *rev_tail)
{
	if (!c || !c->encoded || k < 0 || k >= (int)c->sets.size()) { harcgpu_set_error("bad file set"); return -1; }
	CK(cudaSetDevice(c->device));
	const SetOut &o = c->sets[k];
	if (seq && o.seq_bytes) CK(cudaMemcpyAsync(seq, o.seq, o.seq_bytes, cudaMemcpyDeviceToHost, c->st));
	if (pos && o.pos_bytes) CK(cudaMemcpyAsync(pos, o.pos, o.pos_bytes, cudaMemcpyDeviceToHost, c->st));
	if (noise && o.noise_bytes) CK(cudaMemcpyAsync(noise, o.noise, o.noise_bytes, cudaMemcpyDeviceToHost, c->st));
	if (noisepos && o.noisepos_bytes) CK(cudaMemcpyAsync(noisepos, o.noisepos, o.noisepos_bytes, cudaMemcpyDeviceToHost, c->st));
	if (rev && o.rev_bytes) CK(cudaMemcpyAsync(rev, o.rev, o.rev_bytes, cudaMemcpyDeviceToHost, c->st));
	if (seq_tail) memcpy(seq_tail, o.seq_tail, o.seq_ntail);
	if (rev_tail) memcpy(rev_tail, o.rev_tail, o.rev_ntail);
	CK(cudaStreamSynchronize(c->st));
	return 0;
}

int harcgpu_get_globals(harcgpu_ctx *c, uint32_t *order, uint32_t *order_N, uint8_t *singleton, char *singleton_tail, char *input_N)
{
	if (!c || !c->encoded) { harcgpu_set_error("encode first"); return -1; }
	CK(cudaSetDevice(c->device));
	if (order && c->esz.n_order) CK(cudaMemcpyAsync(order, c->o_order, 4 * (size_t)c->esz.n_order, cudaMemcpyDeviceToHost, c->st));
	if (order_N && c->esz.n_order_N) CK(cudaMemcpyAsync(order_N, c->o_order_N, 4 * (size_t)c->esz.n_order_N, cudaMemcpyDeviceToHost, c->st));
	if (singleton && c->esz.singleton_bytes) CK(cudaMemcpyAsync(singleton, c->o_single, c->esz.singleton_bytes, cudaMemcpyDeviceToHost, c->st));
	if (input_N && c->esz.input_N_bytes) CK(cudaMemcpyAsync(input_N, c->o_inputN, c->esz.input_N_bytes, cudaMemcpyDeviceToHost, c->st));
	if (singleton_tail) memcpy(singleton_tail, c->single_tail, c->esz.singleton_tail);
	CK(cudaStreamSynchronize(c->st));
	return 0;
}

// pack_order.cpp:20-77 (harc:111-112, the -p mode) on the order stream of the last harcgpu_encode
int harcgpu_get_packed_order(harcgpu_ctx *c, void *packed, uint32_t *tail, uint64_t *packed_bytes, uint32_t *tail_entries)
{
	if (!c || !c->encoded) { harcgpu_set_error("encode first"); return -1; }
	CK(cudaSetDevice(c->device));
	u64 pb = 0; u32 nt = 0;
	int rc = s2_pack_order(c, packed, tail, &pb, &nt);
	if (packed_bytes) *packed_bytes = pb;
	if (tail_entries) *tail_entries = nt;
	return rc;
}

// ---- the process contract (reorder.out / encoder.out) -------------------------------------------------------
static bool slurp(const std::string &path, std::vector<char> &out)
{
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) return false;
	struct stat sb;
	if (fstat(fileno(f), &sb)) { fclose(f); return false; }
	out.resize((size_t)sb.st_size);
	size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
	fclose(f);
	return got == out.size();
}
static bool spit(const std::string &path, const void *p, size_t n)
{
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) return false;
	size_t w = n ? fwrite(p, 1, n, f) : 0;
	fclose(f);
	return w == n;
}
#define FAIL(...) do { harcgpu_set_error(__VA_ARGS__); return -1; } while (0)

// reorder.cpp:100-131
int harcgpu_reorder_dir(harcgpu_ctx *c, const char *basedir)
{
	if (!c || !basedir) FAIL("null argument");
	std::string out = std::string(basedir) + "/output/";
	std::vector<char> nb, dna;
	if (!slurp(out + "numreads.bin", nb) || nb.size() < 4) FAIL("cannot read %snumreads.bin", out.c_str());
	u32 n;
	memcpy(&n, nb.data(), 4);
	if (!slurp(out + "input_clean.dna", dna) && n) FAIL("cannot read %sinput_clean.dna", out.c_str());
	if (dna.size() < (size_t)n * (c->L + 1)) FAIL("input_clean.dna holds fewer than %u lines of %d bases", n, c->L);
	printf("Reading file: %sinput_clean.dna\n", out.c_str());
	if (harcgpu_load_reads(c, dna.data(), n)) return -1;
	std::vector<char>().swap(dna);
	if (n > 0) {
		printf("Constructing dictionaries\n");
		if (harcgpu_build_dicts(c)) return -1;
	} else {
		if (harcgpu_build_dicts(c)) return -1;
	}
	printf("Reordering reads\n");
	if (harcgpu_reorder(c)) return -1;
	printf("Reordering done, %u were unmatched\n", c->n_unmatched);
	printf("Writing to file\n");
	size_t m = c->n_matched, s = c->n_single, line = (size_t)c->L + 1;
	std::vector<u32> order(m), order_s(s);
	std::vector<char> rev(m), flag(m), tdna(m * line), sdna(s * line);
	std::vector<u8> pos(m);
	if (harcgpu_get_reorder(c, order.data(), rev.data(), flag.data(), pos.data(), order_s.data())) return -1;
	if (harcgpu_get_reordered_reads(c, tdna.data(), sdna.data())) return -1;
	if (!spit(out + "temp.dna", tdna.data(), tdna.size()) || !spit(out + "temp.dna.singleton", sdna.data(), sdna.size()) ||
	    !spit(out + "read_rev.txt", rev.data(), m) || !spit(out + "tempflag.txt", flag.data(), m) ||
	    !spit(out + "temppos.txt", pos.data(), m) || !spit(out + "read_order.bin", order.data(), 4 * m) ||
	    !spit(out + "read_order.bin.singleton", order_s.data(), 4 * s))
		FAIL("cannot write stage I files under %s", out.c_str());
	printf("Done!\n");
	return 0;
}

// encoder.cpp:108-152
int harcgpu_encode_dir(harcgpu_ctx *c, const char *basedir)
{
	if (!c || !basedir) FAIL("null argument");
	std::string out = std::string(basedir) + "/output/";
	std::vector<char> order, dna, flag, pos, rc, sd, so, N;
	size_t line = (size_t)c->L + 1;
	if (!slurp(out + "read_order.bin", order)) FAIL("cannot read %sread_order.bin", out.c_str());
	slurp(out + "temp.dna", dna); slurp(out + "tempflag.txt", flag); slurp(out + "temppos.txt", pos); slurp(out + "read_rev.txt", rc);
	slurp(out + "temp.dna.singleton", sd); slurp(out + "read_order.bin.singleton", so); slurp(out + "input_N.dna", N);
	u32 m = (u32)(order.size() / 4), ns = (u32)(sd.size() / line), nN = (u32)(N.size() / line);
	if (dna.size() < m * line || flag.size() < m || pos.size() < m || rc.size() < m || so.size() < 4 * (size_t)ns)
		FAIL("stage I files under %s are inconsistent", out.c_str());
	printf("Read length: %d\nNumber of non-singleton reads: %u\nNumber of singleton reads: %u\nNumber of reads with N: %u\n", c->L, m, ns, nN);
	if (harcgpu_set_stream(c, dna.data(), flag.data(), (const u8 *)pos.data(), (const u32 *)order.data(), rc.data(), m)) return -1;
	if (harcgpu_load_pool(c, sd.data(), (const u32 *)so.data(), ns, N.data(), nN)) return -1;
	printf("Encoding reads\n");
	if (harcgpu_encode(c)) return -1;
	harcgpu_encode_sizes es;
	if (harcgpu_get_encode_sizes(c, &es)) return -1;
	for (int k = 0; k < (int)c->sets.size(); k++) {
		harcgpu_set_sizes z;
		if (harcgpu_get_set_sizes(c, k, &z)) return -1;
		std::vector<u8> seq(z.seq_bytes), ps(z.pos_bytes), np(z.noisepos_bytes), rv(z.rev_bytes);
		std::vector<char> noise(z.noise_bytes);
		char st[8], rt[8];
		if (harcgpu_get_set(c, k, seq.data(), st, ps.data(), noise.data(), np.data(), rv.data(), rt)) return -1;
		std::string sk = "." + std::to_string(k);
		if (!spit(out + "read_seq.txt" + sk, seq.data(), seq.size()) || !spit(out + "read_seq.txt" + sk + ".tail", st, z.seq_tail) ||
		    !spit(out + "read_pos.txt" + sk, ps.data(), ps.size()) || !spit(out + "read_noise.txt" + sk, noise.data(), noise.size()) ||
		    !spit(out + "read_noisepos.txt" + sk, np.data(), np.size()) || !spit(out + "read_rev.txt" + sk, rv.data(), rv.size()) ||
		    !spit(out + "read_rev.txt" + sk + ".tail", rt, z.rev_tail))
			FAIL("cannot write stage II files under %s", out.c_str());
	}
	std::vector<u32> oo(es.n_order), oN(es.n_order_N);
	std::vector<u8> sg(es.singleton_bytes);
	std::vector<char> iN(es.input_N_bytes);
	char stl[8];
	if (harcgpu_get_globals(c, oo.data(), oN.data(), sg.data(), stl, iN.data())) return -1;
	std::string meta = std::to_string(c->L) + "\n";
	if (!spit(out + "read_order.bin", oo.data(), 4 * oo.size()) || !spit(out + "read_order_N_pe.bin", oN.data(), 4 * oN.size()) ||
	    !spit(out + "read_singleton.txt", sg.data(), sg.size()) || !spit(out + "read_singleton.txt.tail", stl, es.singleton_tail) ||
	    !spit(out + "input_N.dna", iN.data(), iN.size()) || !spit(out + "read_meta.txt", meta.data(), meta.size()))
		FAIL("cannot write stage II files under %s", out.c_str());
	printf("Encoding done:\n%u singleton reads were aligned\n%u reads with N were aligned\n", es.aligned_singletons, es.aligned_N);
	return 0;
}

} // extern "C"
