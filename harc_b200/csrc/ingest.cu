// Fused FASTQ ingest: preprocess.cpp:49-138 + reorder.cpp:240-263 in one pass over the FASTQ bytes in
// HBM.  The reference splits the file on the CPU into input_clean.dna / input_N.dna / read_order_N.bin / numreads.bin
// with getline, and stage I then re-reads input_clean.dna line by line into bitsets.  Here:
//
//   K1 nl_count_kernel     newlines per 4 KiB tile (16-byte loads, byte-wise SIMD compare)         -> scan = line number of a tile
//   K2 seq_offset_kernel   byte offset of the sequence line (line 4r+1) of every record r          preprocess.cpp:81-113 `switch(i)`
//   K3 classify_kernel     warp per record: fixed read length check (preprocess.cpp:92-97), has-N flag (98)   -> scan = output slots
//   K4 emit_kernel         warp per record: clean reads go straight into the packed 2-bit array of stage I (A0 G1 C2 T3,
//                          reorder.cpp:188-195), reads with N as ASCII lines + their record number (preprocess.cpp:100-102)
//
// Streaming and HBM-bound: the file is read twice (K1, K2) plus the sequence lines twice (K3, K4).
#include "ctx.h"

namespace {
constexpr int ING_TILE = 4096; // bytes per block of K1/K2: 256 threads x 16 bytes

__device__ __forceinline__ u32 nl_mask16(const uint4 v) // bit i set <=> byte i of the 16 is '\n'
{
	u32 m = 0;
	const u32 w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const u32 e = __vcmpeq4(w[k], 0x0a0a0a0au) & 0x01010101u; // 1 in the low bit of every matching byte
		m |= ((e | (e >> 7) | (e >> 14) | (e >> 21)) & 0xfu) << (4 * k);
	}
	return m;
}
__device__ __forceinline__ uint4 load16(const char *__restrict__ buf, u64 nbytes, u64 off) // zero-filled behind the end
{
	if (off + 16 <= nbytes) return __ldg(reinterpret_cast<const uint4 *>(buf + off));
	uint4 v = make_uint4(0, 0, 0, 0);
	unsigned char *b = reinterpret_cast<unsigned char *>(&v);
	for (int i = 0; i < 16; i++)
		if (off + i < nbytes) b[i] = (unsigned char)buf[off + i];
	return v;
}

__global__ void __launch_bounds__(256) nl_count_kernel(const char *__restrict__ buf, u64 nbytes, u64 *__restrict__ counts)
{
	__shared__ u32 wsum[8];
	const u64 off = (u64)blockIdx.x * ING_TILE + 16ull * threadIdx.x;
	u32 c = off < nbytes ? __popc(nl_mask16(load16(buf, nbytes, off))) : 0u;
	for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
	if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
	__syncthreads();
	if (threadIdx.x == 0) {
		u32 t = 0;
		for (int i = 0; i < 8; i++) t += wsum[i];
		counts[blockIdx.x] = t;
	}
}

// base[b] = newlines before tile b = number of the line that is open at the tile's first byte.  The newline that ends
// line k starts line k + 1; line 4r + 1 is the sequence line of record r.
__global__ void __launch_bounds__(256) seq_offset_kernel(const char *__restrict__ buf, u64 nbytes, const u64 *__restrict__ base,
                                                         u64 nrec, u64 *__restrict__ seq_off)
{
	__shared__ u32 wsum[8];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u64 off = (u64)blockIdx.x * ING_TILE + 16ull * threadIdx.x;
	const u32 m = off < nbytes ? nl_mask16(load16(buf, nbytes, off)) : 0u;
	const u32 c = __popc(m);
	u32 incl = c;
	for (int o = 1; o < 32; o <<= 1) {
		const u32 t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += t;
	}
	if (lane == 31) wsum[warp] = incl;
	__syncthreads();
	u32 before = incl - c;
	for (int w = 0; w < warp; w++) before += wsum[w];
	u64 line = base[blockIdx.x] + before; // number of the line that the first newline of this thread ends
	u32 mm = m;
	while (mm) {
		const int i = __ffs(mm) - 1;
		mm &= mm - 1;
		const u64 next = line + 1; // the line that starts behind this newline
		if ((next & 3ull) == 1ull) {
			const u64 r = next >> 2;
			if (r < nrec) seq_off[r] = off + i + 1;
		}
		line++;
	}
}

// err[0] = (record number << 20 | length found) of the first record (lowest number) whose sequence line is not L long.
// A warp works on ING_RPW consecutive records at a time and issues the byte loads of all of them before it looks at
// any (the kernel is otherwise bound by the latency of one short dependent chain per warp).
constexpr int ING_RPW = 4;  // records per warp and round
// T = 32-byte slices of a line (readlen + 1 <= 256 -> T <= 8): a template parameter so that no dead slice is issued
template <int T>
__global__ void __launch_bounds__(256) classify_kernel(const char *__restrict__ buf, u64 nbytes, const u64 *__restrict__ seq_off, u64 nrec,
                                                       int L, u32 *__restrict__ isN, unsigned long long *__restrict__ err)
{
	const u64 r0 = (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * ING_RPW;
	const int lane = threadIdx.x & 31;
	if (r0 >= nrec) return;
	u64 off[ING_RPW];
	unsigned char ch[ING_RPW][T];
#pragma unroll
	for (int k = 0; k < ING_RPW; k++) off[k] = r0 + k < nrec ? __ldg(&seq_off[r0 + k]) : nbytes;
#pragma unroll
	for (int k = 0; k < ING_RPW; k++)
#pragma unroll
		for (int t = 0; t < T; t++) {
			const u64 p = off[k] + lane + 32 * t;
			ch[k][t] = p < nbytes ? (unsigned char)__ldg(&buf[p]) : (unsigned char)'\n'; // the end of the file ends a line
		}
#pragma unroll
	for (int k = 0; k < ING_RPW; k++) {
		const u64 r = r0 + k;
		if (r >= nrec) break;
		bool hasN = false;
		int nlpos = 0x7fffffff; // first newline / end of file inside the first L + 1 bytes of the line
#pragma unroll
		for (int t = 0; t < T; t++) {
			const int i = lane + 32 * t;
			if (i <= L) {
				if (ch[k][t] == '\n') nlpos = min(nlpos, i);
				else if (i < L && ch[k][t] == 'N') hasN = true;
			}
		}
		for (int o = 16; o > 0; o >>= 1) nlpos = min(nlpos, __shfl_xor_sync(0xffffffffu, nlpos, o));
		const u32 anyN = __ballot_sync(0xffffffffu, hasN);
		if (lane == 0) {
			isN[r] = anyN ? 1u : 0u;
			if (nlpos != L) {
				// line shorter than L (nlpos < L) or longer (no newline up to byte L): report like preprocess.cpp:92-97
				unsigned long long len = 0;
				if (nlpos < L) len = (unsigned long long)nlpos;
				else { // measure the long line
					u64 p = off[k] + L;
					while (p < nbytes && buf[p] != '\n') p++;
					len = p - off[k];
				}
				atomicMin(&err[0], (r << 20) | (len & 0xfffffull)); // smallest record number wins
			}
		}
	}
}

// H = 128-base halves of a read (readlen <= 128 -> 1, else 2)
template <int H>
__global__ void __launch_bounds__(256) emit_kernel(const char *__restrict__ buf, u64 nbytes, const u64 *__restrict__ seq_off, u64 nrec, int L,
                                                   int NW, const u32 *__restrict__ isN, const u32 *__restrict__ exN, u64 *__restrict__ reads,
                                                   char *__restrict__ outN, u32 *__restrict__ orderN)
{
	const u64 r0 = (((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * ING_RPW;
	const int lane = threadIdx.x & 31;
	if (r0 >= nrec) return;
	// lane k packs output bytes k and k + 32 of a read (bases 4k .. 4k+3): up to 8 input bytes per record
	u64 off[ING_RPW];
	u32 nb[ING_RPW], fl[ING_RPW];
	unsigned char ch[ING_RPW][4 * H];
#pragma unroll
	for (int k = 0; k < ING_RPW; k++) {
		const bool ok = r0 + k < nrec;
		off[k] = ok ? __ldg(&seq_off[r0 + k]) : 0;
		nb[k] = ok ? __ldg(&exN[r0 + k]) : 0u; // reads with N before this record
		fl[k] = ok ? __ldg(&isN[r0 + k]) : 0u;
	}
#pragma unroll
	for (int k = 0; k < ING_RPW; k++)
#pragma unroll
		for (int t = 0; t < 4 * H; t++) {
			const int i = 4 * (lane + 32 * (t >> 2)) + (t & 3);
			ch[k][t] = (r0 + k < nrec && i < L) ? (unsigned char)__ldg(&buf[off[k] + i]) : (unsigned char)'A';
		}
#pragma unroll
	for (int k = 0; k < ING_RPW; k++) {
		const u64 r = r0 + k;
		if (r >= nrec) break;
		if (fl[k]) {
			char *dst = outN + (size_t)nb[k] * (L + 1);
#pragma unroll
			for (int t = 0; t < 4 * H; t++) {
				const int i = 4 * (lane + 32 * (t >> 2)) + (t & 3);
				if (i < L) dst[i] = (char)ch[k][t];
			}
			if (lane == 0) { dst[L] = '\n'; orderN[nb[k]] = (u32)r; }
		} else {
			unsigned char *dst = reinterpret_cast<unsigned char *>(reads + (size_t)(r - nb[k]) * NW);
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const int kb = lane + 32 * h; // output byte
				if (h < H && kb < 8 * NW) {
					u32 v = 0;
#pragma unroll
					for (int t = 0; t < 4; t++) {
						// (ch >> 1) & 3 is A0 C1 T2 G3; the reference's code (reorder.cpp:188-195) is A0 G1 C2 T3; bases
						// behind the end of the read were loaded as 'A' = 0
						const u32 c = ch[k][4 * h + t];
						const u32 b0 = (c >> 1) & 1u, b1 = (c >> 2) & 1u;
						v |= (((b0 ^ b1) << 1) | b1) << (2 * t);
					}
					dst[kb] = (unsigned char)v;
				}
			}
		}
	}
}

__global__ void __launch_bounds__(256) iota_u32_kernel(u32 *v, u32 n)
{
	u32 i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) v[i] = i;
}
} // namespace

// d_fastq: the FASTQ bytes in device memory (16-byte aligned).  Fills c->reads (packed clean reads), c->ing_N (ASCII
// lines of the reads with N), c->ing_orderN; the caller (capi.cu) has reset stage I state.  Two host syncs (line
// count, clean/N split) because the output sizes depend on the data.
int ing_ingest(harcgpu_ctx *c, const char *d_fastq, u64 nbytes, u64 *total_reads, u32 *n_clean, u32 *n_N)
{
	cudaStream_t st = c->st;
	const int L = c->L;
	*total_reads = 0; *n_clean = 0; *n_N = 0;
	c->release(c->ing_N); c->release(c->ing_orderN);
	c->ing_N = nullptr; c->ing_orderN = nullptr; c->ing_nN = 0;
	c->release(c->reads); c->release(c->claim);
	c->reads = nullptr; c->claim = nullptr;
	c->n = 0;
	u64 nlines = 0, nrec = 0;
	u64 *counts = nullptr, *base = nullptr, *scan_tmp = nullptr, *d_tot = nullptr, *seq_off = nullptr;
	u32 *isN = nullptr, *exN = nullptr, *d_tot32 = nullptr;
	unsigned long long *err = nullptr;
	const size_t ntiles = (size_t)((nbytes + ING_TILE - 1) / ING_TILE);
	c->tic();
	if (nbytes) {
		if (c->alloc(&counts, ntiles) || c->alloc(&base, ntiles) || c->alloc(&scan_tmp, scan_tmp_elems(ntiles)) || c->alloc(&d_tot, 1)) return -1;
		nl_count_kernel<<<KL + (unsigned)ntiles, 256, 0, st>>>(d_fastq, nbytes, counts);
		CK(cudaGetLastError());
		if (exclusive_scan_u64(counts, base, ntiles, scan_tmp, d_tot, st)) return -1;
		u64 nl = 0;
		char last = 0;
		CK(cudaMemcpyAsync(&nl, d_tot, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&last, d_fastq + nbytes - 1, 1, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		nlines = nl + (last != '\n' ? 1 : 0); // what getline returns (preprocess.cpp:79)
		nrec = (nlines + 2) / 4;              // records that have a sequence line (line 4r + 1)
		*total_reads = nlines / 4;            // readnum: incremented at the fourth line (preprocess.cpp:120)
	}
	if (*total_reads > 4294967290ull || nrec > 4294967290ull) { // preprocess.cpp:124-128
		harcgpu_set_error("Too many reads. HARC supports at most 4294967290 reads");
		return -1;
	}
	u32 nN = 0;
	if (nrec) {
		if (c->alloc(&seq_off, nrec) || c->alloc(&isN, nrec) || c->alloc(&exN, nrec) || c->alloc(&d_tot32, 1) || c->alloc(&err, 1)) return -1;
		c->release(scan_tmp);
		scan_tmp = nullptr;
		if (c->alloc(&scan_tmp, scan_tmp_elems(nrec))) return -1;
		CK(cudaMemsetAsync(err, 0xff, 8, st));
		seq_offset_kernel<<<KL + (unsigned)ntiles, 256, 0, st>>>(d_fastq, nbytes, base, nrec, seq_off);
		const unsigned cg = cdiv((nrec + ING_RPW - 1) / ING_RPW * 32, 256);
		switch ((L + 1 + 31) / 32) {
		case 1: classify_kernel<1><<<KL + cg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, isN, err); break;
		case 2: classify_kernel<2><<<KL + cg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, isN, err); break;
		case 3: classify_kernel<3><<<KL + cg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, isN, err); break;
		case 4: classify_kernel<4><<<KL + cg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, isN, err); break;
		case 5: classify_kernel<5><<<KL + cg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, isN, err); break;
		case 6: classify_kernel<6><<<KL + cg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, isN, err); break;
		case 7: classify_kernel<7><<<KL + cg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, isN, err); break;
		default: classify_kernel<8><<<KL + cg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, isN, err); break;
		}
		CK(cudaGetLastError());
		if (exclusive_scan_u32(isN, exN, nrec, scan_tmp, d_tot32, st)) return -1;
		unsigned long long herr = 0;
		CK(cudaMemcpyAsync(&nN, d_tot32, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&herr, err, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		if (herr != ~0ull) {
			harcgpu_set_error("Read length not fixed. Found two different read lengths: %d and %llu (read %llu)", L, herr & 0xfffffull, herr >> 20);
			return -1;
		}
	}
	const u32 nclean = (u32)(nrec - nN);
	c->n = nclean;
	if (c->alloc(&c->reads, (size_t)nclean * c->NW) || c->alloc(&c->claim, ((size_t)nclean + 31) / 32)) return -1;
	if (c->alloc(&c->ing_N, (size_t)nN * (L + 1) + 16) || c->alloc(&c->ing_orderN, nN)) return -1;
	c->ing_nN = nN;
	if (nrec) {
		const unsigned eg = cdiv((nrec + ING_RPW - 1) / ING_RPW * 32, 256);
		if (L <= 128) emit_kernel<1><<<KL + eg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, c->NW, isN, exN, c->reads, c->ing_N, c->ing_orderN);
		else emit_kernel<2><<<KL + eg, 256, 0, st>>>(d_fastq, nbytes, seq_off, nrec, L, c->NW, isN, exN, c->reads, c->ing_N, c->ing_orderN);
		CK(cudaGetLastError());
	}
	c->toc("ingest");
	void *tmp[] = { counts, base, scan_tmp, d_tot, seq_off, isN, exN, d_tot32, err };
	for (void *q : tmp) c->release(q);
	*n_clean = nclean;
	*n_N = nN;
	return 0;
}

// input_clean.dna as the reference would have written it: the packed reads back as ASCII lines (device buffer)
int ing_unpack_clean(harcgpu_ctx *c, char *d_out)
{
	const u32 n = c->n;
	if (!n) return 0;
	u32 *order = nullptr;
	if (c->alloc(&order, n)) return -1;
	iota_u32_kernel<<<KL + cdiv(n, 256), 256, 0, c->st>>>(order, n);
	CK(cudaGetLastError());
	int rc = s1_unpack_reads(c, c->reads, order, nullptr, n, d_out);
	c->release(order);
	return rc;
}
