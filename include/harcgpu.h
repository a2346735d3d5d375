/* harcgpu.h -- C ABI of libharcgpu.so: HARC's stage I (hash-based read reordering) and stage II (consensus
 * encoding) as sm_100a CUDA kernels.
 *
 * The reference has no FFI for this path: its boundary is the process + file contract `reorder.out <basedir>` /
 * `encoder.out <basedir>` driven by harc:65-69, with the eleven macros of the generated src/config.h
 * (harc:52-63) as parameters.  This header is what a binding for that path would bind: every entry point cites
 * the reference code it replaces (file:line under shubhamchandak94/HARC).  The two drop-in executables
 * (harc_b200/csrc/reorder_main.cpp, encoder_main.cpp) are thin wrappers over these calls.
 *
 * Conventions: every function returns 0 on success and <0 on error (harcgpu_last_error() gives the text); no
 * exceptions cross the boundary; the caller owns every host buffer; the library owns all device memory inside the
 * opaque context.  There is no CPU fallback: without a CUDA device every compute entry point fails.
 * All multi-byte values are little endian, as in the reference's files.
 */
#ifndef HARCGPU_H
#define HARCGPU_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct harcgpu_ctx harcgpu_ctx;

/* The eleven macros of src/config.h (harc:52-63) as run-time values, plus the two knobs that are num_thr there. */
typedef struct {
	int readlen;       /* harc:62  readlen (<= 255) */
	int maxmatch;      /* harc:52  readlen/2 */
	int thresh;        /* harc:53  4, Hamming threshold in BITS of the 2-bit code (reorder.cpp:543) */
	int thresh_s;      /* harc:54  24, same for the 3-bit code of stage II (encoder.cpp:296) */
	int numdict;       /* harc:55  2 */
	int maxsearch;     /* harc:56  1000 */
	int dict_start[2]; /* harc:57,59 */
	int dict_end[2];   /* harc:58,60 */
	int walkers;       /* concurrent chain walkers = the reference's num_thr in reorder.cpp:455; 0 = choose for the GPU */
	int file_sets;     /* K, number of read_*.txt.<k> output sets = the reference's num_thr in encoder.cpp:169-196; 0 -> 1 */
	int reads_per_walker; /* walkers == 0: one walker per this many reads, capped at what the GPU keeps resident; 0 -> 4096 */
	int extend;        /* left extension of new chains (not in the reference; same file format): 1 on, -1 off,
	                      0 = on when more than one walker runs (one walker without it = the reference at num_thr=1) */
	int lanes_per_walker; /* GPU lanes that cooperate on one walker: 16 or 32; 0 = default (32, one warp) */
	int shard_dicts;   /* one job on several GPUs: 1 = both dictionaries sharded by key hash over the GPUs (all-to-all of the
	                      (key, id) pairs, probes through NVLink peer memory, 1/world of the tables per GPU), 0 = every
	                      GPU builds and holds both dictionaries */
} harcgpu_params;

/* Sizes of everything stage II produced, so the caller can allocate before harcgpu_get_*.  (encoder.cpp:457-508) */
typedef struct {
	uint32_t n_order;        /* entries of the rewritten read_order.bin */
	uint32_t n_order_N;      /* entries of read_order_N_pe.bin */
	uint64_t singleton_bytes, singleton_tail; /* read_singleton.txt, read_singleton.txt.tail */
	uint64_t input_N_bytes;  /* rewritten input_N.dna */
	uint32_t aligned_singletons, aligned_N;   /* the two counters printed at encoder.cpp:507-508 */
} harcgpu_encode_sizes;

typedef struct {
	uint64_t seq_bytes, seq_tail;     /* read_seq.txt.k, .tail        (encoder.cpp:519-551) */
	uint64_t pos_bytes;               /* read_pos.txt.k               (encoder.cpp:661-663,682-683,707-708) */
	uint64_t noise_bytes;             /* read_noise.txt.k             (encoder.cpp:673-681,698-706) */
	uint64_t noisepos_bytes;          /* read_noisepos.txt.k */
	uint64_t rev_bytes, rev_tail;     /* read_rev.txt.k, .tail        (encoder.cpp:554-580) */
} harcgpu_set_sizes;

/* Device-side counters of the last stage I run (SURVEY §5 "metrics"). */
typedef struct {
	uint64_t steps;        /* search rounds that start at shift 0 (one per appended read unless reads are harvested) */
	uint64_t probes;       /* dictionary probes issued */
	uint64_t key_hits;     /* probes whose key was present */
	uint64_t compares;     /* candidate reads fetched and Hamming-tested */
	uint64_t claim_fails;  /* lost CAS claims */
	uint64_t restarts;     /* chain heads picked (= the "unmatched" count of reorder.cpp:701) */
	uint64_t harvested;    /* reads appended from the candidates a round had already fetched (not in the reference) */
} harcgpu_counters;

/* harc:52-63: fill p from the read length exactly as the CLI does. */
int harcgpu_default_params(int readlen, harcgpu_params *p);
/* One context per GPU / process.  `device` is the CUDA ordinal. */
int harcgpu_create(int device, const harcgpu_params *p, harcgpu_ctx **out);
void harcgpu_destroy(harcgpu_ctx *ctx);
const char *harcgpu_last_error(void);
/* Number of kernels this library has launched in this process (every kernel on the path is the library's own). */
uint64_t harcgpu_launch_count(void);
int harcgpu_device_count(void);

/* ---- fused ingest (preprocess.cpp) --------------------------------------------------------- */
typedef struct {
	uint32_t readlen;      /* the context's read length (every sequence line was checked against it) */
	uint64_t total_reads;  /* "Total number of reads" (preprocess.cpp:134): complete four-line records */
	uint32_t n_clean;      /* numreads.bin: reads without N (preprocess.cpp:129-131) */
	uint32_t n_N;          /* lines of input_N.dna = entries of read_order_N.bin */
} harcgpu_ingest_info;
/* harc:44: length of the second line of the file (host side); <0 if there is none. */
int harcgpu_fastq_readlen(const char *fastq, uint64_t nbytes);
/* preprocess.cpp:49-138 without the intermediate files, fused with reorder.cpp:240-263: the FASTQ bytes (host memory)
 * are copied to the device, split into reads without / with N and the clean reads packed 2 bits/base in place, i.e. the
 * context is afterwards in the state harcgpu_load_reads leaves it in.  A sequence line whose length differs from the
 * context's readlen fails like preprocess.cpp:92-97.  Quality values and ids (-q) are not handled here. */
int harcgpu_ingest_fastq(harcgpu_ctx *ctx, const char *fastq, uint64_t nbytes, harcgpu_ingest_info *info);
/* Same, from a buffer already resident in device memory (16-byte aligned). */
int harcgpu_ingest_fastq_device(harcgpu_ctx *ctx, const void *d_fastq, uint64_t nbytes, harcgpu_ingest_info *info);
/* The three files preprocess would have written: input_clean.dna (n_clean lines), input_N.dna (n_N lines),
 * read_order_N.bin (n_N entries).  Any pointer may be NULL. */
int harcgpu_get_ingest(harcgpu_ctx *ctx, char *input_clean, char *input_N, uint32_t *order_N);
/* encoder.cpp:823-872 with the singletons of harcgpu_reorder and the reads with N of harcgpu_ingest_fastq, both
 * already on the device. */
int harcgpu_load_pool_ingested(harcgpu_ctx *ctx);

/* ---- stage I (reorder.cpp) ------------------------------------------------------------------------------ */
/* reorder.cpp:240-263 readDnaFile + 203-209 stringtobitset.  ascii = contents of input_clean.dna: n lines of
 * readlen chars in {A,C,G,T} + '\n' (host memory).  Copies to the device and packs 2 bits/base. */
int harcgpu_load_reads(harcgpu_ctx *ctx, const char *ascii, uint32_t n);
/* Same, from a buffer already resident in device memory, 16-byte aligned (bench: the kernel-only figure). */
int harcgpu_load_reads_device(harcgpu_ctx *ctx, const void *d_ascii, uint32_t n);
/* reorder.cpp:277-394 constructdictionary (both dictionaries). */
int harcgpu_build_dicts(harcgpu_ctx *ctx);
/* Test hook: canonical dump (keys ascending, ids ascending inside a bin) of dictionary l of stage (1|2).
 * With NULL buffers only *numkeys and *nids are written. */
int harcgpu_dump_dict(harcgpu_ctx *ctx, int stage, int l, uint64_t *keys, uint32_t *counts, uint32_t *ids,
                      uint32_t *numkeys, uint32_t *nids);
/* Test hook: the library's stable radix sort of (u64 key, u32 value) pairs on host arrays, in place.  mode 0: all 64
 * key bits, eight passes; 1: the dictionary build's route (four passes over the top half + fix-up of the runs that share
 * it, full sort only if a run is too long); 2: bits [begin_bit, end_bit) only. */
int harcgpu_debug_sort(harcgpu_ctx *ctx, uint64_t *keys, uint32_t *vals, uint64_t n, int mode, int begin_bit, int end_bit);
/* reorder.cpp:434-703 reorder() incl. 863-915 updaterefcount. */
int harcgpu_reorder(harcgpu_ctx *ctx);
int harcgpu_reorder_counts(harcgpu_ctx *ctx, uint32_t *n_matched, uint32_t *n_singleton, uint32_t *n_unmatched);
/* The five stage I streams: read_order.bin, read_rev.txt ('d'/'r'), tempflag.txt ('0'/'1'),
 * temppos.txt, read_order.bin.singleton.  Any pointer may be NULL. */
int harcgpu_get_reorder(harcgpu_ctx *ctx, uint32_t *order, char *rev, char *flag, uint8_t *pos, uint32_t *order_s);
/* reorder.cpp:722-830 writetofile: temp.dna (n_matched lines, reverse-complemented where flagged 'r') and
 * temp.dna.singleton (n_singleton lines).  Either pointer may be NULL. */
int harcgpu_get_reordered_reads(harcgpu_ctx *ctx, char *temp_dna, char *temp_dna_singleton);
int harcgpu_get_counters(harcgpu_ctx *ctx, harcgpu_counters *c);

/* ---- stage II (encoder.cpp) ----------------------------------------------------------------------------- */
/* encoder.cpp:185-225: the reordered stream as encoder.out reads it from temp.dna, tempflag.txt, temppos.txt,
 * read_order.bin, read_rev.txt (host buffers).  Not needed after harcgpu_reorder() on the same context: the
 * stream is then already resident on the device. */
int harcgpu_set_stream(harcgpu_ctx *ctx, const char *temp_dna, const char *flag, const uint8_t *pos,
                       const uint32_t *order, const char *rev, uint32_t n);
/* encoder.cpp:823-872 readsingletons: pool = singletons ++ reads with N.  singleton_ascii/order_s may be NULL
 * after harcgpu_reorder() on the same context (the singletons are then taken from the device). */
int harcgpu_load_pool(harcgpu_ctx *ctx, const char *singleton_ascii, const uint32_t *order_s, uint32_t n_s,
                      const char *N_ascii, uint32_t n_N);
/* Optional: start the upload of input_N.dna on a copy stream and return at once, so that it overlaps stage I.  A later
 * harcgpu_load_pool with the same N_ascii / n_N takes the uploaded copy.  N_ascii must stay valid (and should be
 * page-locked for the copy to be asynchronous) until then. */
int harcgpu_stage_nreads(harcgpu_ctx *ctx, const char *N_ascii, uint32_t n_N);
/* Same with the N reads already resident in device memory (16-byte aligned) and the singletons taken from
 * harcgpu_reorder() on this context (bench: the kernel-only figure). */
int harcgpu_load_pool_device(harcgpu_ctx *ctx, const void *d_N_ascii, uint32_t n_N);
/* encoder.cpp:154-510 encode() + 512-616 packbits() + 619-717 buildcontig/writecontig. */
int harcgpu_encode(harcgpu_ctx *ctx);
int harcgpu_get_encode_sizes(harcgpu_ctx *ctx, harcgpu_encode_sizes *s);
int harcgpu_get_set_sizes(harcgpu_ctx *ctx, int k, harcgpu_set_sizes *s);
/* Copy file set k to host buffers of the sizes reported above.  Any pointer may be NULL. */
int harcgpu_get_set(harcgpu_ctx *ctx, int k, uint8_t *seq, char *seq_tail, uint8_t *pos, char *noise,
                    uint8_t *noisepos, uint8_t *rev, char *rev_tail);
int harcgpu_get_globals(harcgpu_ctx *ctx, uint32_t *order, uint32_t *order_N, uint8_t *singleton,
                        char *singleton_tail, char *input_N);

/* pack_order.cpp:20-77 (run by harc:111-112 in the order-preserving mode -p) on the read_order.bin stream of the last
 * harcgpu_encode: `packed` receives the new read_order.bin (header {int numbits; uint32 numreads}, then numbits u32 words
 * per block of 32 entries), `tail` the read_order.bin.tail entries.  With NULL buffers only the sizes are written. */
int harcgpu_get_packed_order(harcgpu_ctx *ctx, void *packed, uint32_t *tail, uint64_t *packed_bytes, uint32_t *tail_entries);

/* ---- one job on several GPUs of one box (one process and one context per GPU) ----------------------------------
 * Not in the reference (it is one process); what is split is what its threads split: the key extraction of
 * reorder.cpp:284-302, the walker starts of reorder.cpp:476-497 and the contig ranges of encoder.cpp:169-180.
 *
 * Every GPU owns one allocation, its arena, which the other GPUs map (CUDA IPC between processes).  GPU r uploads and
 * packs its slice [base, base + n_local) of the clean reads; the pack kernel stores every packed read into the replica of
 * every GPU over NVLink.  The two dictionaries are sharded by key (params.shard_dicts = 1): GPU r extracts the keys of its
 * slice, the (key, id) pairs go to the owner of the key's hash range (an all-to-all done by plain stores into the owners'
 * receive buffers), the owner sorts them and builds its shard of the key table; every GPU also holds a Bloom filter over
 * all shards, so a walker asks a remote table only for keys that are almost certainly there.  With shard_dicts = 0 every
 * GPU builds both dictionaries over all reads instead.  The claimed-read bitmap (reorder.cpp:449 remainingreads + the
 * lock arrays of reorder.cpp:436-442) is cut into `world` contiguous id ranges, range r lives on GPU r and is read and
 * claimed (atomicAnd) by the others through NVLink peer memory.  GPU r's walkers start and restart only inside range r
 * but may claim any read, so the chains of all GPUs partition the read set exactly as the threads of the reference do.
 * Stage II then runs per GPU on its own chains (file set k = rank, encoder.cpp:169-196) against the same pool; which
 * contig gets a pool read is settled by an all-reduce(min) over the priority array, done by the caller's hook (NCCL
 * through torch.distributed in this repo), as is the all-gather of the singleton ids.  The GPUs synchronise through
 * barrier kernels inside these calls, so EVERY rank must make the same calls in the same order:
 *   job_init -> [exchange the 64-byte handles] -> job_connect -> per pass: job_load_reads -> job_build_dicts ->
 *   job_reorder -> [all-gather the singleton ids] -> load_pool_ids -> encode (calls the hook once) -> get_set(0) /
 *   get_globals. */
int harcgpu_job_init(harcgpu_ctx *ctx, int rank, int world, uint32_t n_total, uint32_t base, uint32_t n_local,
                     void *ipc_handle_out /* 64 bytes, may be NULL */, void **local_ptr_out /* may be NULL */);
/* handles: world x 64 bytes in rank order (other processes); or local_ptrs: the arenas of contexts of THIS process
 * (what job_init gave as local_ptr_out), which lets one GPU stand in for several in tests. */
int harcgpu_job_connect(harcgpu_ctx *ctx, const void *handles, void *const *local_ptrs);
/* Only for ranks that are contexts of ONE process sharing one GPU (tests): the ranks then meet through this host hook
 * (it must return 0 once every rank has called it) instead of the barrier kernel, because kernels of different streams
 * of one GPU are not guaranteed to run side by side.  NULL restores the barrier kernel. */
int harcgpu_job_set_barrier(harcgpu_ctx *ctx, int (*fn)(void *user), void *user);
/* reorder.cpp:240-263 for this rank's slice (host / 16-byte aligned device buffer of n_local lines). */
int harcgpu_job_load_reads(harcgpu_ctx *ctx, const char *ascii, uint32_t n_local);
int harcgpu_job_load_reads_device(harcgpu_ctx *ctx, const void *d_ascii, uint32_t n_local);
/* reorder.cpp:277-394 (harcgpu_build_dicts on a context of a job calls this). */
int harcgpu_job_build_dicts(harcgpu_ctx *ctx);
/* reorder.cpp:434-703 (harcgpu_reorder on a context of a job calls this). */
int harcgpu_job_reorder(harcgpu_ctx *ctx);
/* Hook called once per harcgpu_encode of a context of a job with the device array of `count` int64 priorities: it must
 * return 0 after replacing the array by its element-wise minimum over all ranks. */
int harcgpu_set_pool_exchange(harcgpu_ctx *ctx, int (*fn)(void *user, void *d_best, uint64_t count), void *user);
/* harcgpu_load_pool with the singletons given as ids into the reads of this context (the concatenation of all ranks'
 * read_order.bin.singleton; host or device memory); unaligned pool reads are written by rank 0 only. */
int harcgpu_load_pool_ids(harcgpu_ctx *ctx, const uint32_t *singleton_ids, uint32_t n_s, const char *N_ascii, uint32_t n_N);
/* Device pointer and element count of a result that the caller's plumbing moves between GPUs without a host copy:
 * "singleton_ids" (read_order.bin.singleton of the last reorder), "order" (read_order.bin of the last reorder),
 * "out_order" (read_order.bin of the last encode).  Valid until the next call that recomputes it. */
int harcgpu_device_result(harcgpu_ctx *ctx, const char *name, const void **ptr, uint64_t *count);

/* ---- the process contract, in-process --------------------------------------------------------------------- */
/* `reorder.out <basedir>` (reorder.cpp:100-131) and `encoder.out <basedir>` (encoder.cpp:108-152): read and write
 * the files of SURVEY Appendix A under <basedir>/output/. */
int harcgpu_reorder_dir(harcgpu_ctx *ctx, const char *basedir);
int harcgpu_encode_dir(harcgpu_ctx *ctx, const char *basedir);

/* Timing of the last call in milliseconds of device time (CUDA events on the context's stream): phases are
 * "pack","dict","walk","finalize","encode".  Returns <0 for an unknown phase. */
double harcgpu_last_ms(harcgpu_ctx *ctx, const char *phase);
/* The context keeps the device blocks it has released for the next pass (no driver allocation in steady state); this
 * gives them back to the driver, e.g. before the caller needs the memory for something else.  Results stay valid. */
int harcgpu_trim(harcgpu_ctx *ctx);
/* Raw CUDA stream of the context (cudaStream_t) so a caller can bracket calls with its own events. */
void *harcgpu_stream(harcgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
