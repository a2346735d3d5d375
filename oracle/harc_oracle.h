/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of HARC's reorder + encode hot path.
 *
 * Plain C99, single thread, written from the behaviour of the reference (file:line cited per function in
 * harc_oracle.c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (libharcgpu.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks this restatement byte for byte against the
 * reference's own binaries (oracle/_ref, built from /root/reference by oracle/Makefile) run with num_thr=1
 * (the only setting at which the reference is deterministic), and against the fixtures under
 * tests/golden/ that oracle/make_golden.py generated from those binaries.
 */
#ifndef HARC_ORACLE_H
#define HARC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int readlen;        /* config.h readlen (harc:62) */
	int maxmatch;       /* readlen/2 (harc:52) */
	int thresh;         /* 4  (harc:53), in BITS of the 2-bit code */
	int thresh_s;       /* 24 (harc:54), in BITS of the 3-bit code */
	int numdict;        /* 2  (harc:55) */
	int maxsearch;      /* 1000 (harc:56) */
	int dict_start[2];  /* harc:57,59 */
	int dict_end[2];    /* harc:58,60 */
} oracle_params;

/* harc:52-63 */
void oracle_default_params(int readlen, oracle_params *p);

/* reorder.cpp:203-209 (stringtobitset): n lines of L chars + '\n' -> n x W little-endian u64 words, W=ceil(2L/64) */
void oracle_pack2(const char *ascii, uint32_t n, int L, uint64_t *out);
/* encoder.cpp:815-821 with the 3-bit code of encoder.cpp:731-745; W3=ceil(3L/64) */
void oracle_pack3(const char *ascii, uint32_t n, int L, uint64_t *out);

/* Canonical dictionary (reorder.cpp:277-394 / encoder.cpp:886-992 minus the MPHF): keys ascending, ids
 * ascending inside a bin.  bits = 2 or 3.  Caller frees *keys,*counts,*ids with oracle_free. */
int oracle_dict_canonical(const uint64_t *reads, uint32_t n, int words, int bits, int dstart, int dend,
                          uint64_t **keys, uint32_t **counts, uint32_t **ids, uint32_t *numkeys);
void oracle_free(void *p);

/* popcount(ref ^ (read & mask[j])) etc. exposed for unit tests (reorder.cpp:543,608) */
int oracle_hamming_fwd(const uint64_t *ref_shifted, const uint64_t *read, int L, int j);

/* reorder.cpp main(): reads <basedir>/output/{numreads.bin,input_clean.dna}; writes temp.dna, temp.dna.singleton,
 * read_rev.txt, tempflag.txt, temppos.txt, read_order.bin, read_order.bin.singleton.
 * walkers = the reference's num_thr; >1 is emulated deterministically (round-robin, one step per walker per turn).
 * Returns the number of unmatched reads (the count printed at reorder.cpp:701), or <0 on error. */
int64_t oracle_reorder_dir(const char *basedir, const oracle_params *p, int walkers);

/* encoder.cpp main(): consumes the stage I files + input_N.dna and writes the stage II files of SURVEY
 * Appendix A with K = nsets per-"thread" file sets.  aligned[0] = singletons aligned, aligned[1] = N reads aligned
 * (encoder.cpp:507-508).  Returns 0 or <0. */
int oracle_encode_dir(const char *basedir, const oracle_params *p, int nsets, uint32_t aligned[2]);

#ifdef __cplusplus
}
#endif
#endif
