/* Synthetic workload generator for bench.py and the large-size tests (NOT part of the product, NOT the oracle).
 * Same read model as the reference's util/gen_fastq_noRC / util/gen_fastq (gen_fastq.cpp:97-130): uniform start
 * positions on a seeded i.i.d. ACGT genome, optional reverse complement of odd reads (covering ref[pos+1..pos+L]),
 * optional errors: each base mutates w.p. 1/100 to one of the three other bases or N, each w.p. 1/4.  It writes what
 * preprocess.cpp:98-109 would have written -- the lines of input_clean.dna and of input_N.dna -- straight into caller
 * buffers, skipping the FASTQ detour.
 *
 * Reproducibility: the random stream is seeded per fixed-size BLOCK (SIM_GBLOCK genome bases, SIM_RBLOCK reads), never per
 * thread, so (seed, sizes) names the workload exactly on any machine and any range of blocks can be generated on its
 * own (one job on several GPUs: every rank makes only its slice of the reads).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define SIM_GBLOCK (1u << 20)
#define SIM_RBLOCK (1u << 14)

static inline uint64_t rng(uint64_t *s)
{
	uint64_t x = *s;
	x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
	*s = x;
	return x * 0x2545F4914F6CDD1DULL;
}
static inline double rngu(uint64_t *s) { return ((rng(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline uint64_t block_seed(uint64_t seed, uint64_t salt, uint64_t block)
{
	uint64_t z = seed * 0x9E3779B97F4A7C15ULL + salt + block * 0xD1B54A32D192ED03ULL; /* splitmix64 finaliser */
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	z ^= z >> 31;
	return z ? z : 1;
}

uint32_t sim_read_block(void) { return SIM_RBLOCK; }

void sim_genome(char *g, uint64_t n, uint64_t seed, int threads)
{
	const int64_t nb = (int64_t)((n + SIM_GBLOCK - 1) / SIM_GBLOCK);
	if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(static)
	for (int64_t blk = 0; blk < nb; blk++) {
		uint64_t s = block_seed(seed, 0x1234567ULL, (uint64_t)blk);
		uint64_t a = (uint64_t)blk * SIM_GBLOCK, b = a + SIM_GBLOCK < n ? a + SIM_GBLOCK : n;
		for (uint64_t i = a; i < b;) {
			uint64_t r = rng(&s);
			for (int k = 0; k < 32 && i < b; k++, i++) { g[i] = "ACGT"[r & 3]; r >>= 2; }
		}
	}
}

/* Reads [first, first + count) of the workload (first must be a multiple of SIM_RBLOCK): out = count lines of L chars +
 * '\n'; hasN[i] = 1 if line first + i holds an N.  Returns the number of lines with N. */
uint64_t sim_reads(const char *g, uint64_t glen, uint64_t first, uint64_t count, int L, int rc, int errors, uint64_t seed, char *out,
                   uint8_t *hasN, int threads)
{
	static const char trans[4][4] = { { 'G', 'C', 'T', 'N' }, { 'A', 'C', 'T', 'N' }, { 'A', 'G', 'T', 'N' }, { 'A', 'G', 'C', 'N' } };
	const double lg = log(0.99);
	uint64_t totalN = 0;
	const int64_t nb = (int64_t)((count + SIM_RBLOCK - 1) / SIM_RBLOCK);
	if (first % SIM_RBLOCK) return (uint64_t)-1;
	if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(static) reduction(+ : totalN)
	for (int64_t blk = 0; blk < nb; blk++) {
		uint64_t s = block_seed(seed, 0x7654321ULL, first / SIM_RBLOCK + (uint64_t)blk);
		uint64_t a = (uint64_t)blk * SIM_RBLOCK, b = a + SIM_RBLOCK < count ? a + SIM_RBLOCK : count;
		for (uint64_t i = a; i < b; i++) {
			char *o = out + i * (uint64_t)(L + 1);
			uint64_t pos = rng(&s) % (glen - L - 1);
			if (rc && ((first + i) & 1)) {
				for (int k = 0; k < L; k++) {
					char c = g[pos + L - k];
					o[k] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
				}
			} else
				memcpy(o, g + pos, L);
			o[L] = '\n';
			int n_flag = 0;
			if (errors) {
				/* geometric gaps between mutated bases: P(gap = k) = 0.99^k * 0.01 */
				int k = (int)(log(rngu(&s)) / lg);
				while (k < L) {
					int sym = o[k] == 'A' ? 0 : o[k] == 'G' ? 1 : o[k] == 'C' ? 2 : 3;
					char c = trans[sym][rng(&s) & 3];
					o[k] = c;
					if (c == 'N') n_flag = 1;
					k += 1 + (int)(log(rngu(&s)) / lg);
				}
			}
			hasN[i] = (uint8_t)n_flag;
			totalN += n_flag;
		}
	}
	return totalN;
}

/* preprocess.cpp:98-109: split into clean lines and N lines (stable), read_order_N = index of each N read (+ first) */
void sim_split(const char *all, const uint8_t *hasN, uint64_t n, int L, char *clean, char *withN, uint32_t *order_N, uint64_t first)
{
	uint64_t line = (uint64_t)L + 1, c = 0, k = 0;
	for (uint64_t i = 0; i < n; i++) {
		if (hasN[i]) { memcpy(withN + k * line, all + i * line, line); order_N[k++] = (uint32_t)(first + i); }
		else { memcpy(clean + c * line, all + i * line, line); c++; }
	}
}
