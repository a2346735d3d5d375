/* Synthetic workload generator for bench.py and the large-size tests (NOT part of the product, NOT the oracle).
 * Same read model as the reference's util/gen_fastq_noRC / util/gen_fastq (gen_fastq.cpp:97-130): uniform start
 * positions on a seeded i.i.d. ACGT genome, optional reverse complement of odd reads (covering ref[pos+1..pos+L]),
 * optional errors: each base mutates w.p. 1/100 to one of the three other bases or N, each w.p. 1/4.  It writes what
 * preprocess.cpp:98-109 would have written -- the lines of input_clean.dna and of input_N.dna -- straight into caller
 * buffers, skipping the FASTQ detour.  Seeded xorshift64*, so a (seed, sizes) pair names the workload exactly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

static inline uint64_t rng(uint64_t *s)
{
	uint64_t x = *s;
	x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
	*s = x;
	return x * 0x2545F4914F6CDD1DULL;
}
static inline double rngu(uint64_t *s) { return ((rng(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

void sim_genome(char *g, uint64_t n, uint64_t seed)
{
#pragma omp parallel
	{
		int t = omp_get_thread_num(), T = omp_get_num_threads();
		uint64_t s = seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL + (uint64_t)t * 0xD1B54A32D192ED03ULL;
		if (!s) s = 1;
		uint64_t a = n * t / T, b = n * (t + 1) / T;
		for (uint64_t i = a; i < b;) {
			uint64_t r = rng(&s);
			for (int k = 0; k < 32 && i < b; k++, i++) { g[i] = "ACGT"[r & 3]; r >>= 2; }
		}
	}
}

/* all: n lines of L chars + '\n'; hasN[i] = 1 if line i holds an N.  Returns number of lines with N. */
uint64_t sim_reads(const char *g, uint64_t glen, uint64_t n, int L, int rc, int errors, uint64_t seed, char *all, uint8_t *hasN)
{
	static const char trans[4][4] = { { 'G', 'C', 'T', 'N' }, { 'A', 'C', 'T', 'N' }, { 'A', 'G', 'T', 'N' }, { 'A', 'G', 'C', 'N' } };
	const double lg = log(0.99);
	uint64_t totalN = 0;
#pragma omp parallel reduction(+ : totalN)
	{
		int t = omp_get_thread_num(), T = omp_get_num_threads();
		uint64_t s = seed * 0xA0761D6478BD642FULL + 0x7654321ULL + (uint64_t)t * 0xE7037ED1A0B428DBULL;
		if (!s) s = 1;
		uint64_t a = n * t / T, b = n * (t + 1) / T;
		for (uint64_t i = a; i < b; i++) {
			char *o = all + i * (uint64_t)(L + 1);
			uint64_t pos = rng(&s) % (glen - L - 1);
			if (rc && (i & 1)) {
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

/* preprocess.cpp:98-109: split into clean lines and N lines (stable), read_order_N = index of each N read */
void sim_split(const char *all, const uint8_t *hasN, uint64_t n, int L, char *clean, char *withN, uint32_t *order_N)
{
	uint64_t line = (uint64_t)L + 1, c = 0, k = 0;
	for (uint64_t i = 0; i < n; i++) {
		if (hasN[i]) { memcpy(withN + k * line, all + i * line, line); order_N[k++] = (uint32_t)i; }
		else { memcpy(clean + c * line, all + i * line, line); c++; }
	}
}
