/* TEST INFRASTRUCTURE ONLY -- see harc_oracle.h.  Restatement of HARC's stage I (reorder.cpp) and stage II
 * (encoder.cpp) in plain, single-threaded C99.  It restates WHAT the reference computes with arrays and word
 * arithmetic; the MPHF (BooPHF.h) is replaced by binary search over the sorted unique keys, because MPHF
 * indices never reach any output.  All file:line citations are into /root/reference/src/. */
#define _FILE_OFFSET_BITS 64
#include "harc_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXW 12 /* words per read: ceil(3*256/64) */

/* ------------------------------------------------------------------ parameters (harc:52-63) */
void oracle_default_params(int L, oracle_params *p)
{
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
}

/* ------------------------------------------------------------------ bit helpers */
static int popc64(uint64_t x) { return __builtin_popcountll(x); }

/* bits [pos, pos+n) of a little-endian word array, n <= 64 (bitset & mask >> start, to_ullong) */
static uint64_t getbits(const uint64_t *w, int words, int pos, int n)
{
	int q = pos >> 6, r = pos & 63;
	uint64_t v = w[q] >> r;
	if (r && q + 1 < words) v |= w[q + 1] << (64 - r);
	if (n < 64) v &= (((uint64_t)1) << n) - 1;
	return v;
}

/* bitset >>= s / <<= s on `nbits` significant bits */
static void shr(uint64_t *w, int words, int s)
{
	for (int i = 0; i < words; i++) {
		uint64_t v = w[i] >> s;
		if (i + 1 < words) v |= w[i + 1] << (64 - s);
		w[i] = v;
	}
}
static void shl(uint64_t *w, int words, int s, int nbits)
{
	for (int i = words - 1; i >= 0; i--) {
		uint64_t v = w[i] << s;
		if (i > 0) v |= w[i - 1] >> (64 - s);
		w[i] = v;
	}
	if (nbits & 63) w[words - 1] &= (((uint64_t)1) << (nbits & 63)) - 1;
}

/* 2-bit code of reorder.cpp:188-195: value = bit(2i) + 2*bit(2i+1): A=0 G=1 C=2 T=3 */
static int code2(char c)
{
	switch (c) { case 'A': return 0; case 'G': return 1; case 'C': return 2; case 'T': return 3; }
	return 0;
}
static const char dec2[4] = { 'A', 'G', 'C', 'T' }; /* revinttochar, reorder.cpp:59 */
/* 3-bit code of encoder.cpp:731-745: N=1 (bit 3i), G=2 (bit 3i+1), C=4 (bit 3i+2), T=6, A=0 */
static int code3(char c)
{
	switch (c) { case 'A': return 0; case 'N': return 1; case 'G': return 2; case 'C': return 4; case 'T': return 6; }
	return 0;
}
static const char dec3[8] = { 'A', 'N', 'G', '#', 'C', '#', 'T', '#' }; /* encoder.cpp:73 */
static char comp(char c) /* chartorevchar, reorder.cpp:135-138, encoder.cpp:721-725 */
{
	switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; }
	return 'N';
}
static void revcomp(const char *s, char *o, int L) /* reorder.cpp:856-861 */
{
	for (int j = 0; j < L; j++) o[j] = comp(s[L - 1 - j]);
}

static void enc_str(const char *s, int L, int bits, uint64_t *out, int words)
{
	memset(out, 0, 8 * words);
	for (int i = 0; i < L; i++) {
		uint64_t v = bits == 2 ? code2(s[i]) : code3(s[i]);
		int pos = bits * i;
		out[pos >> 6] |= v << (pos & 63);
		if ((pos & 63) + bits > 64) out[(pos >> 6) + 1] |= v >> (64 - (pos & 63));
	}
}
static void dec_str(const uint64_t *w, int words, int L, int bits, char *s)
{
	for (int i = 0; i < L; i++) {
		int v = (int)getbits(w, words, bits * i, bits);
		s[i] = bits == 2 ? dec2[v] : dec3[v];
	}
}

void oracle_pack2(const char *a, uint32_t n, int L, uint64_t *out)
{
	int W = (2 * L + 63) / 64;
	for (uint32_t i = 0; i < n; i++) enc_str(a + (size_t)i * (L + 1), L, 2, out + (size_t)i * W, W);
}
void oracle_pack3(const char *a, uint32_t n, int L, uint64_t *out)
{
	int W = (3 * L + 63) / 64;
	for (uint32_t i = 0; i < n; i++) enc_str(a + (size_t)i * (L + 1), L, 3, out + (size_t)i * W, W);
}
void oracle_free(void *p) { free(p); }

int oracle_hamming_fwd(const uint64_t *ref, const uint64_t *read, int L, int j)
{
	int W = (2 * L + 63) / 64, nb = 2 * (L - j), c = 0;
	for (int i = 0; i < W; i++) {
		uint64_t m = nb >= 64 * (i + 1) ? ~(uint64_t)0 : (nb <= 64 * i ? 0 : ((((uint64_t)1) << (nb - 64 * i)) - 1));
		c += popc64(ref[i] ^ (read[i] & m));
	}
	return c;
}

/* ------------------------------------------------------------------ dictionary
 * reorder.cpp:277-394 / encoder.cpp:886-992: key_i = (read_i & mask) >> bits*start; bins hold ids ascending.
 * reorder.cpp:396-432 (findpos/remove): deletion compacts the bin, so the live ids are always the ascending
 * prefix ids[start .. start+live); the "last read stays + empty_bin" trick is the same as live == 0. */
typedef struct {
	uint32_t numkeys;
	uint64_t *keys;  /* ascending */
	uint32_t *start; /* numkeys+1 */
	uint32_t *live;  /* numkeys */
	uint32_t *ids;   /* n */
} odict;

typedef struct { uint64_t k; uint32_t id; } kid;
static int cmp_kid(const void *a, const void *b)
{
	const kid *x = a, *y = b;
	if (x->k != y->k) return x->k < y->k ? -1 : 1;
	return x->id < y->id ? -1 : (x->id > y->id);
}

static void dict_build(odict *d, const uint64_t *reads, uint32_t n, int words, int bits, int ds, int de)
{
	kid *a = malloc(sizeof(kid) * (n ? n : 1));
	for (uint32_t i = 0; i < n; i++) {
		a[i].k = getbits(reads + (size_t)i * words, words, bits * ds, bits * (de - ds + 1));
		a[i].id = i;
	}
	qsort(a, n, sizeof(kid), cmp_kid);
	uint32_t nk = 0;
	for (uint32_t i = 0; i < n; i++) if (i == 0 || a[i].k != a[i - 1].k) nk++;
	d->numkeys = nk;
	d->keys = malloc(8 * (nk ? nk : 1));
	d->start = malloc(4 * (nk + 1));
	d->live = malloc(4 * (nk ? nk : 1));
	d->ids = malloc(4 * (n ? n : 1));
	uint32_t b = 0;
	for (uint32_t i = 0; i < n; i++) {
		if (i == 0 || a[i].k != a[i - 1].k) { d->keys[b] = a[i].k; d->start[b] = i; b++; }
		d->ids[i] = a[i].id;
	}
	d->start[nk] = n;
	for (uint32_t i = 0; i < nk; i++) d->live[i] = d->start[i + 1] - d->start[i];
	free(a);
}
static void dict_free(odict *d) { free(d->keys); free(d->start); free(d->live); free(d->ids); }

static int64_t dict_find(const odict *d, uint64_t key) /* stands in for bphf->lookup + the ull==ull1 check */
{
	int64_t lo = 0, hi = (int64_t)d->numkeys - 1;
	while (lo <= hi) {
		int64_t m = (lo + hi) >> 1;
		if (d->keys[m] == key) return m;
		if (d->keys[m] < key) lo = m + 1; else hi = m - 1;
	}
	return -1;
}
static void dict_remove(odict *d, int64_t bin, uint32_t id) /* reorder.cpp:409-432 */
{
	uint32_t s = d->start[bin], l = d->live[bin];
	uint32_t lo = 0, hi = l;
	while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (d->ids[s + m] < id) lo = m + 1; else hi = m; }
	if (lo >= l || d->ids[s + lo] != id) return;
	memmove(d->ids + s + lo, d->ids + s + lo + 1, 4 * (size_t)(l - lo - 1));
	d->live[bin] = l - 1;
}

int oracle_dict_canonical(const uint64_t *reads, uint32_t n, int words, int bits, int ds, int de,
                          uint64_t **keys, uint32_t **counts, uint32_t **ids, uint32_t *numkeys)
{
	odict d;
	dict_build(&d, reads, n, words, bits, ds, de);
	*keys = d.keys; *ids = d.ids; *numkeys = d.numkeys;
	*counts = d.live;
	free(d.start);
	return 0;
}

/* ------------------------------------------------------------------ file helpers */
static char *slurp(const char *path, size_t *len)
{
	FILE *f = fopen(path, "rb");
	if (!f) { *len = 0; return NULL; }
	fseeko(f, 0, SEEK_END);
	size_t n = (size_t)ftello(f);
	fseeko(f, 0, SEEK_SET);
	char *b = malloc(n + 1);
	if (n && fread(b, 1, n, f) != n) { fclose(f); free(b); *len = 0; return NULL; }
	b[n] = 0;
	fclose(f);
	*len = n;
	return b;
}
static int spit(const char *dir, const char *name, const void *buf, size_t len)
{
	char p[4096];
	snprintf(p, sizeof p, "%s/output/%s", dir, name);
	FILE *f = fopen(p, "wb");
	if (!f) return -1;
	if (len) fwrite(buf, 1, len, f);
	fclose(f);
	return 0;
}
typedef struct { char *p; size_t n, cap; } buf;
static void bput(buf *b, const void *s, size_t n)
{
	if (b->n + n > b->cap) { b->cap = (b->n + n) * 2 + 64; b->p = realloc(b->p, b->cap); }
	memcpy(b->p + b->n, s, n);
	b->n += n;
}
static void bputc(buf *b, char c) { bput(b, &c, 1); }
static void bput32(buf *b, uint32_t v) { bput(b, &v, 4); }

/* ------------------------------------------------------------------ stage I */
typedef struct {
	uint64_t ref[MAXW], revref[MAXW];
	int *count; /* [4][L], A C G T (chartoint, reorder.cpp:139-142) */
	uint32_t current, prev;
	int prev_unmatched, done;
	int64_t remainingpos;
	buf rc, flag, pos, order, order_s;
	uint32_t unmatched;
} walker;

static int c2i(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }
static const char i2c[4] = { 'A', 'C', 'G', 'T' };

/* reorder.cpp:863-915 */
static void updaterefcount(const uint64_t *cur, walker *w, int L, int W, int reset, int rev, int shift)
{
	char s[260], s1[260], *current;
	dec_str(cur, W, L, 2, s);
	if (!rev) current = s; else { revcomp(s, s1, L); current = s1; }
	int *cnt = w->count;
	if (reset) {
		memset(cnt, 0, sizeof(int) * 4 * L);
		for (int i = 0; i < L; i++) cnt[c2i(current[i]) * L + i] = 1;
	} else {
		for (int i = 0; i < L - shift; i++) {
			for (int j = 0; j < 4; j++) cnt[j * L + i] = cnt[j * L + i + shift];
			cnt[c2i(current[i]) * L + i] += 1;
			int max = 0, ind = 0;
			for (int j = 0; j < 4; j++) if (cnt[j * L + i] > max) { max = cnt[j * L + i]; ind = j; }
			current[i] = i2c[ind];
		}
		for (int i = L - shift; i < L; i++) {
			for (int j = 0; j < 4; j++) cnt[j * L + i] = 0;
			cnt[c2i(current[i]) * L + i] = 1;
		}
	}
	enc_str(current, L, 2, w->ref, W);
	char r[260];
	revcomp(current, r, L);
	enc_str(r, L, 2, w->revref, W);
}

typedef struct {
	const oracle_params *p;
	int L, W;
	uint32_t n;
	const uint64_t *read;
	odict dict[2];
	unsigned char *remaining;
} stage1;

/* one trip round the while(!done) loop of reorder.cpp:499-689 for one walker */
static void walker_step(stage1 *S, walker *w)
{
	const oracle_params *p = S->p;
	int L = S->L, W = S->W;
	/* 505-514: delete current from its bins */
	for (int l = 0; l < p->numdict; l++) {
		uint64_t key = getbits(S->read + (size_t)w->current * W, W, 2 * p->dict_start[l], 2 * (p->dict_end[l] - p->dict_start[l] + 1));
		int64_t b = dict_find(&S->dict[l], key);
		if (b >= 0) dict_remove(&S->dict[l], b, w->current);
	}
	int flag = 0;
	uint32_t k = 0;
	for (int j = 0; j < p->maxmatch && !flag; j++) {
		for (int rev = 0; rev < 2 && !flag; rev++) {       /* 520-580 forward, 585-643 reverse */
			const uint64_t *q = rev ? w->revref : w->ref;
			for (int l = 0; l < p->numdict && !flag; l++) {
				if (!rev && p->dict_end[l] + j >= L) continue;      /* 522 */
				if (rev && p->dict_start[l] <= j) continue;        /* 587 */
				uint64_t key = getbits(q, W, 2 * p->dict_start[l], 2 * (p->dict_end[l] - p->dict_start[l] + 1));
				int64_t b = dict_find(&S->dict[l], key);
				if (b < 0 || S->dict[l].live[b] == 0) continue;
				int64_t s = S->dict[l].start[b], e = s + S->dict[l].live[b];
				for (int64_t i = e - 1; i >= s && i >= e - p->maxsearch; i--) { /* 540, 605 */
					uint32_t rid = S->dict[l].ids[i];
					const uint64_t *r = S->read + (size_t)rid * W;
					int c = 0;
					for (int x = 0; x < W; x++) {
						/* mask[j]: low 2(L-j) bits; revmask[j]: bits >= 2j (706-718) */
						uint64_t m;
						int lo = rev ? 2 * j : 0, hi = rev ? 2 * L : 2 * (L - j);
						int a = lo - 64 * x, z = hi - 64 * x;
						if (z <= 0 || a >= 64) m = 0;
						else {
							m = ~(uint64_t)0;
							if (a > 0) m &= ~(uint64_t)0 << a;
							if (z < 64) m &= (((uint64_t)1) << z) - 1;
						}
						c += popc64(q[x] ^ (r[x] & m));
					}
					if (c <= p->thresh && S->remaining[rid]) { /* 543-552 */
						S->remaining[rid] = 0;
						k = rid;
						flag = 1;
						break;
					}
				}
				if (flag) {                                          /* 560-578, 624-641 */
					w->current = k;
					updaterefcount(S->read + (size_t)k * W, w, L, W, 0, rev, j);
					if (w->prev_unmatched) {
						bputc(&w->rc, 'd'); bput32(&w->order, w->prev); bputc(&w->flag, '0'); bputc(&w->pos, (char)L);
					}
					bputc(&w->rc, rev ? 'r' : 'd'); bput32(&w->order, k); bputc(&w->flag, '1'); bputc(&w->pos, (char)j);
					w->prev_unmatched = 0;
				}
			}
		}
		if (!flag) { shl(w->revref, W, 2, 2 * L); shr(w->ref, W, 2); } /* 647-648 */
	}
	if (!flag) {                                                       /* 650-688 */
		for (int64_t j = w->remainingpos; j >= 0; j--)
			if (S->remaining[j]) {
				w->current = (uint32_t)j; w->remainingpos = j - 1; S->remaining[j] = 0; flag = 1; w->unmatched++;
				break;
			}
		if (!flag) {
			if (w->prev_unmatched) bput32(&w->order_s, w->prev);
			w->done = 1;
		} else {
			updaterefcount(S->read + (size_t)w->current * W, w, L, W, 1, 0, 0);
			if (w->prev_unmatched) bput32(&w->order_s, w->prev);
			w->prev_unmatched = 1;
			w->prev = w->current;
		}
	}
}

int64_t oracle_reorder_dir(const char *basedir, const oracle_params *p, int T)
{
	char path[4096];
	size_t len;
	int L = p->readlen, W = (2 * L + 63) / 64;
	snprintf(path, sizeof path, "%s/output/numreads.bin", basedir);
	char *nb = slurp(path, &len);
	if (!nb || len < 4) return -1;
	uint32_t n;
	memcpy(&n, nb, 4);
	free(nb);
	snprintf(path, sizeof path, "%s/output/input_clean.dna", basedir);
	char *ascii = slurp(path, &len);
	if (n && (!ascii || len < (size_t)n * (L + 1))) return -2;
	uint64_t *read = calloc((size_t)(n ? n : 1) * W, 8);
	oracle_pack2(ascii, n, L, read);                                  /* 240-263 */
	stage1 S;
	S.p = p; S.L = L; S.W = W; S.n = n; S.read = read;
	for (int l = 0; l < p->numdict; l++) dict_build(&S.dict[l], read, n, W, 2, p->dict_start[l], p->dict_end[l]);
	S.remaining = malloc(n ? n : 1);
	memset(S.remaining, 1, n);
	walker *w = calloc(T, sizeof(walker));
	uint32_t firstread = 0;
	for (int t = 0; t < T; t++) {                                     /* 476-497 */
		w[t].count = calloc(4 * L, sizeof(int));
		w[t].remainingpos = (int64_t)n - 1;
		w[t].current = firstread;
		if (n == 0 || S.remaining[firstread] == 0) w[t].done = 1;
		else { S.remaining[firstread] = 0; w[t].unmatched++; }
		firstread += n / T;
		if (!w[t].done) {
			updaterefcount(read + (size_t)w[t].current * W, &w[t], L, W, 1, 0, 0);
			w[t].prev_unmatched = 1;
			w[t].prev = w[t].current;
		}
	}
	for (int alive = 1; alive;) {
		alive = 0;
		for (int t = 0; t < T; t++) if (!w[t].done) { walker_step(&S, &w[t]); alive = 1; }
	}
	/* writetofile, reorder.cpp:722-830: per-thread files concatenated in tid order */
	buf dna = {0}, dna_s = {0}, rc = {0}, flag = {0}, pos = {0}, order = {0}, order_s = {0};
	int64_t unmatched = 0;
	char s[260], s1[260];
	for (int t = 0; t < T; t++) {
		unmatched += w[t].unmatched;
		size_t m = w[t].rc.n;
		for (size_t i = 0; i < m; i++) {
			uint32_t id;
			memcpy(&id, w[t].order.p + 4 * i, 4);
			dec_str(read + (size_t)id * W, W, L, 2, s);
			if (w[t].rc.p[i] == 'd') bput(&dna, s, L); else { revcomp(s, s1, L); bput(&dna, s1, L); }
			bputc(&dna, '\n');
		}
		for (size_t i = 0; i < w[t].order_s.n / 4; i++) {
			uint32_t id;
			memcpy(&id, w[t].order_s.p + 4 * i, 4);
			dec_str(read + (size_t)id * W, W, L, 2, s);
			bput(&dna_s, s, L);
			bputc(&dna_s, '\n');
		}
		bput(&rc, w[t].rc.p, w[t].rc.n); bput(&flag, w[t].flag.p, w[t].flag.n); bput(&pos, w[t].pos.p, w[t].pos.n);
		bput(&order, w[t].order.p, w[t].order.n); bput(&order_s, w[t].order_s.p, w[t].order_s.n);
		free(w[t].rc.p); free(w[t].flag.p); free(w[t].pos.p); free(w[t].order.p); free(w[t].order_s.p); free(w[t].count);
	}
	spit(basedir, "temp.dna", dna.p, dna.n);
	spit(basedir, "temp.dna.singleton", dna_s.p, dna_s.n);
	spit(basedir, "read_rev.txt", rc.p, rc.n);
	spit(basedir, "tempflag.txt", flag.p, flag.n);
	spit(basedir, "temppos.txt", pos.p, pos.n);
	spit(basedir, "read_order.bin", order.p, order.n);
	spit(basedir, "read_order.bin.singleton", order_s.p, order_s.n);
	free(dna.p); free(dna_s.p); free(rc.p); free(flag.p); free(pos.p); free(order.p); free(order_s.p);
	for (int l = 0; l < p->numdict; l++) dict_free(&S.dict[l]);
	free(S.remaining); free(w); free(read); free(ascii);
	return unmatched;
}

/* ------------------------------------------------------------------ stage II */
static char enc_noise(char ref, char rd) /* encoder.cpp:752-771 */
{
	static const char *to = "ACGTN";
	static const char tab[5][5] = {
		/* ref A */ { 0, '0', '1', '2', '3' },
		/* ref C */ { '0', 0, '1', '2', '3' },
		/* ref G */ { '1', '2', 0, '0', '3' },
		/* ref T */ { '2', '1', '0', 0, '3' },
		/* ref N */ { '0', '2', '1', '3', 0 },
	};
	int a = (int)(strchr(to, ref) - to), b = (int)(strchr(to, rd) - to);
	return tab[a][b];
}

typedef struct { long pos; const char *read; char own[260]; uint32_t order; char rc; } centry;

typedef struct {
	buf seq, pos, noise, noisepos, rev, order, order_N;
} fileset;

/* encoder.cpp:619-652 */
static char *buildcontig(const centry *e, uint32_t m, int L, size_t *reflen)
{
	if (m == 1) {
		char *r = malloc(L + 1);
		memcpy(r, e[0].read, L);
		*reflen = L;
		return r;
	}
	size_t len = L;
	for (uint32_t i = 1; i < m; i++) len += e[i].pos;
	long (*cnt)[4] = calloc(len, sizeof(long[4]));
	size_t cur = 0;
	for (uint32_t i = 0; i < m; i++) {
		if (i) cur += e[i].pos;
		for (int x = 0; x < L; x++) cnt[cur + x][c2i(e[i].read[x])] += 1;
	}
	char *ref = malloc(len + 1);
	for (size_t i = 0; i < len; i++) {
		long max = 0, ind = 0;
		for (int j = 0; j < 4; j++) if (cnt[i][j] > max) { max = cnt[i][j]; ind = j; }
		ref[i] = i2c[ind];
	}
	free(cnt);
	*reflen = len;
	return ref;
}

/* encoder.cpp:654-717; e[i].pos are DELTAS here */
static void writecontig(const char *ref, size_t reflen, const centry *e, uint32_t m, int L, fileset *f)
{
	bput(&f->seq, ref, reflen);
	if (m == 1) {
		bputc(&f->noise, '\n'); bputc(&f->pos, (char)L); bput32(&f->order, e[0].order); bputc(&f->rev, e[0].rc);
		return;
	}
	long cur = 0;
	for (uint32_t i = 0; i < m; i++) {
		if (i) cur += e[i].pos;
		long prevj = 0;
		for (long j = 0; j < L; j++)
			if (e[i].read[j] != ref[cur + j]) {
				bputc(&f->noise, enc_noise(ref[cur + j], e[i].read[j]));
				bputc(&f->noisepos, (char)(j - prevj));
				prevj = j;
			}
		bputc(&f->noise, '\n');
		bputc(&f->pos, i ? (char)e[i].pos : (char)L);
		if (memchr(e[i].read, 'N', L)) bput32(&f->order_N, e[i].order); else bput32(&f->order, e[i].order);
		bputc(&f->rev, e[i].rc);
	}
}

typedef struct {
	const oracle_params *p;
	int L, W3;
	uint32_t npool, ns;
	uint64_t *pool; /* 3-bit packed singletons ++ N reads */
	uint32_t *order_s;
	unsigned char *remaining;
	odict dict[2];
	int ds[2], de[2];
} stage2;

static int ham3(const uint64_t *a, const uint64_t *b, int W3)
{
	int c = 0;
	for (int i = 0; i < W3; i++) c += popc64(a[i] ^ b[i]);
	return c;
}

/* encoder.cpp:231-418: slide the window along the contig and pull pool reads in.  On entry e[] holds
 * deltas; on exit a merged array (deltas again) is returned in *out (malloc'd). */
static uint32_t align_pool(stage2 *S, const char *ref, size_t reflen, centry *e, uint32_t m, centry **out)
{
	int L = S->L, W3 = S->W3;
	e[0].pos = 0;                                              /* 243 */
	long cum = 0;
	for (uint32_t i = 0; i < m; i++) { cum += e[i].pos; e[i].pos = cum; }   /* 246-251 */
	centry *ins = NULL;
	uint32_t nins = 0, capins = 0;
	uint64_t fw[MAXW], rv[MAXW];
	char tmp[260], tmp2[260];
	memcpy(tmp, ref, L);
	enc_str(tmp, L, 3, fw, W3);
	revcomp(tmp, tmp2, L);
	enc_str(tmp2, L, 3, rv, W3);
	uint32_t del[2048];
	for (size_t j = 0; j + L <= reflen; j++) {
		for (int rev = 0; rev < 2; rev++)
			for (int l = 0; l < 2; l++) {
				const uint64_t *q = rev ? rv : fw;
				uint64_t key = getbits(q, W3, 3 * S->ds[l], 3 * (S->de[l] - S->ds[l] + 1));
				int64_t b = dict_find(&S->dict[l], key);
				if (b < 0 || S->dict[l].live[b] == 0) continue;
				int64_t s = S->dict[l].start[b], en = s + S->dict[l].live[b];
				uint32_t ndel = 0;
				for (int64_t i = en - 1; i >= s && i >= en - S->p->maxsearch; i--) {   /* 293, 357: no break */
					uint32_t rid = S->dict[l].ids[i];
					if (ham3(q, S->pool + (size_t)rid * W3, W3) <= S->p->thresh_s && S->remaining[rid]) {
						S->remaining[rid] = 0;
						if (nins == capins) { capins = capins * 2 + 16; ins = realloc(ins, capins * sizeof(centry)); }
						centry *c = &ins[nins++];
						c->pos = (long)j;
						dec_str(S->pool + (size_t)rid * W3, W3, L, 3, tmp);
						if (rev) revcomp(tmp, c->own, L); else memcpy(c->own, tmp, L);
						c->read = NULL; /* fixed up after the realloc storm */
						c->order = S->order_s[rid];
						c->rc = rev ? 'r' : 'd';
						del[ndel++] = rid;
					}
				}
				for (uint32_t x = 0; x < ndel; x++)                                   /* 321-336 */
					for (int l1 = 0; l1 < 2; l1++) {
						const uint64_t *r = S->pool + (size_t)del[x] * W3;
						int64_t bb = dict_find(&S->dict[l1], getbits(r, W3, 3 * S->ds[l1], 3 * (S->de[l1] - S->ds[l1] + 1)));
						if (bb >= 0) dict_remove(&S->dict[l1], bb, del[x]);
					}
			}
		if (j + L < reflen) {                                     /* 402-408 */
			shr(fw, W3, 3);
			uint64_t v = code3(ref[j + L]);
			int pos = 3 * (L - 1);
			fw[pos >> 6] |= v << (pos & 63);
			if ((pos & 63) + 3 > 64) fw[(pos >> 6) + 1] |= v >> (64 - (pos & 63));
			shl(rv, W3, 3, 3 * L);
			rv[0] |= (uint64_t)code3(comp(ref[j + L]));
		}
	}
	/* list semantics of 254-268 + insert-before-pos_it: stable merge by position, originals first on ties */
	centry *r = malloc(sizeof(centry) * (m + nins));
	uint32_t a = 0, b = 0, o = 0;
	while (a < m || b < nins) {
		if (b >= nins || (a < m && e[a].pos <= ins[b].pos)) r[o++] = e[a++];
		else { r[o] = ins[b++]; o++; }
	}
	for (uint32_t i = 0; i < o; i++) if (!r[i].read) r[i].read = r[i].own;
	long prev = 0;
	for (uint32_t i = 0; i < o; i++) { long d = r[i].pos - prev; prev = r[i].pos; r[i].pos = d; }  /* 412-417 */
	free(ins);
	*out = r;
	return o;
}

/* encoder.cpp:512-616 */
static void pack_seq(const buf *in, buf *out, buf *tail)
{
	size_t n = in->n;
	for (size_t i = 0; i + 4 <= n; i += 4) {
		unsigned char b = (unsigned char)(64 * c2i(in->p[i + 3]) + 16 * c2i(in->p[i + 2]) + 4 * c2i(in->p[i + 1]) + c2i(in->p[i]));
		bputc(out, (char)b);
	}
	bput(tail, in->p + (n / 4) * 4, n % 4);
}
static void pack_rev(const buf *in, buf *out, buf *tail)
{
	size_t n = in->n;
	for (size_t i = 0; i + 8 <= n; i += 8) {
		unsigned char b = 0;
		for (int k = 0; k < 8; k++) b |= (unsigned char)((in->p[i + k] == 'r') << k);
		bputc(out, (char)b);
	}
	bput(tail, in->p + (n / 8) * 8, n % 8);
}

int oracle_encode_dir(const char *basedir, const oracle_params *p, int T, uint32_t aligned[2])
{
	char path[4096], name[256];
	int L = p->readlen, W3 = (3 * L + 63) / 64;
	size_t l_dna, l_flag, l_pos, l_order, l_rc, l_s, l_os, l_N;
#define IN(var, nm) snprintf(path, sizeof path, "%s/output/%s", basedir, nm); char *var = slurp(path, &l_##var)
	IN(order, "read_order.bin");
	IN(dna, "temp.dna");
	IN(flag, "tempflag.txt");
	IN(pos, "temppos.txt");
	IN(rc, "read_rev.txt");
	IN(s, "temp.dna.singleton");
	IN(os, "read_order.bin.singleton");
	IN(N, "input_N.dna");
#undef IN
	uint32_t numreads = (uint32_t)(l_order / 4);                      /* 781-813 */
	uint32_t ns = (uint32_t)(l_s / (L + 1)), nN = (uint32_t)(l_N / (L + 1));
	if (l_dna < (size_t)numreads * (L + 1) || l_flag < numreads || l_pos < numreads || l_rc < numreads) return -1;
	stage2 S;
	memset(&S, 0, sizeof S);
	S.p = p; S.L = L; S.W3 = W3; S.npool = ns + nN; S.ns = ns;
	S.pool = calloc((size_t)(S.npool ? S.npool : 1) * W3, 8);          /* 823-872 */
	oracle_pack3(s, ns, L, S.pool);
	oracle_pack3(N, nN, L, S.pool + (size_t)ns * W3);
	S.order_s = malloc(4 * (size_t)(S.npool ? S.npool : 1));
	for (uint32_t i = 0; i < ns; i++) memcpy(&S.order_s[i], os + 4 * (size_t)i, 4);
	for (uint32_t i = ns; i < S.npool; i++) S.order_s[i] = i - ns;
	if (L > 50) { S.ds[0] = 0; S.de[0] = 20; S.ds[1] = 21; S.de[1] = 41; }           /* 132-145 */
	else { S.ds[0] = 0; S.de[0] = 20 * L / 50; S.ds[1] = 20 * L / 50 + 1; S.de[1] = 41 * L / 50; }
	S.remaining = malloc(S.npool ? S.npool : 1);
	memset(S.remaining, 1, S.npool);
	if (S.npool) for (int l = 0; l < 2; l++) dict_build(&S.dict[l], S.pool, S.npool, W3, 3, S.ds[l], S.de[l]);

	uint32_t per = numreads ? 1 + (numreads - 1) / T : 0;              /* 169-180 */
	fileset *fs = calloc(T, sizeof(fileset));
	uint32_t start = 0;
	for (int t = 0; t < T; t++) {
		uint32_t a = start > numreads ? numreads : start;
		uint32_t stop = a + per > numreads ? numreads : a + per;
		start = stop;
		centry *e = NULL;
		uint32_t m = 0, cap = 0;
		for (uint32_t i = a; i <= stop; i++) {
			int last = (i == stop);
			if (last && a == stop) break;                              /* 438 */
			if (last || flag[i] == '0' || m > 10000000) {              /* 226 */
				if (m) {
					size_t reflen;
					char *ref = buildcontig(e, m, L, &reflen);
					if (!last && S.npool) {                            /* 231; the last contig of a range is not aligned (438-441) */
						centry *mg;
						uint32_t mm = align_pool(&S, ref, reflen, e, m, &mg);
						writecontig(ref, reflen, mg, mm, L, &fs[t]);
						free(mg);
					} else
						writecontig(ref, reflen, e, m, L, &fs[t]);
					free(ref);
				}
				m = 0;
			}
			if (last) break;
			if (m == cap) { cap = cap * 2 + 64; e = realloc(e, cap * sizeof(centry)); }
			e[m].pos = (unsigned char)pos[i];
			e[m].read = dna + (size_t)i * (L + 1);
			e[m].order = 0;
			memcpy(&e[m].order, order + 4 * (size_t)i, 4);
			e[m].rc = rc[i];
			m++;
		}
		free(e);
	}
	/* 457-503 */
	buf ord = {0}, ordN = {0}, sing = {0}, Nout = {0};
	for (int t = 0; t < T; t++) { bput(&ord, fs[t].order.p, fs[t].order.n); bput(&ordN, fs[t].order_N.p, fs[t].order_N.n); }
	uint32_t matched_s = ns, matched_N = nN;
	char tmp[260];
	for (uint32_t i = 0; i < ns; i++) if (S.remaining[i]) {
		matched_s--;
		bput32(&ord, S.order_s[i]);
		dec_str(S.pool + (size_t)i * W3, W3, L, 3, tmp);
		bput(&sing, tmp, L);
	}
	for (uint32_t i = ns; i < S.npool; i++) if (S.remaining[i]) {
		matched_N--;
		dec_str(S.pool + (size_t)i * W3, W3, L, 3, tmp);
		bput(&Nout, tmp, L);
		bputc(&Nout, '\n');
		bput32(&ordN, S.order_s[i]);
	}
	if (aligned) { aligned[0] = matched_s; aligned[1] = matched_N; }
	snprintf(tmp, sizeof tmp, "%d\n", L);
	spit(basedir, "read_meta.txt", tmp, strlen(tmp));
	spit(basedir, "read_order.bin", ord.p, ord.n);
	spit(basedir, "read_order_N_pe.bin", ordN.p, ordN.n);
	spit(basedir, "input_N.dna", Nout.p, Nout.n);
	buf pk = {0}, tl = {0};
	pack_seq(&sing, &pk, &tl);
	spit(basedir, "read_singleton.txt", pk.p, pk.n);
	spit(basedir, "read_singleton.txt.tail", tl.p, tl.n);
	for (int t = 0; t < T; t++) {
		buf a = {0}, b = {0};
		pack_seq(&fs[t].seq, &a, &b);
		snprintf(name, sizeof name, "read_seq.txt.%d", t); spit(basedir, name, a.p, a.n);
		snprintf(name, sizeof name, "read_seq.txt.%d.tail", t); spit(basedir, name, b.p, b.n);
		free(a.p); free(b.p);
		buf c = {0}, d = {0};
		pack_rev(&fs[t].rev, &c, &d);
		snprintf(name, sizeof name, "read_rev.txt.%d", t); spit(basedir, name, c.p, c.n);
		snprintf(name, sizeof name, "read_rev.txt.%d.tail", t); spit(basedir, name, d.p, d.n);
		free(c.p); free(d.p);
		snprintf(name, sizeof name, "read_pos.txt.%d", t); spit(basedir, name, fs[t].pos.p, fs[t].pos.n);
		snprintf(name, sizeof name, "read_noise.txt.%d", t); spit(basedir, name, fs[t].noise.p, fs[t].noise.n);
		snprintf(name, sizeof name, "read_noisepos.txt.%d", t); spit(basedir, name, fs[t].noisepos.p, fs[t].noisepos.n);
		free(fs[t].seq.p); free(fs[t].pos.p); free(fs[t].noise.p); free(fs[t].noisepos.p); free(fs[t].rev.p);
		free(fs[t].order.p); free(fs[t].order_N.p);
	}
	free(pk.p); free(tl.p); free(ord.p); free(ordN.p); free(sing.p); free(Nout.p); free(fs);
	if (S.npool) for (int l = 0; l < 2; l++) dict_free(&S.dict[l]);
	free(S.pool); free(S.order_s); free(S.remaining);
	free(order); free(dna); free(flag); free(pos); free(rc); free(s); free(os); free(N);
	return 0;
}
