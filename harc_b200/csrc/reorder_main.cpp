// Drop-in for the reference's `reorder.out <basedir>` (reorder.cpp:100-131, launched at harc:67).
// The read length and the other config.h macros (harc:52-63) are derived at run time from the first line of
// output/input_clean.dna, so nothing is recompiled per run.  Extra options: --walkers N (the reference's num_thr),
// --device D.  Exit code 0 on success, 1 on any error (harc runs under `set -e`).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "harcgpu.h"

static int first_line_len(const std::string &path)
{
	FILE *f = fopen(path.c_str(), "rb");
	if (!f) return -1;
	int n = 0, ch;
	while ((ch = fgetc(f)) != EOF && ch != '\n') n++;
	fclose(f);
	return n;
}

int main(int argc, char **argv)
{
	if (argc < 2) { fprintf(stderr, "usage: reorder.out <basedir> [--walkers N] [--device D] [--readlen L]\n"); return 1; }
	std::string base = argv[1];
	int walkers = 0, device = 0, L = -1;
	for (int i = 2; i + 1 < argc; i += 2) {
		if (!strcmp(argv[i], "--walkers")) walkers = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "--readlen")) L = atoi(argv[i + 1]);
	}
	if (L < 0) L = first_line_len(base + "/output/input_clean.dna");
	if (L <= 0) L = first_line_len(base + "/output/input_N.dna"); // no clean read at all
	if (L <= 0) { fprintf(stderr, "cannot determine the read length from %s/output/input_clean.dna\n", base.c_str()); return 1; }
	harcgpu_params p;
	harcgpu_ctx *ctx = nullptr;
	if (harcgpu_default_params(L, &p)) { fprintf(stderr, "%s\n", harcgpu_last_error()); return 1; }
	p.walkers = walkers;
	if (harcgpu_create(device, &p, &ctx) || harcgpu_reorder_dir(ctx, base.c_str())) {
		fprintf(stderr, "reorder.out: %s\n", harcgpu_last_error());
		return 1;
	}
	harcgpu_destroy(ctx);
	return 0;
}
