// Drop-in for the reference's `encoder.out <basedir>` (encoder.cpp:108-152, launched at harc:69).
// Options: --sets K (number of read_*.txt.<k> file sets, the reference's num_thr; default 1), --device D.
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
	if (argc < 2) { fprintf(stderr, "usage: encoder.out <basedir> [--sets K] [--device D] [--readlen L]\n"); return 1; }
	std::string base = argv[1];
	int sets = 1, device = 0, L = -1;
	for (int i = 2; i + 1 < argc; i += 2) {
		if (!strcmp(argv[i], "--sets")) sets = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "--device")) device = atoi(argv[i + 1]);
		else if (!strcmp(argv[i], "--readlen")) L = atoi(argv[i + 1]);
	}
	const char *names[] = { "temp.dna", "temp.dna.singleton", "input_N.dna", "input_clean.dna" };
	for (int i = 0; i < 4 && L <= 0; i++) L = first_line_len(base + "/output/" + names[i]);
	if (L <= 0) { fprintf(stderr, "cannot determine the read length under %s/output\n", base.c_str()); return 1; }
	harcgpu_params p;
	harcgpu_ctx *ctx = nullptr;
	if (harcgpu_default_params(L, &p)) { fprintf(stderr, "%s\n", harcgpu_last_error()); return 1; }
	p.file_sets = sets;
	if (harcgpu_create(device, &p, &ctx) || harcgpu_encode_dir(ctx, base.c_str())) {
		fprintf(stderr, "encoder.out: %s\n", harcgpu_last_error());
		return 1;
	}
	harcgpu_destroy(ctx);
	return 0;
}
