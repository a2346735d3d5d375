// Test infrastructure only (oracle).  Include-wrapper around the reference's UNMODIFIED translation units:
// renames their main() and calls their own readDnaFile/readsingletons + constructdictionary, then dumps the
// dictionary in canonical form (bins sorted by key, ids in stored order) so a key-storing GPU table can be
// compared bit for bit.  MPHF index order never reaches any output of HARC, bin contents do.
//   stage I : reorder.cpp:240-263 (readDnaFile), 277-394 (constructdictionary)
//   stage II: encoder.cpp:823-872 (readsingletons), 886-992 (constructdictionary)     [-DSTAGE2]
// Dump format (little endian): u32 numdict; per dict: u32 numkeys, u32 nids; then numkeys x {u64 key, u32 count};
// then nids x u32 ids (bin after bin, in key order).
#define main ref_main
#include REF_TU
#undef main
#include <cstdio>
#include <utility>

int main(int argc, char **argv)
{
	if (argc < 3) { fprintf(stderr, "usage: dictdump <basedir> <out.bin>\n"); return 2; }
	std::string basedir = argv[1];
	outdir = basedir + "/output/";
	omp_set_num_threads(num_thr);
#ifdef STAGE2
	infile = basedir + "/output/temp.dna";
	infile_order = basedir + "/output/read_order.bin";
	infile_N = basedir + "/output/input_N.dna";
	getDataParams();
	setglobalarrays();
	uint32_t n = numreads_s + numreads_N;
	std::bitset<3*readlen> *read = new std::bitset<3*readlen>[n];
	uint32_t *order_s = new uint32_t[n];
	readsingletons(read, order_s);
	if (readlen > 50) { dict_start[0] = 0; dict_end[0] = 20; dict_start[1] = 21; dict_end[1] = 41; }
	else { dict_start[0] = 0; dict_end[0] = 20*readlen/50; dict_start[1] = 20*readlen/50 + 1; dict_end[1] = 41*readlen/50; }
	const int nd = numdict_s;
	bbhashdict dict[2];
	std::bitset<3*readlen> m[2];
	const int bits = 3;
#else
	infile = basedir + "/output/input_clean.dna";
	infilenumreads = basedir + "/output/numreads.bin";
	std::ifstream f_numreads(infilenumreads, std::ios::binary);
	f_numreads.read((char*)&numreads, sizeof(uint32_t));
	setglobalarrays();
	uint32_t n = numreads;
	std::bitset<2*readlen> *read = new std::bitset<2*readlen>[n];
	readDnaFile(read);
	const int nd = numdict;
	bbhashdict dict[numdict];
	std::bitset<2*readlen> m[numdict];
	const int bits = 2;
#endif
	if (n > 0) constructdictionary(read, dict);
	generateindexmasks(m);
	FILE *fo = fopen(argv[2], "wb");
	uint32_t u = nd; fwrite(&u, 4, 1, fo);
	for (int j = 0; j < nd; j++) {
		uint32_t nk = n ? dict[j].numkeys : 0;
		std::vector<std::pair<uint64_t, uint32_t>> bins(nk);
		for (uint32_t b = 0; b < nk; b++) {
			uint32_t rid = dict[j].read_id[dict[j].startpos[b]];
			bins[b] = std::make_pair(((read[rid] & m[j]) >> bits*dict_start[j]).to_ullong(), b);
		}
		std::sort(bins.begin(), bins.end());
		fwrite(&nk, 4, 1, fo); fwrite(&n, 4, 1, fo);
		for (uint32_t b = 0; b < nk; b++) {
			uint32_t c = dict[j].startpos[bins[b].second+1] - dict[j].startpos[bins[b].second];
			fwrite(&bins[b].first, 8, 1, fo); fwrite(&c, 4, 1, fo);
		}
		for (uint32_t b = 0; b < nk; b++) {
			uint32_t s = dict[j].startpos[bins[b].second], e = dict[j].startpos[bins[b].second+1];
			fwrite(dict[j].read_id + s, 4, e - s, fo);
		}
	}
	fclose(fo);
	return 0;
}
