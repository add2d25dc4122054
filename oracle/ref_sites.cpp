// ref_sites.cpp -- TEST INFRASTRUCTURE: runs the UNMODIFIED reference's .sites
// ingest (read_sites, find_compress_cols, compress_sites,
// make_sequences_from_sites; sequences.cpp:173-352,523-609, called as
// arg-sample.cpp:965-1007 calls them) and dumps what it produced, so that
// tests/test_sites_ingest.py can pin argweaver_b200's awb_sites_* against it.
// Linked with oracle/_ref/libargweaver.a by oracle/Makefile; copies no
// reference code.
//
//   ref_sites <in.sites> <out.awf> <compress> [region_start region_end]
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "argweaver/sequences.h"
#include "flatio.h"

using namespace argweaver;

int main(int argc, char **argv)
{
    if (argc < 4) {
        fprintf(stderr, "usage: ref_sites in.sites out.awf compress [start end]\n");
        return 2;
    }
    const int compress = atoi(argv[3]);
    const int rs = argc > 5 ? atoi(argv[4]) : -1;
    const int re = argc > 5 ? atoi(argv[5]) : -1;
    Sites sites;
    if (!read_sites(argv[1], &sites, rs, re))
        return 1;
    FILE *f = awf_create(argv[2]);
    const int nseqs = sites.get_num_seqs(), ncols = sites.get_num_sites();
    awf_write_int(f, "nseqs", nseqs);
    awf_write_int(f, "read_start", sites.start_coord);
    awf_write_int(f, "read_end", sites.end_coord);
    awf_write1(f, "read_positions", AWF_I32, ncols, ncols ? &sites.positions[0] : NULL);
    std::vector<unsigned char> cols((size_t) ncols * nseqs);
    for (int i = 0; i < ncols; i++)
        for (int j = 0; j < nseqs; j++)
            cols[(size_t) i * nseqs + j] = sites.cols[i][j];
    awf_write2(f, "cols", AWF_U8, ncols, nseqs, cols.data());
    SitesMapping mapping;
    const bool ok = find_compress_cols(&sites, compress, &mapping);
    awf_write_int(f, "compress_ok", ok ? 1 : 0);
    if (ok) {
        compress_sites(&sites, &mapping);
        Sequences sequences;
        make_sequences_from_sites(&sites, &sequences);
        awf_write_int(f, "start", sites.start_coord);
        awf_write_int(f, "end", sites.end_coord);
        awf_write1(f, "positions", AWF_I32, ncols, ncols ? &sites.positions[0] : NULL);
        awf_write1(f, "all_sites", AWF_I32, mapping.all_sites.size(),
                   mapping.all_sites.size() ? &mapping.all_sites[0] : NULL);
        const int L = sequences.length();
        std::vector<unsigned char> dense((size_t) nseqs * L);
        for (int j = 0; j < nseqs; j++)
            for (int i = 0; i < L; i++)
                dense[(size_t) j * L + i] = sequences.seqs[j][i];
        awf_write2(f, "seqs", AWF_U8, nseqs, L, dense.data());
    }
    fclose(f);
    return 0;
}
