// awb_sites.cpp -- .sites ingest and site compression (host side of SURVEY
// section 8f, N-2): the step right before the threading path.
//
// Replaces, with the same results,
//   read_sites                 sequences.cpp:173-303   (NAMES / REGION / pos<TAB>column)
//   validate_site_column       sequences.cpp:159-167
//   find_compress_cols         sequences.cpp:523-597   (-c / --compress-seq)
//   compress_sites             sequences.cpp:601-609
//   make_sequences_from_sites  sequences.cpp:323-352   (dense rows, default 'A')
// as called by arg-sample (arg-sample.cpp:965-1007).  The path itself does not
// need the dense rows: awb_problem takes the variant columns as they are
// (var_pos / var_cols), which is ~3 % of the bytes of the dense alignment.
//
// Plain C ABI (include/argweaver_b200.h); host-only code, compiled into
// libargweaver_b200.so.

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "argweaver_b200.h"

int awb_fail_msg(const std::string &msg);      // awb_api.cu

struct awb_sites {
    std::string chrom;
    int start_coord, end_coord;                // 0-based, half open
    std::vector<std::string> names;
    std::vector<int> positions;                // ascending, unique
    std::vector<unsigned char> cols;           // [ncols][nseqs]
    // SitesMapping (sequences.h:272): filled by awb_sites_compress
    bool compressed;
    int old_start, old_end, new_start, new_end;
    std::vector<int> old_sites, new_sites, all_sites;
};

static int base_code(unsigned char c)
{
    switch (c) {                               // seq.cpp:15-43 (dna2int)
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    }
    return -1;
}

extern "C" int awb_sites_read(const char *filename, int subregion_start,
                              int subregion_end, awb_sites **out)
{
    FILE *f = fopen(filename, "r");
    if (!f)
        return awb_fail_msg(std::string("cannot read file '") + filename + "'");
    awb_sites *s = new awb_sites;
    s->start_coord = s->end_coord = 0;
    s->compressed = false;
    int nseqs = 0, lineno = 0, rc = 0;
    std::string line;
    char *buf = NULL;
    size_t cap = 0;
    ssize_t len;
    while (!rc && (len = getline(&buf, &cap, f)) >= 0) {
        while (len > 0 && (buf[len - 1] == '\n' || buf[len - 1] == '\r'))
            buf[--len] = 0;
        lineno++;
        if (strncmp(buf, "NAMES\t", 6) == 0) {
            s->names.clear();
            const char *p = buf + 6;
            for (;;) {
                const char *q = strchr(p, '\t');
                s->names.push_back(q ? std::string(p, q - p) : std::string(p));
                if (!q) break;
                p = q + 1;
            }
            nseqs = (int) s->names.size();
            for (int i = 0; i < nseqs && !rc; i++)
                if (s->names[i].empty())
                    rc = awb_fail_msg("name for sequence " + std::to_string(i + 1) +
                                      " is zero length (line " + std::to_string(lineno) + ")");
        } else if (strncmp(buf, "REGION\t", 7) == 0) {
            char chrom[51];
            if (sscanf(buf, "REGION\t%50s\t%d\t%d", chrom, &s->start_coord,
                       &s->end_coord) != 3) {
                rc = awb_fail_msg("bad REGION format");
                break;
            }
            s->chrom = chrom;
            s->start_coord--;                  // 0-based
            if (subregion_start != -1) s->start_coord = subregion_start;
            if (subregion_end != -1) s->end_coord = subregion_end;
        } else if (strncmp(buf, "RANGE\t", 6) == 0) {
            rc = awb_fail_msg("deprecated RANGE line detected (use REGION instead)");
        } else {
            int position;
            if (sscanf(buf, "%d\t", &position) != 1) {
                rc = awb_fail_msg("first column is not an integer (line " +
                                  std::to_string(lineno) + ")");
                break;
            }
            // (the reference compares the 1-based position with the 0-based
            // region here, sequences.cpp:244-247; kept)
            if (position < s->start_coord || position >= s->end_coord)
                continue;
            const char *col = strchr(buf, '\t');
            col = col ? col + 1 : buf + len;
            if ((int) strlen(col) != nseqs) {
                rc = awb_fail_msg("the number bases given, " + std::to_string(strlen(col)) +
                                  ", does not match the number of sequences " +
                                  std::to_string(nseqs) + " (line " +
                                  std::to_string(lineno) + ")");
                break;
            }
            const size_t o = s->cols.size();
            s->cols.resize(o + nseqs);
            for (int i = 0; i < nseqs && !rc; i++) {
                const unsigned char c = (unsigned char) toupper((unsigned char) col[i]);
                if (c != 'N' && base_code(c) < 0)
                    rc = awb_fail_msg("invalid sequence characters (line " +
                                      std::to_string(lineno) + ")");
                s->cols[o + i] = c;
            }
            if (rc) break;
            position--;
            if (!s->positions.empty() && s->positions.back() >= position) {
                rc = awb_fail_msg("invalid site location " +
                                  std::to_string(s->positions.back()) + " >= " +
                                  std::to_string(position) + " (line " +
                                  std::to_string(lineno) + "): sites must be sorted and unique");
                break;
            }
            s->positions.push_back(position);
        }
    }
    free(buf);
    fclose(f);
    if (rc) {
        delete s;
        return rc;
    }
    *out = s;
    return 0;
}

extern "C" int awb_sites_from_columns(int nseqs, int start_coord, int end_coord,
                                      int ncols, const int *positions,
                                      const unsigned char *cols, awb_sites **out)
{
    if (nseqs < 1 || ncols < 0 || end_coord < start_coord)
        return awb_fail_msg("awb_sites_from_columns: bad dimensions");
    for (int i = 1; i < ncols; i++)
        if (positions[i] <= positions[i - 1])
            return awb_fail_msg("awb_sites_from_columns: sites must be sorted and unique");
    awb_sites *s = new awb_sites;
    s->start_coord = start_coord;
    s->end_coord = end_coord;
    s->compressed = false;
    s->names.resize(nseqs);
    for (int i = 0; i < nseqs; i++)
        s->names[i] = "n" + std::to_string(i);
    s->positions.assign(positions, positions + ncols);
    s->cols.assign(cols, cols + (size_t) ncols * nseqs);
    *out = s;
    return 0;
}

extern "C" void awb_sites_free(awb_sites *s) { delete s; }
extern "C" int awb_sites_nseqs(const awb_sites *s) { return (int) s->names.size(); }
extern "C" int awb_sites_ncols(const awb_sites *s) { return (int) s->positions.size(); }
extern "C" int awb_sites_start(const awb_sites *s) { return s->start_coord; }
extern "C" int awb_sites_end(const awb_sites *s) { return s->end_coord; }
extern "C" const char *awb_sites_name(const awb_sites *s, int i)
{
    return (i >= 0 && i < (int) s->names.size()) ? s->names[i].c_str() : "";
}
extern "C" const int *awb_sites_positions(const awb_sites *s) { return s->positions.data(); }
extern "C" const unsigned char *awb_sites_columns(const awb_sites *s) { return s->cols.data(); }

// find_compress_cols + compress_sites.  Returns 2 (and leaves the sites as they
// were) when the alignment cannot be compressed at this level.
extern "C" int awb_sites_compress(awb_sites *s, int compress)
{
    if (compress < 1)
        return awb_fail_msg("awb_sites_compress: compress must be >= 1");
    if (s->compressed)
        return awb_fail_msg("awb_sites_compress: already compressed");
    const int ncols = (int) s->positions.size();
    int blocki = 0;
    int next_block = s->start_coord + compress;
    const int half_block = compress / 2;
    std::vector<int> old_sites, new_sites, all_sites;
    int new_end;
    if (compress == 1) {
        for (int i = s->start_coord; i < s->end_coord; i++)
            all_sites.push_back(i);
        for (int i = 0; i < ncols; i++) {
            old_sites.push_back(s->positions[i]);
            new_sites.push_back(s->positions[i] - s->start_coord);
        }
        new_end = s->end_coord - s->start_coord;
    } else {
        for (int i = 0; i < ncols; i++) {
            const int col = s->positions[i];
            while (col >= next_block) {
                all_sites.push_back(next_block - half_block);
                next_block += compress;
                blocki++;
            }
            old_sites.push_back(col);
            new_sites.push_back(blocki);
            all_sites.push_back(col);
            next_block += compress;
            blocki++;
            if (next_block - compress > s->end_coord && i != ncols - 1) {
                awb_fail_msg("unable to compress sequences at given compression level");
                return 2;
            }
        }
        while (s->end_coord >= next_block) {
            all_sites.push_back(next_block - half_block);
            next_block += compress;
            blocki++;
        }
        new_end = (s->end_coord - s->start_coord) / compress;
        if (ncols > 0 && new_sites[ncols - 1] + 1 > new_end)
            new_end = new_sites[ncols - 1] + 1;
    }
    s->old_start = s->start_coord;
    s->old_end = s->end_coord;
    s->new_start = 0;
    s->new_end = new_end;
    s->old_sites.swap(old_sites);
    s->new_sites.swap(new_sites);
    s->all_sites.swap(all_sites);
    s->start_coord = s->new_start;
    s->end_coord = s->new_end;
    for (int i = 0; i < ncols; i++)
        s->positions[i] = s->new_sites[i];
    s->compressed = true;
    return 0;
}

extern "C" int awb_sites_mapping_size(const awb_sites *s)
{
    return s->compressed ? (int) s->all_sites.size() : -1;
}

extern "C" const int *awb_sites_mapping(const awb_sites *s)
{
    return s->compressed ? s->all_sites.data() : NULL;
}

// make_sequences_from_sites: dense rows seqs[nseqs][end - start]
extern "C" int awb_sites_to_sequences(const awb_sites *s, unsigned char *seqs,
                                      unsigned char default_char)
{
    const int nseqs = (int) s->names.size();
    const int seqlen = s->end_coord - s->start_coord;
    const int nsites = (int) s->positions.size();
    if (!default_char) default_char = 'A';
    for (int i = 0; i < nseqs; i++) {
        unsigned char *row = seqs + (size_t) i * seqlen;
        memset(row, default_char, seqlen);
    }
    for (int c = 0; c < nsites; c++) {
        const int j = s->positions[c] - s->start_coord;
        if (j < 0 || j >= seqlen)
            continue;                          // (outside the region: never matched)
        for (int i = 0; i < nseqs; i++)
            seqs[(size_t) i * seqlen + j] = s->cols[(size_t) c * nseqs + i];
    }
    return 0;
}
