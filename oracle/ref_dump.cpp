// ref_dump.cpp -- TEST INFRASTRUCTURE (oracle/_ref). Links against the
// UNMODIFIED reference objects (see oracle/Makefile) and writes one golden
// vector file: the flattened inputs of one thread-sampling call plus everything
// the reference computes for it (states, lineage counts, transition vectors,
// switch matrices, emissions, forward table, rand() draws, sampled path).
//
// It only calls the reference's public C++ interface; it copies no reference
// code.  Call sequence mirrors sample_arg_thread / sample_arg_thread_internal
// (reference src/argweaver/sample_thread.cpp:578-694) and the data flow of
// src/arg-sample.cpp:970-1160 (read sites -> compress -> sequences -> model).
//
// usage: ref_dump --sites F --out O [--region a-b] [--ntimes T] [--maxtime M]
//                 [--popsize N] [--rho r] [--mu m] [--compress c] [--seed s]
//                 [--mode external|internal-leaf|internal-uniform]
//                 [--refine n]

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "argweaver/common.h"
#include "argweaver/compress.h"
#include "argweaver/emit.h"
#include "argweaver/local_tree.h"
#include "argweaver/logging.h"
#include "argweaver/matrices.h"
#include "argweaver/model.h"
#include "argweaver/sample_arg.h"
#include "argweaver/sample_thread.h"
#include "argweaver/sequences.h"
#include "argweaver/states.h"
#include "argweaver/thread.h"
#include "argweaver/trans.h"

#include "flatio.h"

using namespace argweaver;
using namespace std;

static void dump_inputs(FILE *out, const ArgModel &model,
                        const Sequences &sequences, const LocalTrees *trees,
                        int new_chrom, bool internal)
{
    const int T = model.ntimes;
    awf_write_int(out, "ntimes", T);
    awf_write1(out, "times", AWF_F64, T, model.times);
    awf_write1(out, "popsizes", AWF_F64, T, model.popsizes);
    awf_write_double(out, "rho", model.rho);
    awf_write_double(out, "mu", model.mu);
    awf_write_int(out, "internal", internal ? 1 : 0);
    awf_write_int(out, "minage", 0);
    awf_write_int(out, "new_chrom", new_chrom);
    awf_write_int(out, "start_coord", trees->start_coord);
    awf_write_int(out, "end_coord", trees->end_coord);

    // sequences (dense rows, as the reference's Sequences holds them)
    const int nseqs = sequences.get_num_seqs();
    const int seqlen = sequences.length();
    vector<unsigned char> rows((size_t) nseqs * seqlen);
    for (int i = 0; i < nseqs; i++)
        memcpy(&rows[(size_t) i * seqlen], sequences.seqs[i], seqlen);
    awf_write2(out, "seqs", AWF_U8, nseqs, seqlen, &rows[0]);

    const int nleaves = trees->get_num_leaves();
    vector<int> seqids(trees->seqids.begin(), trees->seqids.end());
    awf_write1(out, "seqids", AWF_I32, nleaves, &seqids[0]);

    // local trees
    const int B = trees->get_num_trees();
    const int V = trees->nnodes;
    vector<int> ptrees((size_t) B * V), ages((size_t) B * V),
        mappings((size_t) B * V), child0((size_t) B * V),
        child1((size_t) B * V), sprs((size_t) B * 4), blocklens(B), roots(B);
    int b = 0;
    for (LocalTrees::const_iterator it = trees->begin(); it != trees->end();
         ++it, ++b) {
        const LocalTree *tree = it->tree;
        if (tree->nnodes != V) {
            fprintf(stderr, "tree %d has %d nodes, expected %d\n", b,
                    tree->nnodes, V);
            exit(1);
        }
        for (int j = 0; j < V; j++) {
            ptrees[(size_t) b * V + j] = tree->nodes[j].parent;
            ages[(size_t) b * V + j] = tree->nodes[j].age;
            child0[(size_t) b * V + j] = tree->nodes[j].child[0];
            child1[(size_t) b * V + j] = tree->nodes[j].child[1];
            mappings[(size_t) b * V + j] = it->mapping ? it->mapping[j] : -2;
        }
        sprs[b * 4 + 0] = it->spr.recomb_node;
        sprs[b * 4 + 1] = it->spr.recomb_time;
        sprs[b * 4 + 2] = it->spr.coal_node;
        sprs[b * 4 + 3] = it->spr.coal_time;
        blocklens[b] = it->blocklen;
        roots[b] = tree->root;
    }
    awf_write_int(out, "ntrees", B);
    awf_write_int(out, "nnodes", V);
    awf_write2(out, "ptrees", AWF_I32, B, V, &ptrees[0]);
    awf_write2(out, "ages", AWF_I32, B, V, &ages[0]);
    awf_write2(out, "child0", AWF_I32, B, V, &child0[0]);
    awf_write2(out, "child1", AWF_I32, B, V, &child1[0]);
    awf_write2(out, "mappings", AWF_I32, B, V, &mappings[0]);
    awf_write2(out, "sprs", AWF_I32, B, 4, &sprs[0]);
    awf_write1(out, "blocklens", AWF_I32, B, &blocklens[0]);
    awf_write1(out, "roots", AWF_I32, B, &roots[0]);
}


// Walk the per-block matrices exactly as arghmm_forward_alg does and record them.
static void dump_matrices(FILE *out, const ArgModel &model,
                          const Sequences &sequences, const LocalTrees *trees,
                          int new_chrom, bool internal)
{
    const int T = model.ntimes;
    const int B = trees->get_num_trees();

    ArgHmmMatrixIter iter(&model, &sequences, trees, new_chrom);
    iter.set_internal(internal, 0);

    vector<int> nstates, states_flat, nbr, nrc, ncl, tm_minage;
    vector<int> sw_determ, sw_recombsrc, sw_recoalsrc;
    vector<double> tmD, tmE, tmlnB, tmlnE2, tmlnNegG1, tmG2, tmG3, tmlnG4,
        tmnorecombs, sw_determprob, sw_recombrow, sw_recoalrow, emit;
    vector<int64_t> row_off(1, 0), fw_off(1, 0), sw1_off(1, 0);

    LineageCounts lineages(T);
    States states;
    int b = 0;
    for (iter.begin(); iter.more(); iter.next(), b++) {
        ArgHmmMatrices &m = iter.ref_matrices(NULL);
        const LocalTree *tree = iter.get_tree_spr()->tree;
        iter.get_coal_states(tree, states);
        const int S = states.size();
        const int S1 = max(S, 1);
        nstates.push_back(S);
        for (int j = 0; j < S; j++) {
            states_flat.push_back(states[j].node);
            states_flat.push_back(states[j].time);
        }
        row_off.push_back(row_off.back() + S1);
        fw_off.push_back(fw_off.back() + (int64_t) S1 * m.blocklen);

        lineages.count(tree, internal);
        for (int t = 0; t < T; t++) {
            nbr.push_back(lineages.nbranches[t]);
            nrc.push_back(lineages.nrecombs[t]);
            ncl.push_back(lineages.ncoals[t]);
        }

        const TransMatrix *tm = m.transmat;
        for (int t = 0; t < T; t++) {
            const bool ok = t < T - 1;
            tmD.push_back(ok ? tm->D[t] : 0.0);
            tmE.push_back(ok ? tm->E[t] : 0.0);
            tmlnB.push_back(ok ? tm->lnB[t] : 0.0);
            tmlnE2.push_back(ok ? tm->lnE2[t] : 0.0);
            tmlnNegG1.push_back(ok ? tm->lnNegG1[t] : 0.0);
            tmG2.push_back(ok ? tm->G2[t] : 0.0);
            tmG3.push_back(ok ? tm->G3[t] : 0.0);
            tmlnG4.push_back(ok ? tm->lnG4[t] : 0.0);
            tmnorecombs.push_back(ok ? tm->norecombs[t] : 0.0);
        }
        tm_minage.push_back(tm->minage);

        const TransMatrixSwitch *sw = m.transmat_switch;
        if (sw) {
            const int n1 = max(sw->nstates1, 1);
            const int n2 = max(sw->nstates2, 1);
            const bool have_rows = sw->nstates1 > 0 && sw->nstates2 > 0;
            for (int j = 0; j < n1; j++) {
                // determprob is only defined where determ >= 0 and the row is
                // not one of the two dense rows
                const bool defined = (sw->determ[j] >= 0) &&
                    !(have_rows && (j == sw->recombsrc || j == sw->recoalsrc));
                sw_determ.push_back(sw->determ[j]);
                sw_determprob.push_back(defined ? sw->determprob[j] : 0.0);
            }
            for (int j = 0; j < n2; j++) {
                sw_recombrow.push_back(
                    (have_rows && sw->recombsrc != -1) ? sw->recombrow[j] : 0.0);
                sw_recoalrow.push_back(
                    have_rows ? sw->recoalrow[j] : 0.0);
            }
            sw_recombsrc.push_back(sw->recombsrc);
            sw_recoalsrc.push_back(sw->recoalsrc);
            sw1_off.push_back(sw1_off.back() + n1);
        } else {
            for (int j = 0; j < S1; j++) {
                sw_recombrow.push_back(0.0);
                sw_recoalrow.push_back(0.0);
            }
            sw_recombsrc.push_back(-1);
            sw_recoalsrc.push_back(-1);
            sw1_off.push_back(sw1_off.back());
        }

        for (int i = 0; i < m.blocklen; i++)
            for (int j = 0; j < S1; j++)
                emit.push_back(m.emit[i][j]);
    }
    if (b != B) {
        fprintf(stderr, "block count mismatch %d vs %d\n", b, B);
        exit(1);
    }

    int dummy_i = 0;
    double dummy_d = 0;
    awf_write1(out, "nstates", AWF_I32, B, &nstates[0]);
    awf_write2(out, "states", AWF_I32, states_flat.size() / 2, 2,
               states_flat.empty() ? &dummy_i : &states_flat[0]);
    awf_write1(out, "row_off", AWF_I64, B + 1, &row_off[0]);
    awf_write1(out, "fw_off", AWF_I64, B + 1, &fw_off[0]);
    awf_write1(out, "sw1_off", AWF_I64, B + 1, &sw1_off[0]);
    awf_write2(out, "nbranches", AWF_I32, B, T, &nbr[0]);
    awf_write2(out, "nrecombs", AWF_I32, B, T, &nrc[0]);
    awf_write2(out, "ncoals", AWF_I32, B, T, &ncl[0]);
    awf_write2(out, "tm_D", AWF_F64, B, T, &tmD[0]);
    awf_write2(out, "tm_E", AWF_F64, B, T, &tmE[0]);
    awf_write2(out, "tm_lnB", AWF_F64, B, T, &tmlnB[0]);
    awf_write2(out, "tm_lnE2", AWF_F64, B, T, &tmlnE2[0]);
    awf_write2(out, "tm_lnNegG1", AWF_F64, B, T, &tmlnNegG1[0]);
    awf_write2(out, "tm_G2", AWF_F64, B, T, &tmG2[0]);
    awf_write2(out, "tm_G3", AWF_F64, B, T, &tmG3[0]);
    awf_write2(out, "tm_lnG4", AWF_F64, B, T, &tmlnG4[0]);
    awf_write2(out, "tm_norecombs", AWF_F64, B, T, &tmnorecombs[0]);
    awf_write1(out, "tm_minage", AWF_I32, B, &tm_minage[0]);
    awf_write1(out, "sw_determ", AWF_I32, sw_determ.size(),
               sw_determ.empty() ? &dummy_i : &sw_determ[0]);
    awf_write1(out, "sw_determprob", AWF_F64, sw_determprob.size(),
               sw_determprob.empty() ? &dummy_d : &sw_determprob[0]);
    awf_write1(out, "sw_recombrow", AWF_F64, sw_recombrow.size(),
               &sw_recombrow[0]);
    awf_write1(out, "sw_recoalrow", AWF_F64, sw_recoalrow.size(),
               &sw_recoalrow[0]);
    awf_write1(out, "sw_recombsrc", AWF_I32, B, &sw_recombsrc[0]);
    awf_write1(out, "sw_recoalsrc", AWF_I32, B, &sw_recoalsrc[0]);
    awf_write1(out, "emit", AWF_F64, emit.size(), &emit[0]);
}


int main(int argc, char **argv)
{
    string sites_file, out_file, region, mode = "external";
    int ntimes = 20, compress = 10, seed = 1, refine = 0;
    double maxtime = 200e3, popsize = 1e4, rho = 1.6e-8, mu = 1.8e-8;

    for (int i = 1; i < argc; i++) {
        string a = argv[i];
        if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", argv[i]); return 1; }
        const char *v = argv[++i];
        if (a == "--sites") sites_file = v;
        else if (a == "--out") out_file = v;
        else if (a == "--region") region = v;
        else if (a == "--ntimes") ntimes = atoi(v);
        else if (a == "--maxtime") maxtime = atof(v);
        else if (a == "--popsize") popsize = atof(v);
        else if (a == "--rho") rho = atof(v);
        else if (a == "--mu") mu = atof(v);
        else if (a == "--compress") compress = atoi(v);
        else if (a == "--seed") seed = atoi(v);
        else if (a == "--mode") mode = v;
        else if (a == "--refine") refine = atoi(v);
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    if (sites_file == "" || out_file == "") {
        fprintf(stderr, "need --sites and --out\n");
        return 1;
    }
    setLogLevel(LOG_QUIET);
    srand(seed);

    // read + compress sites (arg-sample.cpp:960-1007)
    int sub0 = -1, sub1 = -1;
    if (region != "") {
        if (sscanf(region.c_str(), "%d-%d", &sub0, &sub1) != 2) {
            fprintf(stderr, "bad region\n");
            return 1;
        }
        sub0 -= 1;
    }
    Sites sites;
    if (!read_sites(sites_file.c_str(), &sites, sub0, sub1)) {
        fprintf(stderr, "cannot read sites\n");
        return 1;
    }
    SitesMapping sites_mapping;
    if (!find_compress_cols(&sites, compress, &sites_mapping)) {
        fprintf(stderr, "cannot compress\n");
        return 1;
    }
    compress_sites(&sites, &sites_mapping);
    Sequences sequences;
    make_sequences_from_sites(&sites, &sequences);
    const int nseqs = sequences.get_num_seqs();

    // model (arg-sample.cpp:1042-1076, compress_model :428-436)
    ArgModel model(ntimes, maxtime, popsize, rho * compress, mu * compress);

    const bool internal = (mode != "external");
    const int narg = internal ? nseqs : nseqs - 1;
    int new_chrom = internal ? -1 : nseqs - 1;

    // build the partial ARG with the reference's own sampler
    Sequences arg_seqs(&sequences, narg, sequences.length());
    LocalTrees trees(0, sequences.length());
    sample_arg_seq(&model, &arg_seqs, &trees);
    for (int i = 0; i < refine; i++)
        resample_arg_leaf(&model, &arg_seqs, &trees);

    if (internal) {
        const int maxt = model.get_removed_root_time();
        vector<int> removal_path(trees.get_num_trees());
        if (mode == "internal-leaf") {
            int node = irand(trees.get_num_leaves());
            sample_arg_removal_leaf_path(&trees, node, &removal_path[0]);
        } else if (mode == "internal-uniform") {
            sample_arg_removal_path_uniform(&trees, &removal_path[0]);
        } else {
            fprintf(stderr, "unknown mode %s\n", mode.c_str());
            return 1;
        }
        remove_arg_thread_path(&trees, &removal_path[0], maxt);
    }

    FILE *out = awf_create(out_file.c_str());
    if (!out) { fprintf(stderr, "cannot write %s\n", out_file.c_str()); return 1; }
    dump_inputs(out, model, sequences, &trees, new_chrom, internal);
    dump_matrices(out, model, sequences, &trees, new_chrom, internal);

    // forward (sample_thread.cpp:582-603 / :644-663)
    const int n = trees.length();
    ArgHmmForwardTable forward(trees.start_coord, n);
    ArgHmmMatrixIter matrix_iter(&model, &sequences, &trees, new_chrom);
    matrix_iter.set_internal(internal, 0);
    arghmm_forward_alg(&trees, &model, &sequences, &matrix_iter, &forward,
                       NULL, false, internal);
    double **fw = forward.get_table();

    // flatten fw using the per-block state counts
    {
        vector<double> fwflat;
        States states;
        int pos = trees.start_coord;
        for (LocalTrees::const_iterator it = trees.begin(); it != trees.end();
             ++it) {
            get_coal_states(it->tree, ntimes, states, internal);
            const int S1 = max((int) states.size(), 1);
            for (int i = pos; i < pos + it->blocklen; i++)
                for (int j = 0; j < S1; j++)
                    fwflat.push_back(fw[i][j]);
            pos += it->blocklen;
        }
        awf_write1(out, "fw", AWF_F64, fwflat.size(), &fwflat[0]);
    }

    // pre-draw the rand() values the traceback will consume, then rewind
    const unsigned tb_seed = 7919u * (unsigned) seed + 13u;
    srand(tb_seed);
    vector<int> rand_ints(n);
    for (int i = 0; i < n; i++)
        rand_ints[i] = rand();
    awf_write1(out, "rand_ints", AWF_I32, n, &rand_ints[0]);
    awf_write_int(out, "rand_max", RAND_MAX);
    srand(tb_seed);

    // traceback (sample_thread.cpp:606-611 / :666-673)
    vector<int> path_alloc(n);
    int *thread_path = &path_alloc[0] - trees.start_coord;
    ArgHmmMatrixIter matrix_iter2(&model, NULL, &trees, new_chrom);
    matrix_iter2.set_internal(internal, 0);
    stochastic_traceback(&trees, &model, &matrix_iter2, fw, thread_path,
                         false, internal);
    awf_write1(out, "path", AWF_I32, n, &path_alloc[0]);

    // how many rand() calls did the traceback consume?
    {
        int next = rand();
        srand(tb_seed);
        int used = -1;
        for (int i = 0; i <= n; i++) {
            if (rand() == next && used < 0) { used = i; }
        }
        awf_write_int(out, "rand_used", used);
    }

    fclose(out);
    fprintf(stderr, "ref_dump: mode=%s nseqs=%d sites=%d trees=%d nnodes=%d -> %s\n",
            mode.c_str(), nseqs, n, trees.get_num_trees(), trees.nnodes,
            out_file.c_str());
    return 0;
}
