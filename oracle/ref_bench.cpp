// ref_bench.cpp -- TEST INFRASTRUCTURE (oracle/_ref).  Runs the UNMODIFIED
// reference's own forward algorithm and stochastic traceback
// (src/argweaver/sample_thread.cpp:394-460, :522-569) on a flattened problem
// file and reports wall time; optionally dumps the reference outputs.
//
// Used for (a) bench.py --impl reference and the cpu_baseline leg, and
// (b) validating generated problems against the reference on the CPU.
//
// usage: ref_bench --in problem.awf [--out result.awf] [--reps R] [--threads N]
//                  [--rand-seed S] [--fw-stride K]
//   --fw-stride K : with --out, keep only the forward rows of sites 0, K, 2K, ...
//                 (and the last site) -- at-size parity tests compare those rows
//                 and the whole sampled path
//   --threads N : N independent processes-worth of work run by N forked
//                 children on the same problem (the reference is single
//                 threaded; this is how arg-sample-genome uses N cores).
//
// Prints one line:  forward_s=<t> trace_s=<t> sites=<n> states_sites=<sum>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <unistd.h>
#include <map>
#include <string>
#include <vector>

#include "argweaver/common.h"
#include "argweaver/local_tree.h"
#include "argweaver/logging.h"
#include "argweaver/matrices.h"
#include "argweaver/model.h"
#include "argweaver/recomb.h"
#include "argweaver/sample_thread.h"
#include "argweaver/sequences.h"
#include "argweaver/states.h"

#include "flatio.h"

using namespace argweaver;
using namespace std;

static double now_s()
{
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

struct Loaded {
    int phase_row1, phase_row2;     // unphased individual (leaf order), or -1
    awf_file *af;
    ArgModel *model;
    Sequences *sequences;
    LocalTrees *trees;
    int new_chrom;
    bool internal;
    vector<char *> rows;
};

static void load(const char *fn, Loaded &L)
{
    awf_file *af = awf_read(fn);
    if (!af) { fprintf(stderr, "cannot read %s\n", fn); exit(1); }
    L.af = af;
    const int T = awf_int(af, "ntimes");
    double *times = (double *) awf_need(af, "times")->data;
    double *popsizes = (double *) awf_need(af, "popsizes")->data;
    L.model = new ArgModel(T, times, popsizes, awf_double(af, "rho"),
                           awf_double(af, "mu"));
    // optional emission mode (ArgModel::infsites_penalty, model.h:348)
    if (awf_find(af, "infsites_penalty"))
        L.model->infsites_penalty = awf_double(af, "infsites_penalty");
    awf_array *seqs = awf_need(af, "seqs");
    const int nseqs = seqs->dims[0], seqlen = seqs->dims[1];
    for (int i = 0; i < nseqs; i++)
        L.rows.push_back((char *) seqs->data + (size_t) i * seqlen);
    L.sequences = new Sequences(&L.rows[0], nseqs, seqlen);
    L.internal = awf_int(af, "internal") != 0;
    L.new_chrom = awf_int(af, "new_chrom");
    // optional emission mode: unphased data (emit.cpp:705-742, :834-842)
    L.phase_row1 = L.phase_row2 = -1;
    if (awf_find(af, "phase_rows")) {
        L.phase_row1 = ((int *) awf_need(af, "phase_rows")->data)[0];
        L.phase_row2 = ((int *) awf_need(af, "phase_rows")->data)[1];
    }

    awf_array *pt = awf_need(af, "ptrees");
    const int B = pt->dims[0], V = pt->dims[1];
    int *ptrees = (int *) pt->data;
    int *ages = (int *) awf_need(af, "ages")->data;
    int *sprs = (int *) awf_need(af, "sprs")->data;
    int *blocklens = (int *) awf_need(af, "blocklens")->data;
    vector<int *> pp(B), aa(B), ss(B);
    for (int b = 0; b < B; b++) {
        pp[b] = ptrees + (size_t) b * V;
        aa[b] = ages + (size_t) b * V;
        ss[b] = sprs + (size_t) b * 4;
    }
    const int start = awf_find(af, "start_coord") ? awf_int(af, "start_coord") : 0;
    L.trees = new LocalTrees(&pp[0], &aa[0], &ss[0], blocklens, B, V, -1, start);

    // sequence ids of the leaves
    awf_array *sid = awf_need(af, "seqids");
    L.trees->seqids.clear();
    for (uint64_t i = 0; i < sid->count; i++)
        L.trees->seqids.push_back(((int *) sid->data)[i]);

    // explicit node mappings, if the file carries them
    awf_array *mp = awf_find(af, "mappings");
    if (mp) {
        int b = 0;
        for (LocalTrees::iterator it = L.trees->begin(); it != L.trees->end();
             ++it, ++b) {
            if (b == 0 || !it->mapping) continue;
            for (int j = 0; j < V; j++)
                it->mapping[j] = ((int *) mp->data)[(size_t) b * V + j];
        }
    }

    // internal mode: child[0] of the root must be the subtree root
    awf_array *sr = awf_find(af, "subtree_roots");
    if (L.internal && sr) {
        int b = 0;
        for (LocalTrees::iterator it = L.trees->begin(); it != L.trees->end();
             ++it, ++b) {
            LocalTree *tree = it->tree;
            int want = ((int *) sr->data)[b];
            int *c = tree->nodes[tree->root].child;
            if (want >= 0 && c[1] == want) {
                int tmp = c[0]; c[0] = c[1]; c[1] = tmp;
            }
        }
    }
}

static int g_fw_stride = 1;

static void run_once(Loaded &L, unsigned rand_seed, double *tf, double *tt,
                     const char *out_file)
{
    LocalTrees *trees = L.trees;
    const ArgModel *model = L.model;
    const int n = trees->length();

    // unphased: a PhaseProbs with the two rows set by hand (its constructor
    // wants name pairs; it returns at once for a phased model)
    PhaseProbs phase_pr(0, 0, L.sequences, trees, model);
    PhaseProbs *pp = NULL;
    if (L.phase_row1 >= 0) {
        const int nl = trees->get_num_leaves();
        phase_pr.treemap1 = L.phase_row1;
        phase_pr.treemap2 = L.phase_row2;
        phase_pr.hap1 = L.phase_row1 < nl ? trees->seqids[L.phase_row1] : L.new_chrom;
        phase_pr.hap2 = L.phase_row2 < nl ? trees->seqids[L.phase_row2] : L.new_chrom;
        phase_pr.seqs = L.sequences;
        phase_pr.offset = 0;
        phase_pr.probs.clear();
        L.model->unphased = true;
        pp = &phase_pr;
    }

    double t0 = now_s();
    ArgHmmForwardTable forward(trees->start_coord, n);
    ArgHmmMatrixIter matrix_iter(model, L.sequences, trees, L.new_chrom);
    matrix_iter.set_internal(L.internal, 0);
    arghmm_forward_alg(trees, model, L.sequences, &matrix_iter, &forward, pp,
                       false, L.internal);
    double t1 = now_s();

    srand(rand_seed);
    vector<int> path_alloc(n);
    int *thread_path = &path_alloc[0] - trees->start_coord;
    double **fw = forward.get_table();
    ArgHmmMatrixIter matrix_iter2(model, NULL, trees, L.new_chrom);
    matrix_iter2.set_internal(L.internal, 0);
    stochastic_traceback(trees, model, &matrix_iter2, fw, thread_path, false,
                         L.internal);
    double t2 = now_s();
    *tf = t1 - t0;
    *tt = t2 - t1;

    // P(phasing as given | sampled state) at the heterozygous sites, then the
    // phase draw (sample_thread.cpp:614-616: one frand() per such site, before
    // the recombination points)
    vector<int> phase_pos;
    vector<double> phase_p;
    if (pp) {
        for (map<int, vector<double> >::iterator it = phase_pr.probs.begin();
             it != phase_pr.probs.end(); ++it) {
            phase_pos.push_back(it->first - trees->start_coord);
            phase_p.push_back(it->second[thread_path[it->first]]);
        }
        if (out_file)
            phase_pr.sample_phase(thread_path);
        L.model->unphased = false;
    }

    if (out_file) {
        // the step after the traceback (sample_thread.cpp:617-622): the libc
        // stream continues where the traceback left it
        vector<int> recomb_pos;
        vector<NodePoint> recombs;
        sample_recombinations(trees, model, &matrix_iter2, thread_path,
                              recomb_pos, recombs, L.internal);
        vector<int> rnode, rtime;
        for (size_t i = 0; i < recombs.size(); i++) {
            recomb_pos[i] -= trees->start_coord;
            rnode.push_back(recombs[i].node);
            rtime.push_back(recombs[i].time);
        }
        const int next_rand = rand();   // where the stream stands afterwards

        FILE *out = awf_create(out_file);
        vector<double> fwflat;
        vector<int> nstates, fwsites;
        States states;
        int pos = trees->start_coord;
        for (LocalTrees::const_iterator it = trees->begin();
             it != trees->end(); ++it) {
            get_coal_states(it->tree, model->ntimes, states, L.internal);
            nstates.push_back(states.size());
            const int S1 = max((int) states.size(), 1);
            for (int i = pos; i < pos + it->blocklen; i++) {
                const int rel = i - trees->start_coord;
                if (rel % g_fw_stride != 0 && rel != n - 1)
                    continue;
                fwsites.push_back(rel);
                for (int j = 0; j < S1; j++)
                    fwflat.push_back(fw[i][j]);
            }
            pos += it->blocklen;
        }
        awf_write1(out, "nstates", AWF_I32, nstates.size(), &nstates[0]);
        awf_write1(out, "fw", AWF_F64, fwflat.size(), &fwflat[0]);
        awf_write1(out, "fw_sites", AWF_I32, fwsites.size(), &fwsites[0]);
        awf_write1(out, "path", AWF_I32, n, &path_alloc[0]);
        int zero = 0;
        awf_write1(out, "recomb_pos", AWF_I32, recomb_pos.size(),
                   recomb_pos.empty() ? &zero : &recomb_pos[0]);
        awf_write1(out, "recomb_node", AWF_I32, rnode.size(),
                   rnode.empty() ? &zero : &rnode[0]);
        awf_write1(out, "recomb_time", AWF_I32, rtime.size(),
                   rtime.empty() ? &zero : &rtime[0]);
        awf_write1(out, "next_rand", AWF_I32, 1, &next_rand);
        double dzero = 0;
        awf_write1(out, "phase_pos", AWF_I32, phase_pos.size(),
                   phase_pos.empty() ? &zero : &phase_pos[0]);
        awf_write1(out, "phase_p", AWF_F64, phase_p.size(),
                   phase_p.empty() ? &dzero : &phase_p[0]);
        fclose(out);
    }
}

int main(int argc, char **argv)
{
    const char *in_file = NULL, *out_file = NULL;
    int reps = 1, threads = 1;
    unsigned rand_seed = 1;
    for (int i = 1; i + 1 < argc; i += 2) {
        string a = argv[i];
        if (a == "--in") in_file = argv[i + 1];
        else if (a == "--out") out_file = argv[i + 1];
        else if (a == "--reps") reps = atoi(argv[i + 1]);
        else if (a == "--threads") threads = atoi(argv[i + 1]);
        else if (a == "--rand-seed") rand_seed = (unsigned) atol(argv[i + 1]);
        else if (a == "--fw-stride") g_fw_stride = max(1, atoi(argv[i + 1]));
        else { fprintf(stderr, "unknown option %s\n", argv[i]); return 1; }
    }
    if (!in_file) { fprintf(stderr, "need --in\n"); return 1; }
    setLogLevel(LOG_QUIET);

    Loaded L;
    load(in_file, L);

    // exact sum over blocks of blocklen * nstates
    double states_sites = 0;
    {
        States states;
        for (LocalTrees::const_iterator it = L.trees->begin();
             it != L.trees->end(); ++it) {
            get_coal_states(it->tree, L.model->ntimes, states, L.internal);
            states_sites += (double) states.size() * it->blocklen;
        }
    }

    double tf = 0, tt = 0;
    double wall0 = now_s();
    if (threads <= 1) {
        for (int r = 0; r < reps; r++) {
            double a, b;
            run_once(L, rand_seed, &a, &b, (r == 0) ? out_file : NULL);
            tf += a;
            tt += b;
        }
    } else {
        // N independent single-threaded workers on N cores
        vector<pid_t> kids;
        for (int w = 0; w < threads; w++) {
            pid_t pid = fork();
            if (pid == 0) {
                for (int r = 0; r < reps; r++) {
                    double a, b;
                    run_once(L, rand_seed + w, &a, &b, NULL);
                }
                _exit(0);
            }
            kids.push_back(pid);
        }
        for (size_t w = 0; w < kids.size(); w++) {
            int st;
            waitpid(kids[w], &st, 0);
        }
    }
    double wall = now_s() - wall0;
    printf("forward_s=%.6f trace_s=%.6f wall_s=%.6f reps=%d threads=%d sites=%d "
           "states_sites=%.0f ntrees=%d\n",
           tf / reps, tt / reps, wall, reps, threads, L.trees->length(),
           states_sites, L.trees->get_num_trees());
    return 0;
}
