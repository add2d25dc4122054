// sample_thread_b200.cpp -- the reference-side adapter of INTEGRATION.md section 2,
// compiled for real (TEST INFRASTRUCTURE: it includes the reference's headers by
// path and is linked with the reference's own objects by oracle/Makefile into
// oracle/_ref/arg-sample-b200 and oracle/_ref/libargweaver_dropin.so; nothing in
// the product links it).
//
// It defines the reference's L2 wrappers
//     sample_arg_thread                 sample_thread.cpp:578-640
//     sample_arg_thread_internal        sample_thread.cpp:644-700
//     cond_sample_arg_thread            sample_thread.cpp:706-775
//     cond_sample_arg_thread_internal   sample_thread.cpp:781-865
//     resample_arg_thread               sample_thread.cpp:869-875
// and the C exports arghmm_sample_thread (:1010) and
// arghmm_sample_arg_thread_internal (:981) with the forward pass and the
// stochastic traceback AND the sampling of the recombination points
// (sample_recombinations, recomb.cpp:151-235) replaced by ONE call into
// libargweaver_b200.so (awb_thread_sample_recombs).  What follows -- the ARG
// surgery add_arg_thread[_path] -- is the reference's own code, called exactly
// as the reference calls it.  The reference's sample_thread.cpp is
// compiled next to this file with those seven names renamed to ref_* on the
// compiler command line (no source is modified or copied), so the original
// bodies stay available as the fallback for model features the device path does
// not cover (--unphased together with --infsites, real --mutmap/--recombmap
// files with several regions).
//
// libc rand(): stochastic_traceback consumes one rand() per sampled site, last
// site first (common.h:272-290).  The adapter draws exactly those values up
// front, in that order, and ships them.  sample_recombinations takes a
// data-dependent number of further draws: the library snapshots the process's
// rand() state, runs glibc's generator on the device and advances the process's
// stream by the number of draws it used (awb_libc_rand_snapshot / _advance).
// Either way the stream is left where the reference would leave it, so every
// later consumer sees the same draws and a seeded arg-sample run reproduces the
// reference's .stats rows.  AWB_ADAPTER_HOST_RECOMBS=1 keeps the reference's own
// sample_recombinations (the round-1 arrangement).

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
#include "argweaver/thread.h"

#include "argweaver_b200.h"

namespace argweaver {

using namespace std;

// the reference's own bodies (sample_thread.cpp compiled with -Dname=ref_name)
void ref_sample_arg_thread(const ArgModel *model, Sequences *sequences,
                           LocalTrees *trees, int new_chrom);
void ref_sample_arg_thread_internal(const ArgModel *model,
                                    const Sequences *sequences,
                                    LocalTrees *trees, int minage,
                                    PhaseProbs *phase_pr);
void ref_cond_sample_arg_thread(const ArgModel *model,
                                const Sequences *sequences, LocalTrees *trees,
                                int new_chrom, State start_state,
                                State end_state);
void ref_cond_sample_arg_thread_internal(const ArgModel *model,
                                         const Sequences *sequences,
                                         LocalTrees *trees, State start_state,
                                         State end_state);

// counters a driver can read (how many thread samples ran on the device / fell
// back to the reference's CPU code)
extern "C" {
long awb_adapter_device_calls = 0;
long awb_adapter_fallback_calls = 0;
double awb_adapter_device_seconds = 0;
double awb_adapter_states_sites = 0;
}

// AWB_ADAPTER_REPORT=1: one line on stderr when the process ends
struct AdapterReport {
    ~AdapterReport() {
        if (getenv("AWB_ADAPTER_REPORT"))
            fprintf(stderr, "argweaver_b200 adapter: device thread samples: %ld "
                    "reference fallbacks: %ld device seconds: %.3f\n",
                    awb_adapter_device_calls, awb_adapter_fallback_calls,
                    awb_adapter_device_seconds);
    }
};
static AdapterReport g_report;

// Can the device path take this model?  One rate region only (arg-sample always
// installs one-region maps, arg-sample.cpp:1113; more regions only come from map
// files), phased data.
static bool device_covers(const ArgModel *model, const PhaseProbs *phase_pr)
{
    if (getenv("AWB_ADAPTER_FORCE_REFERENCE"))
        return false;
    if ((model->unphased || phase_pr) && model->infsites_penalty < 1.0)
        return false;               // (this combination stays on the reference path)
    if (model->has_mutmap() && model->mutmap.size() != 1)
        return false;
    if (model->has_recombmap() && model->recombmap.size() != 1)
        return false;
    return true;
}

// LocalTrees / ArgModel / Sequences -> awb_problem (owns the flat copies)
struct FlatProblem {
    vector<int> ptrees, ages, sprs, mappings, blocklens, seqids, roots;
    vector<unsigned char> seqs;
    ArgModel local;
    awb_problem p;

    FlatProblem(const ArgModel *model, const Sequences *sequences,
                const LocalTrees *trees, int new_chrom, bool internal,
                int minage)
    {
        const int V = trees->nnodes, B = trees->get_num_trees();
        ptrees.reserve((size_t) B * V);
        ages.reserve((size_t) B * V);
        mappings.reserve((size_t) B * V);
        for (LocalTrees::const_iterator it = trees->begin();
             it != trees->end(); ++it) {
            const LocalNode *nodes = it->tree->nodes;
            for (int j = 0; j < V; j++) {
                ptrees.push_back(nodes[j].parent);
                ages.push_back(nodes[j].age);
                mappings.push_back(it->mapping ? it->mapping[j] : j);
            }
            sprs.push_back(it->spr.recomb_node);
            sprs.push_back(it->spr.recomb_time);
            sprs.push_back(it->spr.coal_node);
            sprs.push_back(it->spr.coal_time);
            blocklens.push_back(it->blocklen);
            roots.push_back(internal ? nodes[it->tree->root].child[0] : -1);
        }
        const int L = sequences->length();
        const int nseqs = sequences->get_num_seqs();
        seqs.resize((size_t) nseqs * L);
        for (int i = 0; i < nseqs; i++)
            memcpy(&seqs[(size_t) i * L], sequences->seqs[i], L);
        seqids = trees->seqids;

        // rates of the (single) region, as calc_matrices takes them
        // (matrices.h:342-352, model.h:298-310)
        model->get_local_model_index(model->has_mutmap() ? 0 : -1, local);

        memset(&p, 0, sizeof(p));
        p.ntimes = local.ntimes;
        p.times = local.times;
        p.popsizes = local.popsizes;
        p.rho = local.rho;
        p.mu = local.mu;
        p.nseqs = nseqs;
        p.seqlen = L;
        p.seqs = seqs.data();
        p.nleaves = trees->get_num_leaves();
        p.seqids = seqids.data();
        p.new_chrom = new_chrom;
        p.internal = internal ? 1 : 0;
        p.minage = minage;
        p.ntrees = B;
        p.nnodes = V;
        p.start_coord = trees->start_coord;
        p.ptrees = ptrees.data();
        p.ages = ages.data();
        p.sprs = sprs.data();
        p.mappings = mappings.data();
        p.blocklens = blocklens.data();
        p.subtree_roots = internal ? roots.data() : NULL;
        // ArgModel::infsites_penalty (model.h:348); 1.0 = off
        p.infsites_penalty = model->infsites_penalty;
    }
};

// one context (device + stream) for the life of the process
static awb_ctx *adapter_ctx()
{
    static awb_ctx *ctx = NULL;
    if (!ctx) {
        const char *env = getenv("AWB_DEVICE");
        if (awb_ctx_create(env ? atoi(env) : 0, &ctx)) {
            printError("argweaver_b200: %s", awb_last_error());
            abort();
        }
    }
    return ctx;
}

static void device_fail()
{
    printError("argweaver_b200: %s", awb_last_error());
    abort();                            // the reference's error convention
}

// forward + traceback (+ phase draw, + recombination points) on the device.
// path_alloc[0..n): state per site.  recomb_pos / recombs: filled when given
// (absolute coordinates, as sample_recombinations returns them).  phase_pr:
// unphased data -- its probabilities are filled at the sampled states and
// sample_phase is called where the reference calls it (after the traceback,
// before the recombination points).
static void device_thread(const ArgModel *model, const Sequences *sequences,
                          const LocalTrees *trees, int new_chrom, bool internal,
                          int minage, const double *prior, int last_state,
                          int *path_alloc, vector<int> *recomb_pos = NULL,
                          vector<NodePoint> *recombs = NULL,
                          PhaseProbs *phase_pr = NULL)
{
    const int n = trees->length();
    Timer time;
    FlatProblem fp(model, sequences, trees, new_chrom, internal, minage);
    // unphased: only when both haplotypes are rows of this call (emit.cpp:711-714)
    const int nrows = trees->get_num_leaves() + (internal ? 0 : 1);
    const bool phase = model->unphased && phase_pr &&
        phase_pr->treemap1 >= 0 && phase_pr->treemap1 < nrows &&
        phase_pr->treemap2 >= 0 && phase_pr->treemap2 < nrows;
    if (phase) {
        fp.p.unphased = 1;
        fp.p.phase_row1 = phase_pr->treemap1;
        fp.p.phase_row2 = phase_pr->treemap2;
    }
    // the draws stochastic_traceback would take, in its order
    const int ndraws = last_state >= 0 ? n - 1 : n;
    vector<int> draws(n > 0 ? n : 1);
    for (int i = 0; i < ndraws; i++)
        draws[i] = rand();

    awb_batch *b = NULL;
    if (awb_batch_create(adapter_ctx(), 1, &fp.p, 0, &b))
        device_fail();
    const double *priors[1] = { prior };
    const int *rands[1] = { draws.data() };
    if (awb_batch_upload(b) || awb_batch_setup(b) ||
        awb_batch_forward(b, prior ? priors : NULL) ||
        awb_batch_traceback(b, rands, RAND_MAX, last_state >= 0 ? &last_state : NULL) ||
        awb_batch_sync(b))
        device_fail();
    int bad = -1;
    if (awb_batch_get_status(b, 0, &bad))
        device_fail();
    if (bad >= 0) {
        printError("argweaver_b200: forward column %d has no positive entry", bad);
        abort();                        // sample_thread.cpp:443-444
    }
    if (awb_batch_get_path(b, 0, path_alloc))
        device_fail();

    if (model->unphased && phase_pr) {
        if (phase) {
            vector<double> pp(n > 0 ? n : 1);
            if (awb_batch_phase_probs(b) || awb_batch_get_phase_probs(b, 0, pp.data()))
                device_fail();
            phase_pr->probs.clear();
            for (int i = 0; i < n; i++) {
                if (pp[i] < 0)
                    continue;
                vector<double> v(path_alloc[i] + 1, 0.0);
                v[path_alloc[i]] = pp[i];
                phase_pr->probs[i + trees->start_coord] = v;
            }
        }
        // sample_thread.cpp:614-616 / :675-676
        phase_pr->sample_phase(&path_alloc[-trees->start_coord]);
    }

    if (recomb_pos && !getenv("AWB_ADAPTER_HOST_RECOMBS")) {
        // the process's rand() stream as it stands now; the device runs glibc's
        // generator from there and says how many draws it took
        int state[AWB_RNG_WORDS];
        vector<int> pos(n > 0 ? n : 1), node(n > 0 ? n : 1), rtime(n > 0 ? n : 1);
        int nrec = 0, used = 0;
        if (awb_libc_rand_snapshot(state) ||
            awb_batch_sample_recombs(b, state, RAND_MAX) ||
            awb_batch_get_recomb_count(b, 0, &nrec, &used) ||
            awb_batch_get_recombs(b, 0, nrec, pos.data(), node.data(), rtime.data()))
            device_fail();
        awb_libc_rand_advance(used);
        for (int i = 0; i < nrec; i++) {
            recomb_pos->push_back(pos[i] + trees->start_coord);
            recombs->push_back(NodePoint(node[i], rtime[i]));
        }
    } else if (recomb_pos) {
        ArgHmmMatrixIter matrix_iter2(model, NULL, trees, new_chrom);
        matrix_iter2.set_internal(internal, minage);
        sample_recombinations(trees, model, &matrix_iter2,
                              &path_alloc[-trees->start_coord], *recomb_pos,
                              *recombs, internal);
    }
    awb_batch_destroy(b);
    awb_adapter_device_calls++;
    awb_adapter_device_seconds += time.time();
    printTimerLog(time, LOG_LOW, "forward+trace on the device (%6d blocks):",
                  trees->get_num_trees());
}

void sample_arg_thread(const ArgModel *model, Sequences *sequences,
                       LocalTrees *trees, int new_chrom)
{
    if (!device_covers(model, NULL)) {
        awb_adapter_fallback_calls++;
        ref_sample_arg_thread(model, sequences, trees, new_chrom);
        return;
    }
    int *thread_path_alloc = new int [trees->length()];
    int *thread_path = &thread_path_alloc[-trees->start_coord];

    ArgHmmMatrixIter matrix_iter(model, sequences, trees, new_chrom);
    // (sample_thread.cpp:586-591)
    PhaseProbs phase_pr(new_chrom, trees->get_num_leaves(),
                        sequences, trees, model);
    if (model->unphased)
        printf("treemap = %i %i\n", phase_pr.treemap1, phase_pr.treemap2);

    vector<int> recomb_pos;
    vector<NodePoint> recombs;
    device_thread(model, sequences, trees, new_chrom, false, 0, NULL, -1,
                  thread_path_alloc, &recomb_pos, &recombs,
                  model->unphased ? &phase_pr : NULL);

    // add thread to ARG: the reference's code
    Timer time;
    add_arg_thread(trees, matrix_iter.states_model,
                   model->ntimes, thread_path, new_chrom,
                   recomb_pos, recombs);
    printTimerLog(time, LOG_LOW, "add thread:                         ");
    delete [] thread_path_alloc;
}


void sample_arg_thread_internal(
    const ArgModel *model, const Sequences *sequences, LocalTrees *trees,
    int minage, PhaseProbs *phase_pr)
{
    if (!device_covers(model, phase_pr)) {
        awb_adapter_fallback_calls++;
        ref_sample_arg_thread_internal(model, sequences, trees, minage,
                                       phase_pr);
        return;
    }
    const bool internal = true;
    int *thread_path_alloc = new int [trees->length()];
    int *thread_path = &thread_path_alloc[-trees->start_coord];

    ArgHmmMatrixIter matrix_iter(model, sequences, trees);
    matrix_iter.set_internal(internal, minage);
    if (phase_pr != NULL)               // (sample_thread.cpp:652-653)
        printf("treemap = %i %i\n", phase_pr->treemap1, phase_pr->treemap2);
    vector<int> recomb_pos;
    vector<NodePoint> recombs;
    device_thread(model, sequences, trees, -1, internal, minage, NULL, -1,
                  thread_path_alloc, &recomb_pos, &recombs, phase_pr);

    Timer time;
    add_arg_thread_path(trees, matrix_iter.states_model,
                        model->ntimes, thread_path,
                        recomb_pos, recombs);
    printTimerLog(time, LOG_LOW, "add thread:                         ");
    delete [] thread_path_alloc;
}


void cond_sample_arg_thread(const ArgModel *model, const Sequences *sequences,
                            LocalTrees *trees, int new_chrom,
                            State start_state, State end_state)
{
    if (!device_covers(model, NULL)) {
        awb_adapter_fallback_calls++;
        ref_cond_sample_arg_thread(model, sequences, trees, new_chrom,
                                   start_state, end_state);
        return;
    }
    int *thread_path_alloc = new int [trees->length()];
    int *thread_path = &thread_path_alloc[-trees->start_coord];
    States states;

    // one-hot first column at the given start state (sample_thread.cpp:722-731)
    ArgHmmMatrixIter matrix_iter(model, sequences, trees, new_chrom);
    matrix_iter.get_coal_states(trees->front().tree, states);
    int j = find_vector(states, start_state);
    assert(j != -1);
    vector<double> prior(max((int) states.size(), 1), 0.0);
    prior[j] = 1.0;

    // given last state (:741-745)
    matrix_iter.get_coal_states(trees->back().tree, states);
    const int last = find_vector(states, end_state);
    assert(last != -1);

    vector<int> recomb_pos;
    vector<NodePoint> recombs;
    device_thread(model, sequences, trees, new_chrom, false, 0, prior.data(),
                  last, thread_path_alloc, &recomb_pos, &recombs);
    assert(thread_path[trees->start_coord] == j);

    add_arg_thread(trees, matrix_iter.states_model,
                   model->ntimes, thread_path, new_chrom,
                   recomb_pos, recombs);
    delete [] thread_path_alloc;
}


void cond_sample_arg_thread_internal(
    const ArgModel *model, const Sequences *sequences, LocalTrees *trees,
    const State start_state, const State end_state)
{
    if (!device_covers(model, NULL)) {
        awb_adapter_fallback_calls++;
        ref_cond_sample_arg_thread_internal(model, sequences, trees,
                                            start_state, end_state);
        return;
    }
    const bool internal = true;
    int *thread_path_alloc = new int [trees->length()];
    int *thread_path = &thread_path_alloc[-trees->start_coord];
    States states;

    ArgHmmMatrixIter matrix_iter(model, sequences, trees);
    matrix_iter.set_internal(internal);

    // first column (sample_thread.cpp:795-812): one-hot at the start state, the
    // model's prior when the start is open, 1 for a fully specified tree
    matrix_iter.get_coal_states(trees->front().tree, states);
    vector<double> prior(max((int) states.size(), 1), 0.0);
    const double *prior_p = prior.data();
    int j = -1;
    if (states.size() > 0) {
        if (!start_state.is_null()) {
            j = find_vector(states, start_state);
            assert(j != -1);
            prior[j] = 1.0;
        } else {
            prior_p = NULL;
        }
    } else {
        prior[0] = 1.0;
    }

    // last state (:825-839)
    int last = -1;
    matrix_iter.get_coal_states(trees->back().tree, states);
    if (states.size() > 0) {
        if (!end_state.is_null()) {
            last = find_vector(states, end_state);
            assert(last != -1);
        }
    } else {
        last = 0;
    }

    vector<int> recomb_pos;
    vector<NodePoint> recombs;
    device_thread(model, sequences, trees, -1, internal, 0, prior_p, last,
                  thread_path_alloc, &recomb_pos, &recombs);
    if (j >= 0)
        assert(thread_path[trees->start_coord] == j);

    Timer time;
    add_arg_thread_path(trees, matrix_iter.states_model,
                        model->ntimes, thread_path,
                        recomb_pos, recombs);
    printTimerLog(time, LOG_LOW, "add thread:                         ");
    delete [] thread_path_alloc;
}


void resample_arg_thread(const ArgModel *model, Sequences *sequences,
                         LocalTrees *trees, int chrom)
{
    remove_arg_thread(trees, chrom);
    sample_arg_thread(model, sequences, trees, chrom);
}


extern "C" {

// sample_thread.cpp:981-1006
void arghmm_sample_arg_thread_internal(LocalTrees *trees,
    double *times, int ntimes, double *popsizes, double rho, double mu,
    char **seqs, int nseqs, int seqlen, int *thread_path)
{
    ArgModel model(ntimes, times, popsizes, rho, mu);
    Sequences sequences(seqs, nseqs, seqlen);
    device_thread(&model, &sequences, trees, -1, true, 0, NULL, -1,
                  thread_path);
}

// sample_thread.cpp:1010-1024
LocalTrees *arghmm_sample_thread(
    LocalTrees *trees, double *times, int ntimes,
    double *popsizes, double rho, double mu,
    char **seqs, int nseqs, int seqlen)
{
    ArgModel model(ntimes, times, popsizes, rho, mu);
    Sequences sequences(seqs, nseqs, seqlen);
    int new_chrom = nseqs - 1;
    sample_arg_thread(&model, &sequences, trees, new_chrom);
    return trees;
}

} // extern "C"

} // namespace argweaver
