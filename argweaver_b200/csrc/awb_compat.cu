// awb_compat.cu -- the reference's own extern "C" symbols for the threading-HMM
// path, with identical names and signatures (libargweaver.so as bound by
// argweaver/argweaverc.py:19-349), implemented over the flat awb_* ABI.
//
// What is here (reference file:line of the symbol it replaces):
//   arghmm_new_trees / delete_local_trees / get_local_trees_*  local_tree.cpp:1817-1900
//   arghmm_get_nstates / get_state_spaces / delete_state_spaces states.cpp:209-261
//   arghmm_forward_alg / arghmm_sample_posterior /
//   arghmm_sample_arg_thread_internal / delete_path /
//   delete_double_matrix / delete_forward_matrix               sample_thread.cpp:887-1042
//   new_emissions / delete_emissions                           emit.cpp:1288-1310
//   new_transition_probs(_switch) / delete_transition_probs    trans.cpp:1190-1251
//   forward_step / forward_alg / backward_alg /
//   sample_hmm_posterior / sample_hmm_posterior_step           hmm.cpp:13-98
//
// `LocalTrees` is an opaque handle owned by THIS library (a flat copy of the
// parent/age/SPR/blocklen arrays); ARG surgery (add/remove thread) is outside
// the replaced path and is not provided.  Every function computes on the GPU;
// host code only marshals arrays and draws libc rand() in the reference's order
// (one draw per sampled site, common.h:272-290) so the caller's rand() stream
// ends where the reference's would.  Errors follow the reference's convention
// (no error codes): a failure prints the message and aborts.

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "argweaver_b200.h"
#include "awb_common.cuh"

struct LocalTrees {
    int ntrees, nnodes, start_coord;
    std::vector<int> ptrees, ages, sprs, blocklens;
};

static void die(const char *where)
{
    fprintf(stderr, "argweaver_b200: %s: %s\n", where, awb_last_error());
    abort();
}

#define COMPAT_OK(call, where) do { if (call) die(where); } while (0)

static awb_ctx *compat_ctx()
{
    static awb_ctx *ctx = NULL;
    if (!ctx) {
        int dev = 0;
        const char *env = getenv("AWB_DEVICE");
        if (env) dev = atoi(env);
        if (awb_ctx_create(dev, &ctx))
            die("awb_ctx_create");
    }
    return ctx;
}

// ------------------------------------------------------------------ trees

extern "C" LocalTrees *arghmm_new_trees(int **ptrees, int **ages, int **sprs,
                                        int *blocklens, int ntrees, int nnodes,
                                        int start_coord)
{
    LocalTrees *t = new LocalTrees;
    t->ntrees = ntrees;
    t->nnodes = nnodes;
    t->start_coord = start_coord;
    for (int i = 0; i < ntrees; i++) {
        t->ptrees.insert(t->ptrees.end(), ptrees[i], ptrees[i] + nnodes);
        t->ages.insert(t->ages.end(), ages[i], ages[i] + nnodes);
        t->sprs.insert(t->sprs.end(), sprs[i], sprs[i] + 4);
        t->blocklens.push_back(blocklens[i]);
    }
    return t;
}

extern "C" void delete_local_trees(LocalTrees *trees) { delete trees; }
extern "C" int get_local_trees_ntrees(LocalTrees *trees) { return trees->ntrees; }
extern "C" int get_local_trees_nnodes(LocalTrees *trees) { return trees->nnodes; }

// A problem over host buffers assembled from reference-style arguments.
struct CompatProblem {
    awb_problem p;
    std::vector<unsigned char> seqs;
    std::vector<int> seqids;
    std::vector<double> times, popsizes;
    int nsites;
};

static void fill_problem(CompatProblem &cp, const LocalTrees *trees,
                         const double *times, int ntimes, const double *popsizes,
                         double rho, double mu, char **seqs, int nseqs,
                         int seqlen, bool internal)
{
    memset(&cp.p, 0, sizeof(cp.p));
    cp.times.assign(times, times + ntimes);
    if (popsizes)
        cp.popsizes.assign(popsizes, popsizes + ntimes);
    else
        cp.popsizes.assign(ntimes, 1e4);
    const int nleaves = (trees->nnodes + 1) / 2;
    int nsites = 0;
    for (int b = 0; b < trees->ntrees; b++)
        nsites += trees->blocklens[b];
    cp.nsites = nsites;
    if (seqs) {
        cp.seqs.resize((size_t) nseqs * seqlen);
        for (int i = 0; i < nseqs; i++)
            memcpy(&cp.seqs[(size_t) i * seqlen], seqs[i], seqlen);
    } else {
        // no sequence data needed (state / transition queries)
        nseqs = nleaves + 1;
        seqlen = trees->start_coord + nsites;
        cp.seqs.assign((size_t) nseqs * seqlen, 'A');
    }
    cp.seqids.resize(nleaves);
    for (int i = 0; i < nleaves; i++)
        cp.seqids[i] = i;                  // LocalTrees::set_default_seqids
    awb_problem &p = cp.p;
    p.ntimes = ntimes;
    p.times = cp.times.data();
    p.popsizes = cp.popsizes.data();
    p.rho = rho;
    p.mu = mu;
    p.nseqs = nseqs;
    p.seqlen = seqlen;
    p.seqs = cp.seqs.data();
    p.nleaves = nleaves;
    p.seqids = cp.seqids.data();
    p.new_chrom = internal ? -1 : nleaves;  // ArgHmmMatrixIter default (matrices.h:225)
    p.internal = internal ? 1 : 0;
    p.minage = 0;
    p.ntrees = trees->ntrees;
    p.nnodes = trees->nnodes;
    p.start_coord = trees->start_coord;
    p.ptrees = trees->ptrees.data();
    p.ages = trees->ages.data();
    p.sprs = trees->sprs.data();
    p.mappings = NULL;                      // make_node_mapping (local_tree.h:767)
    p.blocklens = trees->blocklens.data();
    p.subtree_roots = NULL;                 // child[0] of the root in index order
}

struct CompatRun {
    awb_batch *b;
    std::vector<int> nstates;
    std::vector<int64_t> row_off, fw_off, sw1_off;
    explicit CompatRun(const CompatProblem &cp, int flags) : b(NULL)
    {
        COMPAT_OK(awb_batch_create(compat_ctx(), 1, &cp.p, flags, &b),
                  "awb_batch_create");
        COMPAT_OK(awb_batch_upload(b), "awb_batch_upload");
        COMPAT_OK(awb_batch_setup(b), "awb_batch_setup");
        nstates.resize(cp.p.ntrees);
        row_off.resize(cp.p.ntrees + 1);
        fw_off.resize(cp.p.ntrees + 1);
        sw1_off.resize(cp.p.ntrees + 1);
        awb_batch_get_nstates(b, 0, nstates.data());
        awb_batch_get_layout(b, 0, row_off.data(), fw_off.data(), sw1_off.data());
    }
    ~CompatRun() { awb_batch_destroy(b); }
    template <typename T>
    std::vector<T> fetch(const char *name)
    {
        const int64_t nb = awb_batch_debug_bytes(b, 0, name);
        if (nb < 0) die(name);
        std::vector<T> v((size_t) nb / sizeof(T) + 1);
        COMPAT_OK(awb_batch_sync(b), "awb_batch_sync");
        COMPAT_OK(awb_batch_get_debug(b, 0, name, v.data(), nb), name);
        v.resize((size_t) nb / sizeof(T));
        return v;
    }
};

// ------------------------------------------------------------------ states

// states.cpp:209-225: number of states of the tree covering each site
extern "C" void arghmm_get_nstates(LocalTrees *trees, int ntimes, bool internal,
                                   int *nstates)
{
    std::vector<double> times(ntimes);
    for (int i = 0; i < ntimes; i++) times[i] = i;
    CompatProblem cp;
    fill_problem(cp, trees, times.data(), ntimes, NULL, 1e-8, 1e-8, NULL, 0, 0,
                 internal);
    awb_batch *b = NULL;
    COMPAT_OK(awb_batch_create(compat_ctx(), 1, &cp.p, 0, &b), "awb_batch_create");
    std::vector<int> ns(trees->ntrees);
    awb_batch_get_nstates(b, 0, ns.data());
    awb_batch_destroy(b);
    int i = 0;
    for (int t = 0; t < trees->ntrees; t++)
        for (int j = 0; j < trees->blocklens[t]; j++)
            nstates[i++] = ns[t];
}

// states.cpp:229-252 (quirk kept: `internal` is ignored, external states)
extern "C" intstate **get_state_spaces(LocalTrees *trees, int ntimes, bool internal)
{
    (void) internal;
    std::vector<double> times(ntimes);
    for (int i = 0; i < ntimes; i++) times[i] = i;
    CompatProblem cp;
    fill_problem(cp, trees, times.data(), ntimes, NULL, 1e-8, 1e-8, NULL, 0, 0,
                 false);
    CompatRun run(cp, 0);
    std::vector<short> node = run.fetch<short>("st_node");
    std::vector<signed char> time = run.fetch<signed char>("st_time");
    intstate **all = new intstate *[trees->ntrees];
    for (int t = 0; t < trees->ntrees; t++) {
        const int S = run.nstates[t];
        all[t] = new intstate[S > 0 ? S : 1];
        for (int j = 0; j < S; j++) {
            all[t][j][0] = node[run.row_off[t] + j];
            all[t][j][1] = time[run.row_off[t] + j];
        }
    }
    return all;
}

extern "C" void delete_state_spaces(intstate **all_states, int ntrees)
{
    for (int i = 0; i < ntrees; i++)
        delete[] all_states[i];
    delete[] all_states;
}

// ------------------------------------------------------------------ forward / sampling

// sample_thread.cpp:887-926.  Rows are allocated one by one
// (ArgHmmForwardTableOld) so delete_forward_matrix can free them.
extern "C" double **arghmm_forward_alg(LocalTrees *trees, double *times, int ntimes,
                                       double *popsizes, double rho, double mu,
                                       char **seqs, int nseqs, int seqlen,
                                       bool prior_given, double *prior,
                                       bool internal, bool slow)
{
    (void) slow;        // the dense "slow" path computes the same table
    CompatProblem cp;
    fill_problem(cp, trees, times, ntimes, popsizes, rho, mu, seqs, nseqs, seqlen,
                 internal);
    CompatRun run(cp, 0);
    const double *priors[1] = { prior_given ? prior : NULL };
    COMPAT_OK(awb_batch_forward(run.b, prior_given ? priors : NULL),
              "awb_batch_forward");
    COMPAT_OK(awb_batch_sync(run.b), "awb_batch_sync");
    std::vector<double> flat(awb_batch_fw_doubles(run.b, 0));
    COMPAT_OK(awb_batch_get_fw(run.b, 0, flat.data()), "awb_batch_get_fw");
    int bad = -1;
    awb_batch_get_status(run.b, 0, &bad);
    if (bad >= 0) {
        // sample_thread.cpp:443-444,457-458 assert(top > 0.0)
        fprintf(stderr, "argweaver_b200: forward column %d is not positive\n", bad);
        abort();
    }
    double **fw = new double *[cp.nsites];
    int site = 0;
    for (int t = 0; t < trees->ntrees; t++) {
        const int S1 = run.nstates[t] > 0 ? run.nstates[t] : 1;
        for (int i = 0; i < trees->blocklens[t]; i++, site++) {
            fw[site] = new double[S1];
            memcpy(fw[site], &flat[run.fw_off[t] + (int64_t) i * S1],
                   sizeof(double) * S1);
        }
    }
    return fw;
}

// draws in the reference's order: one rand() per sampled site, last site first
static void draw_rands(std::vector<int> &r, int n)
{
    r.resize(n);
    for (int i = 0; i < n; i++)
        r[i] = rand();
}

// sample_thread.cpp:930-977.  Deviation, on purpose: the reference's conversion
// loop re-declares `end` inside the loop body (:958-961), so with more than one
// local tree it rewrites path[0 .. blocklen) for every block and leaves the
// rest of the path unset; here every site is converted with its own block's
// states (identical to the reference for a single tree).
extern "C" intstate *arghmm_sample_posterior(int **ptrees, int **ages, int **sprs,
                                             int *blocklens, int ntrees, int nnodes,
                                             double *times, int ntimes,
                                             double *popsizes, double rho,
                                             double mu, char **seqs, int nseqs,
                                             int seqlen, intstate *path)
{
    LocalTrees *trees = arghmm_new_trees(ptrees, ages, sprs, blocklens, ntrees,
                                         nnodes, 0);
    CompatProblem cp;
    fill_problem(cp, trees, times, ntimes, popsizes, rho, mu, seqs, nseqs, seqlen,
                 false);
    CompatRun run(cp, 0);
    std::vector<int> r;
    draw_rands(r, cp.nsites);
    const int *rp[1] = { r.data() };
    COMPAT_OK(awb_batch_forward(run.b, NULL), "awb_batch_forward");
    COMPAT_OK(awb_batch_traceback(run.b, rp, RAND_MAX, NULL), "awb_batch_traceback");
    COMPAT_OK(awb_batch_sync(run.b), "awb_batch_sync");
    std::vector<int> ipath(cp.nsites);
    COMPAT_OK(awb_batch_get_path(run.b, 0, ipath.data()), "awb_batch_get_path");
    std::vector<short> node = run.fetch<short>("st_node");
    std::vector<signed char> time = run.fetch<signed char>("st_time");
    if (path == NULL)
        path = new intstate[seqlen];
    int site = 0;
    for (int t = 0; t < ntrees; t++) {
        for (int i = 0; i < blocklens[t]; i++, site++) {
            const int s = ipath[site];
            path[site][0] = node[run.row_off[t] + s];
            path[site][1] = time[run.row_off[t] + s];
        }
    }
    delete trees;
    return path;
}

// sample_thread.cpp:981-1006
extern "C" void arghmm_sample_arg_thread_internal(LocalTrees *trees, double *times,
                                                  int ntimes, double *popsizes,
                                                  double rho, double mu, char **seqs,
                                                  int nseqs, int seqlen,
                                                  int *thread_path)
{
    CompatProblem cp;
    fill_problem(cp, trees, times, ntimes, popsizes, rho, mu, seqs, nseqs, seqlen,
                 true);
    CompatRun run(cp, 0);
    std::vector<int> r;
    draw_rands(r, cp.nsites);
    const int *rp[1] = { r.data() };
    COMPAT_OK(awb_batch_forward(run.b, NULL), "awb_batch_forward");
    COMPAT_OK(awb_batch_traceback(run.b, rp, RAND_MAX, NULL), "awb_batch_traceback");
    COMPAT_OK(awb_batch_sync(run.b), "awb_batch_sync");
    COMPAT_OK(awb_batch_get_path(run.b, 0, thread_path), "awb_batch_get_path");
}

extern "C" void delete_path(int *path) { delete[] path; }

extern "C" void delete_double_matrix(double **mat, int nrows)
{
    (void) nrows;
    delete[] mat[0];
    delete[] mat;
}

extern "C" void delete_forward_matrix(double **mat, int nrows)
{
    for (int i = 0; i < nrows; i++)
        delete[] mat[i];
    delete[] mat;
}

static double **new_matrix(int nrows, int ncols)
{
    double **mat = new double *[nrows];
    double *block = new double[(size_t) nrows * (ncols > 0 ? ncols : 1)];
    for (int i = 0; i < nrows; i++)
        mat[i] = block + (size_t) i * ncols;
    return mat;
}

// map caller-supplied (node,time) states to this library's state indices
static std::vector<int> map_states(intstate *istates, int nstates,
                                   const std::vector<short> &node,
                                   const std::vector<signed char> &time,
                                   int64_t row0, int S)
{
    std::vector<int> idx(nstates, -1);
    for (int i = 0; i < nstates; i++)
        for (int j = 0; j < S; j++)
            if (node[row0 + j] == istates[i][0] && time[row0 + j] == istates[i][1]) {
                idx[i] = j;
                break;
            }
    return idx;
}

// ------------------------------------------------------------------ emissions

// emit.cpp:1288-1303: external-mode emissions of one tree
extern "C" double **new_emissions(intstate *istates, int nstates, int *ptree,
                                  int nnodes, int *ages_index, char **seqs,
                                  int nseqs, int seqlen, double *times, int ntimes,
                                  double mu)
{
    // one block of seqlen+1 sites whose first column duplicates site 0 (the
    // first column of a table carries the prior, not an emission)
    LocalTrees t;
    t.ntrees = 1;
    t.nnodes = nnodes;
    t.start_coord = 0;
    t.ptrees.assign(ptree, ptree + nnodes);
    t.ages.assign(ages_index, ages_index + nnodes);
    const int nospr[4] = { -1, -1, -1, -1 };
    t.sprs.assign(nospr, nospr + 4);
    t.blocklens.assign(1, seqlen + 1);
    std::vector<std::vector<char> > rows(nseqs, std::vector<char>(seqlen + 1));
    std::vector<char *> rowp(nseqs);
    for (int i = 0; i < nseqs; i++) {
        rows[i][0] = seqs[i][0];
        memcpy(&rows[i][1], seqs[i], seqlen);
        rowp[i] = rows[i].data();
    }
    CompatProblem cp;
    fill_problem(cp, &t, times, ntimes, NULL, 1e-8, mu, rowp.data(), nseqs,
                 seqlen + 1, false);
    CompatRun run(cp, 0);
    std::vector<unsigned char> kind = run.fetch<unsigned char>("kind");
    std::vector<double> inv = run.fetch<double>("inv_emit");
    std::vector<double> fw = run.fetch<double>("fw");
    std::vector<short> node = run.fetch<short>("st_node");
    std::vector<signed char> time = run.fetch<signed char>("st_time");
    const int S = run.nstates[0];
    std::vector<int> idx = map_states(istates, nstates, node, time, 0, S);
    double **emit = new_matrix(seqlen, nstates);
    for (int i = 0; i < seqlen; i++) {
        for (int j = 0; j < nstates; j++) {
            const int s = idx[j];
            double e = NAN;
            if (s >= 0) {
                if (kind[i + 1] == AWB_SITE_VARIANT)
                    e = fw[(size_t) (i + 1) * S + s];
                else if (kind[i + 1] == AWB_SITE_MASKED)
                    e = 1.0;
                else
                    e = inv[s];
            }
            emit[i][j] = e;
        }
    }
    return emit;
}

extern "C" void delete_emissions(double **emit, int seqlen)
{
    (void) seqlen;
    delete[] emit[0];
    delete[] emit;
}

// ------------------------------------------------------------------ transitions

__global__ void awb_dense_trans_kernel(const double *tv, int T, const short *node,
                                       const signed char *time, const int *ages,
                                       const int *idx, int n, double *out)
{
    const int i = blockIdx.x;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const int si = idx[i], sj = idx[j];
        double v = NAN;
        if (si >= 0 && sj >= 0)
            v = log(awb_get_time(tv, T, time[si], time[sj], ages[node[sj]], 0,
                                 node[si] == node[sj]));
        out[(size_t) i * n + j] = v;
    }
}

// trans.cpp:1190-1209: dense LOG transition matrix within a block (external).
// As in the reference, the passed lineage counts / treelen are ignored.
extern "C" double **new_transition_probs(int nnodes, int *ptree, int *ages,
                                         double treelen, intstate *istates,
                                         int nstates, int ntimes, double *times,
                                         double *time_steps, int *nbranches,
                                         int *nrecombs, int *ncoals,
                                         double *popsizes, double rho)
{
    (void) treelen; (void) time_steps; (void) nbranches; (void) nrecombs; (void) ncoals;
    LocalTrees t;
    t.ntrees = 1;
    t.nnodes = nnodes;
    t.start_coord = 0;
    t.ptrees.assign(ptree, ptree + nnodes);
    t.ages.assign(ages, ages + nnodes);
    const int nospr[4] = { -1, -1, -1, -1 };
    t.sprs.assign(nospr, nospr + 4);
    t.blocklens.assign(1, 1);
    CompatProblem cp;
    fill_problem(cp, &t, times, ntimes, popsizes, rho, 0.0, NULL, 0, 0, false);
    CompatRun run(cp, 0);
    std::vector<double> tv = run.fetch<double>("tmvec");
    std::vector<short> node = run.fetch<short>("st_node");
    std::vector<signed char> time = run.fetch<signed char>("st_time");
    const int S = run.nstates[0];
    std::vector<int> idx = map_states(istates, nstates, node, time, 0, S);

    double *d_tv, *d_out;
    short *d_node;
    signed char *d_time;
    int *d_ages, *d_idx;
    cudaMalloc(&d_tv, tv.size() * sizeof(double));
    cudaMalloc(&d_out, sizeof(double) * nstates * nstates + 8);
    cudaMalloc(&d_node, sizeof(short) * (S + 1));
    cudaMalloc(&d_time, S + 1);
    cudaMalloc(&d_ages, sizeof(int) * nnodes);
    cudaMalloc(&d_idx, sizeof(int) * (nstates + 1));
    cudaMemcpy(d_tv, tv.data(), tv.size() * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemcpy(d_node, node.data(), sizeof(short) * S, cudaMemcpyHostToDevice);
    cudaMemcpy(d_time, time.data(), S, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ages, ages, sizeof(int) * nnodes, cudaMemcpyHostToDevice);
    cudaMemcpy(d_idx, idx.data(), sizeof(int) * nstates, cudaMemcpyHostToDevice);
    if (nstates > 0)
        awb_dense_trans_kernel<<<nstates, 128>>>(d_tv, ntimes, d_node, d_time,
                                                 d_ages, d_idx, nstates, d_out);
    double **mat = new_matrix(nstates, nstates);
    if (cudaMemcpy(mat[0], d_out, sizeof(double) * nstates * nstates,
                   cudaMemcpyDeviceToHost) != cudaSuccess) {
        fprintf(stderr, "argweaver_b200: new_transition_probs: CUDA failure\n");
        abort();
    }
    cudaFree(d_tv); cudaFree(d_out); cudaFree(d_node); cudaFree(d_time);
    cudaFree(d_ages); cudaFree(d_idx);
    return mat;
}

// trans.cpp:1212-1245: dense LOG switch matrix between two consecutive trees
extern "C" double **new_transition_probs_switch(
    int *ptree, int *last_ptree, int nnodes, int recomb_node, int recomb_time,
    int coal_node, int coal_time, int *ages_index, int *last_ages_index,
    double treelen, double last_treelen, intstate *istates1, int nstates1,
    intstate *istates2, int nstates2, int ntimes, double *times,
    double *time_steps, int *nbranches, int *nrecombs, int *ncoals,
    double *popsizes, double rho)
{
    (void) treelen; (void) last_treelen; (void) time_steps; (void) nbranches;
    (void) nrecombs; (void) ncoals;
    LocalTrees t;
    t.ntrees = 2;
    t.nnodes = nnodes;
    t.start_coord = 0;
    t.ptrees.assign(last_ptree, last_ptree + nnodes);
    t.ptrees.insert(t.ptrees.end(), ptree, ptree + nnodes);
    t.ages.assign(last_ages_index, last_ages_index + nnodes);
    t.ages.insert(t.ages.end(), ages_index, ages_index + nnodes);
    const int sprs[8] = { -1, -1, -1, -1, recomb_node, recomb_time, coal_node,
                          coal_time };
    t.sprs.assign(sprs, sprs + 8);
    t.blocklens.assign(2, 1);
    CompatProblem cp;
    fill_problem(cp, &t, times, ntimes, popsizes, rho, 0.0, NULL, 0, 0, false);
    CompatRun run(cp, AWB_KEEP_DEBUG);
    std::vector<short> node = run.fetch<short>("st_node");
    std::vector<signed char> time = run.fetch<signed char>("st_time");
    std::vector<int> determ = run.fetch<int>("sw_determ");
    std::vector<double> dprob = run.fetch<double>("sw_determprob");
    std::vector<double> rrow = run.fetch<double>("sw_recombrow");
    std::vector<double> crow = run.fetch<double>("sw_recoalrow");
    std::vector<int> rsrc = run.fetch<int>("sw_recombsrc");
    std::vector<int> csrc = run.fetch<int>("sw_recoalsrc");
    const int S1 = run.nstates[0], S2 = run.nstates[1];
    std::vector<int> idx1 = map_states(istates1, nstates1, node, time, run.row_off[0], S1);
    std::vector<int> idx2 = map_states(istates2, nstates2, node, time, run.row_off[1], S2);
    double **mat = new_matrix(nstates1, nstates2);
    for (int i = 0; i < nstates1; i++) {
        for (int j = 0; j < nstates2; j++) {
            const int a = idx1[i], c = idx2[j];
            double v = NAN;
            if (a >= 0 && c >= 0) {
                // TransMatrixSwitch::get (trans.h:207-219)
                double pr;
                if (a == csrc[1]) pr = crow[run.row_off[1] + c];
                else if (a == rsrc[1]) pr = rrow[run.row_off[1] + c];
                else pr = (determ[run.sw1_off[1] + a] == c) ? dprob[run.sw1_off[1] + a] : 0.0;
                v = log(pr);
            }
            mat[i][j] = v;
        }
    }
    return mat;
}

extern "C" void delete_transition_probs(double **transmat, int nstates)
{
    (void) nstates;
    delete[] transmat[0];
    delete[] transmat;
}

// ------------------------------------------------------------------ generic dense HMM

// common.h:182-202 (terms more than 15 below the max are dropped)
__device__ inline double awb_logsum_block(double v, bool valid, double *red)
{
    // block-wide max then block-wide sum; red has >= 33 doubles
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    double m = valid ? v : -INFINITY;
    for (int d = 16; d >= 1; d >>= 1)
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    double mx = -INFINITY;
    for (int w = 0; w < nwarps; w++) mx = fmax(mx, red[w]);
    __syncthreads();
    double s = (valid && (v - mx > -15.0)) ? exp(v - mx) : 0.0;
    for (int d = 16; d >= 1; d >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < nwarps; w++) tot += red[w];
    __syncthreads();
    return mx + log(tot);
}

// one CTA per target state k: col2[k] = logsum_j(col1[j] + trans[j][k]) + emit[k]
__global__ void awb_hmm_step_kernel(const double *col1, double *col2, int n1, int n2,
                                    const double *trans /*[n1][n2]*/,
                                    const double *emit, int transpose)
{
    __shared__ double red[33];
    const int k = blockIdx.x;
    // the threshold rule needs the global max first (common.h:182-202)
    double m = -INFINITY;
    for (int j = threadIdx.x; j < n1; j += blockDim.x) {
        const double t = transpose ? trans[(size_t) k * n1 + j] : trans[(size_t) j * n2 + k];
        m = fmax(m, col1[j] + t);
    }
    for (int d = 16; d >= 1; d >>= 1)
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    double mx = -INFINITY;
    for (int w = 0; w < (blockDim.x + 31) / 32; w++) mx = fmax(mx, red[w]);
    __syncthreads();
    double s = 0.0;
    for (int j = threadIdx.x; j < n1; j += blockDim.x) {
        const double t = transpose ? trans[(size_t) k * n1 + j] : trans[(size_t) j * n2 + k];
        const double v = col1[j] + t;
        if (v - mx > -15.0) s += exp(v - mx);
    }
    for (int d = 16; d >= 1; d >>= 1)
        s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (blockDim.x + 31) / 32; w++) tot += red[w];
        col2[k] = mx + log(tot) + (emit ? emit[k] : 0.0);
    }
}

struct DevBuf {
    double *p;
    explicit DevBuf(size_t n) : p(NULL) { cudaMalloc(&p, sizeof(double) * (n ? n : 1)); }
    ~DevBuf() { cudaFree(p); }
};

static void flatten(double **m, int nrows, int ncols, std::vector<double> &out)
{
    out.resize((size_t) nrows * ncols);
    for (int i = 0; i < nrows; i++)
        memcpy(&out[(size_t) i * ncols], m[i], sizeof(double) * ncols);
}

static void cuda_check(const char *where)
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        fprintf(stderr, "argweaver_b200: %s: %s\n", where, cudaGetErrorString(e));
        abort();
    }
}

// hmm.cpp:13-24
extern "C" void forward_step(double *col1, double *col2, int nstates1, int nstates2,
                             double **trans, double *emit)
{
    compat_ctx();
    std::vector<double> tr;
    flatten(trans, nstates1, nstates2, tr);
    DevBuf d_tr(tr.size()), d_c1(nstates1), d_c2(nstates2), d_em(nstates2);
    cudaMemcpy(d_tr.p, tr.data(), sizeof(double) * tr.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_c1.p, col1, sizeof(double) * nstates1, cudaMemcpyHostToDevice);
    cudaMemcpy(d_em.p, emit, sizeof(double) * nstates2, cudaMemcpyHostToDevice);
    awb_hmm_step_kernel<<<nstates2, 128>>>(d_c1.p, d_c2.p, nstates1, nstates2,
                                           d_tr.p, d_em.p, 0);
    cuda_check("forward_step");
    cudaMemcpy(col2, d_c2.p, sizeof(double) * nstates2, cudaMemcpyDeviceToHost);
}

// hmm.cpp:27-43 (first column of fw must already be filled)
extern "C" void forward_alg(int n, int nstates, double **trans, double **emit,
                            double **fw)
{
    compat_ctx();
    std::vector<double> tr, em;
    flatten(trans, nstates, nstates, tr);
    flatten(emit, n, nstates, em);
    DevBuf d_tr(tr.size()), d_em(em.size()), d_fw((size_t) n * nstates);
    cudaMemcpy(d_tr.p, tr.data(), sizeof(double) * tr.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_em.p, em.data(), sizeof(double) * em.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_fw.p, fw[0], sizeof(double) * nstates, cudaMemcpyHostToDevice);
    for (int i = 1; i < n; i++)
        awb_hmm_step_kernel<<<nstates, 128>>>(
            d_fw.p + (size_t) (i - 1) * nstates, d_fw.p + (size_t) i * nstates,
            nstates, nstates, d_tr.p, d_em.p + (size_t) i * nstates, 0);
    cuda_check("forward_alg");
    std::vector<double> out((size_t) n * nstates);
    cudaMemcpy(out.data(), d_fw.p, sizeof(double) * out.size(), cudaMemcpyDeviceToHost);
    for (int i = 1; i < n; i++)
        memcpy(fw[i], &out[(size_t) i * nstates], sizeof(double) * nstates);
}

// hmm.cpp:48-64 (last column of bw must already be filled)
extern "C" void backward_alg(int n, int nstates, double **trans, double **emit,
                             double **bw)
{
    compat_ctx();
    std::vector<double> tr, em;
    flatten(trans, nstates, nstates, tr);
    flatten(emit, n, nstates, em);
    DevBuf d_tr(tr.size()), d_bw((size_t) n * nstates), d_tmp(nstates);
    cudaMemcpy(d_tr.p, tr.data(), sizeof(double) * tr.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_bw.p + (size_t) (n - 1) * nstates, bw[n - 1], sizeof(double) * nstates,
               cudaMemcpyHostToDevice);
    std::vector<double> col(nstates);
    for (int i = n - 2; i >= 0; i--) {
        // col2[k] + emit[i+1][k], then bw[i][j] = logsum_k(trans[j][k] + that)
        cudaMemcpy(col.data(), d_bw.p + (size_t) (i + 1) * nstates,
                   sizeof(double) * nstates, cudaMemcpyDeviceToHost);
        for (int k = 0; k < nstates; k++)
            col[k] += em[(size_t) (i + 1) * nstates + k];
        cudaMemcpy(d_tmp.p, col.data(), sizeof(double) * nstates, cudaMemcpyHostToDevice);
        awb_hmm_step_kernel<<<nstates, 128>>>(d_tmp.p, d_bw.p + (size_t) i * nstates,
                                              nstates, nstates, d_tr.p, NULL, 1);
    }
    cuda_check("backward_alg");
    std::vector<double> out((size_t) n * nstates);
    cudaMemcpy(out.data(), d_bw.p, sizeof(double) * out.size(), cudaMemcpyDeviceToHost);
    for (int i = 0; i < n - 1; i++)
        memcpy(bw[i], &out[(size_t) i * nstates], sizeof(double) * nstates);
}

// posterior weights of one step on the device: A[j] = exp(col[j]+trans[j][k]-logsum)
__global__ void awb_hmm_post_kernel(const double *col, const double *trans, int n,
                                    int k, double *A)
{
    __shared__ double red[33];
    const int j = threadIdx.x;
    const bool valid = j < n;
    const double v = valid ? col[j] + trans[(size_t) j * n + k] : -INFINITY;
    const double tot = awb_logsum_block(v, valid, red);
    if (valid)
        A[j] = exp(v - tot);
}

static int sample_host(const std::vector<double> &w)
{
    // common.h:272-290
    double total = 0.0;
    for (size_t i = 0; i < w.size(); i++) total += w[i];
    const double pick = rand() / double(RAND_MAX) * total;
    double x = 0.0;
    for (size_t i = 0; i < w.size(); i++) {
        x += w[i];
        if (x >= pick) return (int) i;
    }
    return (int) w.size() - 1;
}

// hmm.cpp:87-98
extern "C" int sample_hmm_posterior_step(int nstates1, double **trans, double *col1,
                                         int state2)
{
    compat_ctx();
    if (nstates1 > 1024) {
        fprintf(stderr, "argweaver_b200: sample_hmm_posterior_step: > 1024 states\n");
        abort();
    }
    std::vector<double> tr;
    flatten(trans, nstates1, nstates1, tr);
    DevBuf d_tr(tr.size()), d_c(nstates1), d_A(nstates1);
    cudaMemcpy(d_tr.p, tr.data(), sizeof(double) * tr.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_c.p, col1, sizeof(double) * nstates1, cudaMemcpyHostToDevice);
    awb_hmm_post_kernel<<<1, ((nstates1 + 31) / 32) * 32>>>(d_c.p, d_tr.p, nstates1,
                                                            state2, d_A.p);
    cuda_check("sample_hmm_posterior_step");
    std::vector<double> A(nstates1);
    cudaMemcpy(A.data(), d_A.p, sizeof(double) * nstates1, cudaMemcpyDeviceToHost);
    return sample_host(A);
}

// hmm.cpp:67-84 (path[n-1] must already be sampled)
extern "C" void sample_hmm_posterior(int n, int nstates, double **trans, double **fw,
                                     int *path)
{
    compat_ctx();
    if (nstates > 1024) {
        fprintf(stderr, "argweaver_b200: sample_hmm_posterior: > 1024 states\n");
        abort();
    }
    std::vector<double> tr, f;
    flatten(trans, nstates, nstates, tr);
    flatten(fw, n, nstates, f);
    DevBuf d_tr(tr.size()), d_fw(f.size()), d_A(nstates);
    cudaMemcpy(d_tr.p, tr.data(), sizeof(double) * tr.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_fw.p, f.data(), sizeof(double) * f.size(), cudaMemcpyHostToDevice);
    std::vector<double> A(nstates);
    for (int i = n - 2; i >= 0; i--) {
        awb_hmm_post_kernel<<<1, ((nstates + 31) / 32) * 32>>>(
            d_fw.p + (size_t) i * nstates, d_tr.p, nstates, path[i + 1], d_A.p);
        cudaMemcpy(A.data(), d_A.p, sizeof(double) * nstates, cudaMemcpyDeviceToHost);
        path[i] = sample_host(A);
    }
    cuda_check("sample_hmm_posterior");
}

// ------------------------------------------------------------------ total_prob.cpp

// total_prob.cpp:316-326
extern "C" double arghmm_likelihood(LocalTrees *trees, double *times, int ntimes,
                                    double mu, char **seqs, int nseqs, int seqlen)
{
    CompatProblem cp;
    fill_problem(cp, trees, times, ntimes, NULL, 0.0, mu, seqs, nseqs, seqlen, false);
    double lnl = 0.0;
    COMPAT_OK(awb_arg_likelihood(&cp.p, &lnl), "awb_arg_likelihood");
    return lnl;
}

// total_prob.cpp:345-352
extern "C" double arghmm_prior_prob(LocalTrees *trees, double *times, int ntimes,
                                    double *popsizes, double rho)
{
    CompatProblem cp;
    fill_problem(cp, trees, times, ntimes, popsizes, rho, 0.0, NULL, 0, 0, false);
    double lnl = 0.0;
    COMPAT_OK(awb_arg_prior(&cp.p, &lnl), "awb_arg_prior");
    return lnl;
}

// total_prob.cpp:369-377
extern "C" double arghmm_joint_prob(LocalTrees *trees, double *times, int ntimes,
                                    double *popsizes, double mu, double rho,
                                    char **seqs, int nseqs, int seqlen)
{
    CompatProblem cp;
    fill_problem(cp, trees, times, ntimes, popsizes, rho, mu, seqs, nseqs, seqlen, false);
    double lik = 0.0, prior = 0.0;
    COMPAT_OK(awb_arg_joint(&cp.p, &lik, &prior), "awb_arg_joint");
    return lik + prior;
}
