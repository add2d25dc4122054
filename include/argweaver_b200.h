/* argweaver_b200.h -- C ABI of the B200-native threading-HMM path.
 *
 * This library replaces ONE path of ARGweaver (mdrasmus/argweaver): the
 * threading HMM that arg-sample runs for every MCMC step --
 *   emissions            src/argweaver/emit.cpp:650-869      (calc_emissions)
 *   transition setup     src/argweaver/trans.cpp:26-115      (calc_transition_probs)
 *                        src/argweaver/trans.cpp:538-739     (calc_transition_probs_switch)
 *   forward recursion    src/argweaver/sample_thread.cpp:186-296,345-460
 *   stochastic traceback src/argweaver/sample_thread.cpp:470-569
 * All compute runs in hand-written CUDA kernels for sm_100a; there is no CPU
 * fallback: every entry point fails (non-zero return + awb_last_error()) when
 * no CUDA device is usable.
 *
 * Two layers are exported:
 *
 *  (1) the flat ABI (awb_*): plain pointers and sizes, no C++ types.  It is what
 *      the reference's L2 wrappers (sample_arg_thread, sample_arg_thread_internal,
 *      cond_sample_arg_thread*, sample_thread.cpp:578-865) would call after
 *      flattening LocalTrees / ArgModel / Sequences; INTEGRATION.md shows that
 *      adapter.
 *
 *  (2) the reference's own extern "C" symbols for this path, with identical
 *      signatures (argweaver/argweaverc.py:19-349 binds them through ctypes);
 *      declared at the end of this header.
 */
#ifndef ARGWEAVER_B200_H
#define ARGWEAVER_B200_H

#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define AWB_MAX_NTIMES 64     /* model time points supported by this build */
#define AWB_MAX_NNODES 1024   /* nodes per local tree (k <= 512 sequences)  */
#define AWB_MAX_NSTATES 2048  /* HMM states per block                        */

/* One thread-sampling problem ("chain"): the flattened arguments of
 * arghmm_forward_alg(trees, model, sequences, matrix_iter, ...) and
 * stochastic_traceback(...) (sample_thread.h:130-138).                     */
typedef struct awb_problem {
    /* ArgModel (model.h:41-354): time grid, population sizes, rates already
     * multiplied by the site-compression factor (arg-sample.cpp:428-436)     */
    int ntimes;
    const double *times;        /* [ntimes] */
    const double *popsizes;     /* [ntimes] */
    double rho, mu;

    /* Sequences (sequences.h:31): dense char rows, borrowed, not NUL-terminated */
    int nseqs, seqlen;
    const unsigned char *seqs;  /* [nseqs][seqlen] */
    int nleaves;
    const int *seqids;          /* [nleaves] LocalTrees::seqids              */
    int new_chrom;              /* external mode: row being threaded          */

    /* StatesModel (states.h:156): mode */
    int internal;               /* 0: thread a new leaf; 1: re-thread a subtree */
    int minage;                 /* always 0 at live call sites                */

    /* LocalTrees (local_tree.h:504): one entry per local block               */
    int ntrees, nnodes;
    int start_coord;            /* first site (index into seqs rows)          */
    const int *ptrees;          /* [ntrees][nnodes] parent, -1 = root         */
    const int *ages;            /* [ntrees][nnodes] time index                */
    const int *sprs;            /* [ntrees][4] recomb_node, recomb_time, coal_node, coal_time */
    const int *mappings;        /* [ntrees][nnodes] or NULL (= identity except broken node) */
    const int *blocklens;       /* [ntrees] */
    const int *subtree_roots;   /* [ntrees] child[0] of the root (internal) or NULL */

    /* Optional: the alignment as its variant columns only, the form a .sites
     * file holds (Sites, sequences.h:211-268) -- every other column of every row
     * is `default_char`, as make_sequences_from_sites fills it
     * (sequences.cpp:323-352).  When var_cols != NULL it REPLACES seqs (which
     * may be NULL): ~3 % of the bytes go to the device.  Positions are column
     * indices of the (imaginary) dense rows, ascending and unique. */
    int nvar;
    const int *var_pos;            /* [nvar] */
    const unsigned char *var_cols; /* [nvar][nseqs] */
    unsigned char default_char;    /* 0: 'A' */

    /* Optional emission mode (ArgModel::infsites_penalty, model.h:348): states
     * that would need a second mutation at a site have their emission multiplied
     * by this penalty (emit.cpp:457-589, :848-862).  0 or >= 1: off. */
    double infsites_penalty;

    /* Optional emission mode: unphased data (ArgModel::unphased, emit.cpp:705-742,
     * :834-842).  phase_row1 / phase_row2 are the two haplotypes of one
     * individual as positions in the leaf order (PhaseProbs::treemap1/2,
     * sequences.h:174-206: 0..nleaves-1, or nleaves for the new chromosome in
     * external mode); at the sites where they differ the emission is the mean
     * over the two phasings.  awb_batch_phase_probs gives, after the traceback,
     * the probability of the data's phasing at the sampled state of each such
     * site -- what PhaseProbs::sample_phase draws against.  Not combined with
     * infsites_penalty. */
    int unphased;               /* 0: off */
    int phase_row1, phase_row2;
} awb_problem;

typedef struct awb_ctx awb_ctx;       /* one CUDA device + stream             */
typedef struct awb_batch awb_batch;   /* a set of independent problems        */

/* flags for awb_batch_create */
#define AWB_KEEP_DEBUG 1   /* keep per-block setup arrays readable (tests)   */
/* Do not keep the forward table: the window is cut into segments, the forward
 * pass stores the first column of every segment, and the traceback rebuilds one
 * segment's table at a time (the forward recursion runs twice).  Memory per
 * problem drops from 8 B per site*state to one segment, so several times more
 * problems fit a batch.  awb_batch_get_fw is not available in this mode. */
#define AWB_CHECKPOINT 2

const char *awb_last_error(void);
int awb_device_count(void);

int awb_ctx_create(int device, awb_ctx **out);
void awb_ctx_destroy(awb_ctx *ctx);
/* CUDA events on the context's stream (8 slots) for timing by the caller */
int awb_ctx_record(awb_ctx *ctx, int slot);
int awb_ctx_elapsed_ms(awb_ctx *ctx, int slot0, int slot1, float *ms);
int awb_ctx_sync(awb_ctx *ctx);

/* Validate the problems and compute the table layout (host-only work; device
 * memory is taken at the first awb_batch_upload).  The problem structs are
 * copied; the ARRAYS they point to are read by asynchronous copies
 * (cudaMemcpyAsync straight from the caller's memory, truly asynchronous when
 * it is pinned) queued by awb_batch_upload, so they must stay valid and
 * unchanged until awb_batch_sync() -- or any call that returns results -- has
 * returned.  The same holds for the rand() draws given to
 * awb_batch_upload_rand / awb_batch_traceback and the priors given to
 * awb_batch_forward. */
int awb_batch_create(awb_ctx *ctx, int nproblems, const awb_problem *problems,
                     int flags, awb_batch **out);
void awb_batch_destroy(awb_batch *b);

/* host -> device copy of every problem's inputs (trees, SPRs, sequences) */
int awb_batch_upload(awb_batch *b);
int64_t awb_batch_h2d_bytes(const awb_batch *b);

/* Per-block setup (states, lineage counts, transition vectors, switch
 * matrices, site classification, variant-site emissions), then the forward
 * recursion for every problem of the batch.  Asynchronous on the ctx stream.
 * priors: NULL, or per problem a host pointer (or NULL) to a caller-supplied
 * first column (prior_given, sample_thread.cpp:416-429). */
int awb_batch_setup(awb_batch *b);
int awb_batch_forward(awb_batch *b, const double *const *priors);

/* Stochastic traceback.  rand_ints[i] points to the libc rand() draws for
 * problem i in consumption order (last site first; common.h:272-290), one per
 * site; rand_max is RAND_MAX.  last_states: NULL or per problem the given last
 * state (-1 = sample it). */
int awb_batch_traceback(awb_batch *b, const int *const *rand_ints, int rand_max,
                        const int *last_states);

/* Upload the rand() draws ahead of time; awb_batch_traceback may then be
 * called with rand_ints == NULL. */
int awb_batch_upload_rand(awb_batch *b, const int *const *rand_ints);

int awb_batch_sync(awb_batch *b);

/* milliseconds spent in the most recent setup / forward / traceback launches,
 * measured with CUDA events on the launching stream */
int awb_batch_timings(awb_batch *b, float *setup_ms, float *forward_ms,
                      float *traceback_ms);
/* sum over blocks of blocklen * nstates for problem i */
double awb_batch_states_sites(const awb_batch *b, int i);
int64_t awb_batch_fw_doubles(const awb_batch *b, int i);
int awb_batch_nsites(const awb_batch *b, int i);
int awb_batch_kernel_launches(const awb_batch *b);
/* AWB_CHECKPOINT: segments of the longest window, and segment tables kept per
 * window (chosen at the first upload from the free device memory; the last
 * that many segments of the forward pass are not rebuilt for the traceback).
 * Both 1 without AWB_CHECKPOINT. */
int awb_batch_segments(const awb_batch *b);
/* which forward kernel the batch's shape selects: 1 = the register-resident
 * fast kernel (awb_forward_fast.cuh), 0 = the generic one (awb_forward.cuh) */
int awb_batch_forward_kernel(const awb_batch *b);
int awb_batch_resident_segments(const awb_batch *b);

/* results (device -> host) */
int awb_batch_get_path(awb_batch *b, int i, int *path /*[nsites]*/);
int awb_batch_get_logz(awb_batch *b, int i, double *logz);
int awb_batch_get_status(awb_batch *b, int i, int *first_bad_site);
int awb_batch_get_fw(awb_batch *b, int i, double *fw /*[fw_doubles]*/);
int awb_batch_get_nstates(awb_batch *b, int i, int *nstates /*[ntrees]*/);
int awb_batch_get_layout(awb_batch *b, int i, int64_t *row_off, int64_t *fw_off,
                         int64_t *sw1_off /* each [ntrees+1] */);
/* AWB_KEEP_DEBUG only: named per-block arrays, see awb_api.cu (debug_fetch) */
int awb_batch_get_debug(awb_batch *b, int i, const char *name, void *dst,
                        int64_t dst_bytes);
int64_t awb_batch_debug_bytes(awb_batch *b, int i, const char *name);

/* One-shot convenience over host buffers: create + upload + setup + forward +
 * traceback + download, for a single problem. */
int awb_thread_sample(const awb_problem *p, const int *rand_ints, int rand_max,
                      int *path, double *logz);
/* the same with a caller-supplied first column (prior_given; prior == NULL:
 * the model's prior) and a given last state (last_state_given; -1: sample it;
 * the draws then start with the second-to-last site, n-1 of them) -- what
 * cond_sample_arg_thread[_internal] pass (sample_thread.cpp:700-865).  Fails
 * when a forward column has no positive entry (the reference asserts). */
int awb_thread_sample_cond(const awb_problem *p, const double *prior,
                           int last_state, const int *rand_ints, int rand_max,
                           int *path, double *logz);
int awb_forward_table(const awb_problem *p, const double *prior, double *fw,
                      double *logz);

/* Unphased data: P(phasing as given | sampled state) for the heterozygous sites of
 * problem i's unphased individual, after awb_batch_traceback.  p[nsites]: -1 at
 * the other sites. */
int awb_batch_phase_probs(awb_batch *b);
int awb_batch_get_phase_probs(awb_batch *b, int i, double *p);

/* Device time per kernel class, measured with CUDA events around every launch on
 * the batch's stream (bench.py's roofline figure).  awb_batch_kernel_times(b, 1)
 * starts (and resets) the collection; awb_batch_get_kernel_times sums what has
 * been launched since the start or the previous read: ms[AWB_KERNEL_CLASSES] in the order kind, block setup,
 * time matrices, switch setup, emission, forward, traceback, recombination;
 * forward_bytes = algorithmic bytes of the forward launches (8 B per site*state
 * they computed). */
#define AWB_KERNEL_CLASSES 8
int awb_batch_kernel_times(awb_batch *b, int enable);
int awb_batch_get_kernel_times(awb_batch *b, float *ms, double *forward_bytes,
                               int *forward_launches);

/* Recombination points of the sampled thread, the step right after the
 * traceback (sample_recombinations, recomb.cpp:151-235, with
 * recomb_prob_unnormalized :14-117 and get_possible_recomb :122-141), on the
 * device: path and per-block tables never leave it.  The reference draws a
 * data-dependent number of values from libc rand() here, so the kernel runs
 * glibc's own generator (TYPE_3 additive feedback, random_r.c) from a snapshot
 * of the caller's state -- AWB_RNG_WORDS ints per window: r[0..30], front index,
 * rear index, type -- and reports how many draws it took.
 *   awb_libc_rand_snapshot   the calling process's rand() state (glibc only)
 *   awb_libc_rand_advance    consume that many rand() values
 *   awb_rng_draw             the same generator on the host (advances `state`)
 * Positions are site indices of the window (0-based); node -1 is the new leaf
 * (external mode, recomb.cpp:146). */
#define AWB_RNG_WORDS 34
int awb_libc_rand_snapshot(int *state /* [AWB_RNG_WORDS] */);
void awb_libc_rand_advance(long long ndraws);
int awb_rng_draw(int *state, int n, int *out /* [n] or NULL */);
int awb_batch_sample_recombs(awb_batch *b, const int *rng_states /* [n][AWB_RNG_WORDS] */,
                             int rand_max);
int awb_batch_get_recomb_count(awb_batch *b, int i, int *nrecombs, int *draws);
int awb_batch_get_recombs(awb_batch *b, int i, int count, int *pos, int *node,
                          int *time);
/* one problem: traceback + recombination points.  rng_state NULL: snapshot the
 * process's libc stream (taken after the caller has drawn rand_ints from it)
 * and advance it by the draws used, which leaves it where the reference's
 * sample_arg_thread would (sample_thread.cpp:578-632). */
int awb_thread_sample_recombs(const awb_problem *p, const double *prior,
                              int last_state, const int *rand_ints, int rand_max,
                              const int *rng_state, int *path, double *logz,
                              int cap, int *nrecombs, int *pos, int *node,
                              int *time, int *draws);

/* .sites ingest and site compression, the step in front of the path
 * (arg-sample.cpp:965-1007): read_sites (sequences.cpp:173-303),
 * find_compress_cols + compress_sites (:523-609), make_sequences_from_sites
 * (:323-352).  Host-only.  Positions are 0-based; columns are [ncols][nseqs]
 * upper-case characters.  awb_sites_positions / awb_sites_columns can be passed
 * on as awb_problem.var_pos / var_cols (with nseqs, seqlen = end - start). */
typedef struct awb_sites awb_sites;
int awb_sites_read(const char *filename, int subregion_start /* -1: none */,
                   int subregion_end, awb_sites **out);
int awb_sites_from_columns(int nseqs, int start_coord, int end_coord, int ncols,
                           const int *positions, const unsigned char *cols,
                           awb_sites **out);
void awb_sites_free(awb_sites *s);
int awb_sites_nseqs(const awb_sites *s);
int awb_sites_ncols(const awb_sites *s);
int awb_sites_start(const awb_sites *s);
int awb_sites_end(const awb_sites *s);
const char *awb_sites_name(const awb_sites *s, int i);
const int *awb_sites_positions(const awb_sites *s);
const unsigned char *awb_sites_columns(const awb_sites *s);
/* -c / --compress-seq: 0 = done, 2 = cannot be compressed at this level (the
 * sites are left as they were), 1 = error */
int awb_sites_compress(awb_sites *s, int compress);
/* SitesMapping::all_sites after awb_sites_compress (compressed -> old coordinate) */
int awb_sites_mapping_size(const awb_sites *s);
const int *awb_sites_mapping(const awb_sites *s);
int awb_sites_to_sequences(const awb_sites *s, unsigned char *seqs /*[nseqs][end-start]*/,
                           unsigned char default_char /* 0: 'A' */);

/* Log-likelihood and log-prior of a COMPLETE ARG (the `likelihood` and `prior`
 * columns of arg-sample's .stats file, arg-sample.cpp:444-488):
 * calc_arg_likelihood (total_prob.cpp:19-42) and calc_arg_prior (:262-299).
 * `arg` uses the tree, model and sequence fields of awb_problem (ntrees trees over
 * nleaves = (nnodes+1)/2 sequences, leaf j reading row seqids[j]); the threading
 * fields (new_chrom, internal, minage, mappings, subtree_roots) are ignored. */
int awb_arg_likelihood(const awb_problem *arg, double *lnl);
int awb_arg_prior(const awb_problem *arg, double *lnl);
int awb_arg_joint(const awb_problem *arg, double *likelihood, double *prior);

/* ------------------------------------------------------------------------
 * Reference-compatible symbols (same names / signatures as libargweaver.so).
 * `LocalTrees` is an opaque handle owned by this library.
 * ---------------------------------------------------------------------- */
typedef struct LocalTrees LocalTrees;
typedef int intstate[2];

/* local_tree.cpp:1817-1825, :1895 */
LocalTrees *arghmm_new_trees(int **ptrees, int **ages, int **sprs,
                             int *blocklens, int ntrees, int nnodes,
                             int start_coord);
void delete_local_trees(LocalTrees *trees);
int get_local_trees_ntrees(LocalTrees *trees);
int get_local_trees_nnodes(LocalTrees *trees);

/* states.cpp:209-261 */
void arghmm_get_nstates(LocalTrees *trees, int ntimes, bool internal,
                        int *nstates);
intstate **get_state_spaces(LocalTrees *trees, int ntimes, bool internal);
void delete_state_spaces(intstate **all_states, int ntrees);

/* sample_thread.cpp:887-1042 */
double **arghmm_forward_alg(LocalTrees *trees, double *times, int ntimes,
                            double *popsizes, double rho, double mu,
                            char **seqs, int nseqs, int seqlen,
                            bool prior_given, double *prior, bool internal,
                            bool slow);
intstate *arghmm_sample_posterior(int **ptrees, int **ages, int **sprs,
                                  int *blocklens, int ntrees, int nnodes,
                                  double *times, int ntimes, double *popsizes,
                                  double rho, double mu, char **seqs, int nseqs,
                                  int seqlen, intstate *path);
void arghmm_sample_arg_thread_internal(LocalTrees *trees, double *times,
                                       int ntimes, double *popsizes, double rho,
                                       double mu, char **seqs, int nseqs,
                                       int seqlen, int *thread_path);
void delete_path(int *path);
void delete_double_matrix(double **mat, int nrows);
void delete_forward_matrix(double **mat, int nrows);

/* emit.cpp:1288-1310 */
double **new_emissions(intstate *istates, int nstates, int *ptree, int nnodes,
                       int *ages_index, char **seqs, int nseqs, int seqlen,
                       double *times, int ntimes, double mu);
void delete_emissions(double **emit, int seqlen);

/* trans.cpp:1190-1251 (both return LOG probabilities, dense) */
double **new_transition_probs(int nnodes, int *ptree, int *ages, double treelen,
                              intstate *istates, int nstates, int ntimes,
                              double *times, double *time_steps, int *nbranches,
                              int *nrecombs, int *ncoals, double *popsizes,
                              double rho);
double **new_transition_probs_switch(
    int *ptree, int *last_ptree, int nnodes, int recomb_node, int recomb_time,
    int coal_node, int coal_time, int *ages_index, int *last_ages_index,
    double treelen, double last_treelen, intstate *istates1, int nstates1,
    intstate *istates2, int nstates2, int ntimes, double *times,
    double *time_steps, int *nbranches, int *nrecombs, int *ncoals,
    double *popsizes, double rho);
void delete_transition_probs(double **transmat, int nstates);

/* total_prob.cpp:316-377 */
double arghmm_likelihood(LocalTrees *trees, double *times, int ntimes, double mu,
                         char **seqs, int nseqs, int seqlen);
double arghmm_prior_prob(LocalTrees *trees, double *times, int ntimes,
                         double *popsizes, double rho);
double arghmm_joint_prob(LocalTrees *trees, double *times, int ntimes,
                         double *popsizes, double mu, double rho, char **seqs,
                         int nseqs, int seqlen);

/* hmm.cpp:13-98 generic dense log-space HMM */
void forward_step(double *col1, double *col2, int nstates1, int nstates2,
                  double **trans, double *emit);
void forward_alg(int n, int nstates, double **trans, double **emit,
                 double **fw);
void backward_alg(int n, int nstates, double **trans, double **emit,
                  double **bw);
void sample_hmm_posterior(int n, int nstates, double **trans, double **fw,
                          int *path);
int sample_hmm_posterior_step(int nstates1, double **trans, double *col1,
                              int state2);

#ifdef __cplusplus
}
#endif

#endif /* ARGWEAVER_B200_H */
