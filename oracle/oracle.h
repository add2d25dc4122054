/* oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of ARGweaver's threading-HMM hot path, written from
 * the reference's algorithm (each function in oracle.c cites the reference
 * file:line it follows).  It is the checker used by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg.  Nothing in the
 * product (argweaver_b200/) may include, link or call it.
 *
 * Parity status: PINNED.  oracle.c is validated against golden vectors dumped
 * from the unmodified reference compiled in oracle/_ref (see oracle/ref_dump.cpp,
 * tests/test_oracle_vs_reference.py, tests/golden/).
 */
#ifndef AWB_ORACLE_H
#define AWB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flattened thread-sampling problem: same content as awb_problem in
 * include/argweaver_b200.h, declared independently on purpose. */
typedef struct {
    int ntimes;
    const double *times;      /* [ntimes] */
    const double *popsizes;   /* [ntimes] */
    double rho, mu;           /* already multiplied by the compression factor */

    int nseqs, seqlen;
    const unsigned char *seqs; /* [nseqs][seqlen] dense rows */
    int nleaves;
    const int *seqids;        /* [nleaves] row of seqs for tree leaf j */
    int new_chrom;            /* external mode: row of the sequence to thread */

    int internal;             /* 0 = thread a new leaf, 1 = re-thread a subtree */
    int minage;               /* StatesModel.minage (always 0 in live code) */

    int ntrees, nnodes;
    int start_coord;
    const int *ptrees;        /* [ntrees][nnodes] parent index, -1 = root */
    const int *ages;          /* [ntrees][nnodes] time index */
    const int *sprs;          /* [ntrees][4] recomb_node, recomb_time, coal_node, coal_time */
    const int *mappings;      /* [ntrees][nnodes] previous-tree node -> this tree (-1 broken) */
    const int *blocklens;     /* [ntrees] */
    const int *subtree_roots; /* [ntrees] child[0] of the root (internal mode), else -1 */
} orc_problem;

/* Everything the reference computes for one call, flattened.
 * Per-state arrays use rows of max(nstates[b],1) entries at row_off[b];
 * per-site tables (emit, fw) use fw_off[b] + (i-start_b)*max(nstates[b],1). */
typedef struct {
    int *nstates;             /* [ntrees] */
    int64_t *state_off;       /* [ntrees+1] offsets into states (true counts) */
    int *states;              /* [sum nstates][2] (node,time) */
    int64_t *row_off;         /* [ntrees+1] */
    int64_t *fw_off;          /* [ntrees+1] */
    int64_t *sw1_off;         /* [ntrees+1] offsets of determ/determprob rows */
    int *nbranches, *nrecombs, *ncoals;   /* [ntrees][ntimes] */
    double *tm[9];            /* D,E,lnB,lnE2,lnNegG1,G2,G3,lnG4,norecombs: [ntrees][ntimes] */
    int *tm_minage;           /* [ntrees] */
    int *sw_determ;           /* rows of max(nstates[b-1],1) for b>=1 */
    double *sw_determprob;
    double *sw_recombrow;     /* rows at row_off[b] */
    double *sw_recoalrow;
    int *sw_recombsrc, *sw_recoalsrc;     /* [ntrees] */
    double *emit;             /* fw layout */
    double *fw;               /* fw layout */
    int *path;                /* [seqlen of trees] */
    double logZ;              /* sum over columns of log(column norm) */
    int first_bad_site;       /* -1, or first site whose column max was <= 0 */
} orc_result;

/* Allocate / free a result sized for the problem. */
orc_result *orc_result_new(const orc_problem *p);
void orc_result_free(orc_result *r);

/* Setup: states, lineage counts, transition vectors, switch matrices, emissions.
 * (states.cpp, local_tree.cpp:34-219, trans.cpp:26-115,538-739, emit.cpp:650-845) */
void orc_setup(const orc_problem *p, orc_result *r);

/* Forward recursion over all blocks (sample_thread.cpp:394-460).
 * prior: NULL -> calc_state_priors; else the caller-supplied first column. */
void orc_forward(const orc_problem *p, orc_result *r, const double *prior);

/* Stochastic traceback (sample_thread.cpp:522-569) driven by pre-drawn libc
 * rand() integers, consumed in the reference's order (last site first).
 * last_state: -1 -> sample the last column; else the given state index.
 * Returns the number of draws consumed. */
int orc_traceback(const orc_problem *p, orc_result *r, const int *rand_ints,
                  int rand_max, int last_state);

/* Convenience: setup + forward + traceback. */
int orc_thread_sample(const orc_problem *p, orc_result *r, const int *rand_ints,
                      int rand_max);

/* Small pieces exported for unit tests */
int orc_get_states(int nnodes, const int *ptree, const int *ages, int root,
                   int subtree_root, int ntimes, int internal, int minage,
                   int *states /* [.][2] */);
void orc_count_lineages(int nnodes, const int *ptree, const int *ages, int root,
                        int subtree_root, int ntimes, int internal,
                        int *nbranches, int *nrecombs, int *ncoals);
void orc_time_steps(const double *times, int ntimes, double *time_steps,
                    double *coal_time_steps);
double orc_get_time(const double *const tm[9], int a, int b, int c, int minage,
                    int same_node);

/* Generic dense log-space HMM (hmm.cpp:13-98) */
void orc_hmm_forward_alg(int n, int nstates, const double *trans /*[S][S]*/,
                         const double *emit /*[n][S]*/, double *fw /*[n][S]*/);

#ifdef __cplusplus
}
#endif

#endif /* AWB_ORACLE_H */
