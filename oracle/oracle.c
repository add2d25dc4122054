/* oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Plain-C restatement of the reference's threading-HMM hot path.  Citations are
 * relative to the reference tree (mdrasmus/argweaver, src/argweaver/).  The
 * arithmetic deliberately follows the reference's operation order so that the
 * restatement agrees with the compiled reference to ~1e-15; it makes no attempt
 * to be fast.
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define IMAX(a, b) ((a) > (b) ? (a) : (b))
#define IMIN(a, b) ((a) < (b) ? (a) : (b))

/* ------------------------------------------------------------------ model */

/* model.h:322-333 (setup_time_steps), model.cpp:9-23 (get_coal_time_steps) */
void orc_time_steps(const double *times, int ntimes, double *time_steps,
                    double *coal_time_steps)
{
    for (int i = 0; i < ntimes - 1; i++)
        time_steps[i] = times[i + 1] - times[i];
    time_steps[ntimes - 1] = INFINITY;

    double *times2 = (double *) malloc(sizeof(double) * (2 * ntimes + 1));
    for (int i = 0; i < ntimes - 1; i++) {
        times2[2 * i] = times[i];
        times2[2 * i + 1] = sqrt((times[i + 1] + 1.0) * (times[i] + 1.0));
    }
    times2[2 * ntimes] = times[ntimes - 1];
    /* NOTE: the reference leaves times2[2*ntimes-2], times2[2*ntimes-1]
     * unset and never reads past coal_time_steps[2*ntimes-3] + the INF slot */
    times2[2 * ntimes - 2] = times[ntimes - 1];
    times2[2 * ntimes - 1] = times[ntimes - 1];
    for (int i = 0; i < 2 * ntimes - 2; i++)
        coal_time_steps[i] = times2[IMIN(i + 1, 2 * ntimes)] - times2[i];
    coal_time_steps[2 * ntimes - 2] = INFINITY;
    coal_time_steps[2 * ntimes - 1] = INFINITY; /* never consumed */
    free(times2);
}

typedef struct {
    int ntimes;
    const double *times;
    const double *popsizes;
    double *time_steps;
    double *coal_time_steps;
    double rho, mu;
    double mintime;      /* model.h:211 */
    int removed_root_time; /* model.h:207 */
} omodel;

/* ------------------------------------------------------------------- tree */

typedef struct {
    int V;
    const int *parent;
    const int *age;
    int *c0, *c1;
    int root;
} otree;

/* local_tree.h:188-229 (set_ptree): children in node-index order; the order of
 * the root's children is overridden when the caller states the subtree root
 * (internal mode: child[0] = subtree root, thread.cpp:1521-1539). */
static void tree_init(otree *t, int V, const int *parent, const int *age,
                      int subtree_root)
{
    t->V = V;
    t->parent = parent;
    t->age = age;
    t->c0 = (int *) malloc(sizeof(int) * V);
    t->c1 = (int *) malloc(sizeof(int) * V);
    t->root = -1;
    for (int i = 0; i < V; i++)
        t->c0[i] = t->c1[i] = -1;
    for (int i = 0; i < V; i++) {
        int p = parent[i];
        if (p != -1) {
            if (t->c0[p] == -1)
                t->c0[p] = i;
            else
                t->c1[p] = i;
        } else {
            t->root = i;
        }
    }
    if (subtree_root >= 0 && t->root >= 0 && t->c1[t->root] == subtree_root) {
        int tmp = t->c0[t->root];
        t->c0[t->root] = t->c1[t->root];
        t->c1[t->root] = tmp;
    }
}

static void tree_free(otree *t)
{
    free(t->c0);
    free(t->c1);
}

static int tree_is_leaf(const otree *t, int i) { return t->c0[i] == -1; }

/* local_tree.h:394-404 */
static int tree_sibling(const otree *t, int node)
{
    int p = t->parent[node];
    if (p == -1)
        return -1;
    return t->c0[p] == node ? t->c1[p] : t->c0[p];
}

/* local_tree.h:274-304 */
static void tree_postorder(const otree *t, int *order)
{
    char *visit = (char *) calloc(t->V, 1);
    int i;
    for (i = 0; i < t->V; i++) {
        if (!tree_is_leaf(t, i))
            break;
        order[i] = i;
    }
    int end = i;
    for (i = 0; i < t->V; i++) {
        int p = t->parent[order[i]];
        if (p != -1) {
            visit[p]++;
            if (visit[p] == 2)
                order[end++] = p;
        }
    }
    free(visit);
}

/* local_tree.h:361-368 */
static double tree_dist(const otree *t, int node, const double *times)
{
    int p = t->parent[node];
    if (p != -1)
        return times[t->age[p]] - times[t->age[node]];
    return 0.0;
}

/* ----------------------------------------------------------------- states */

/* states.cpp:53-74 (external), :109-166 (internal) */
static int get_states(const otree *t, int ntimes, int internal, int minage,
                      int *states)
{
    int n = 0;
    if (!internal) {
        for (int i = 0; i < t->V; i++) {
            int time = t->age[i];
            int p = t->parent[i];
            if (p == -1) {
                for (; time < ntimes - 1; time++) {
                    states[2 * n] = i; states[2 * n + 1] = time; n++;
                }
            } else {
                int pa = t->age[p];
                for (; time <= pa; time++) {
                    states[2 * n] = i; states[2 * n + 1] = time; n++;
                }
            }
        }
        return n;
    }

    if (t->age[t->root] < ntimes)
        return 0; /* fully specified tree */

    int subtree_root = t->c0[t->root];
    minage = IMAX(minage, t->age[subtree_root]);

    char *ignore = (char *) calloc(t->V, 1);
    int *stack = (int *) malloc(sizeof(int) * t->V);
    int sp = 0;
    ignore[t->root] = 1;
    stack[sp++] = subtree_root;
    while (sp > 0) {
        int node = stack[--sp];
        ignore[node] = 1;
        if (!tree_is_leaf(t, node)) {
            stack[sp++] = t->c0[node];
            stack[sp++] = t->c1[node];
        }
    }
    for (int i = 0; i < t->V; i++) {
        int time = IMAX(t->age[i], minage);
        int p = t->parent[i];
        if (ignore[i])
            continue;
        if (p == t->root) {
            for (; time < ntimes - 1; time++) {
                states[2 * n] = i; states[2 * n + 1] = time; n++;
            }
        } else {
            int pa = t->age[p];
            for (; time <= pa; time++) {
                states[2 * n] = i; states[2 * n + 1] = time; n++;
            }
        }
    }
    free(ignore);
    free(stack);
    return n;
}

int orc_get_states(int nnodes, const int *ptree, const int *ages, int root,
                   int subtree_root, int ntimes, int internal, int minage,
                   int *states)
{
    otree t;
    (void) root;
    tree_init(&t, nnodes, ptree, ages, internal ? subtree_root : -1);
    int n = get_states(&t, ntimes, internal, minage, states);
    tree_free(&t);
    return n;
}

/* states.h:53-131 NodeStateLookup */
typedef struct {
    const int *states;
    int nstates, nnodes;
    int *node_offset, *state_lookup, *per_node;
} olookup;

static void lookup_init(olookup *L, const int *states, int nstates, int nnodes)
{
    const int MAXTIME = 1000000;
    L->states = states;
    L->nstates = nstates;
    L->nnodes = nnodes;
    L->node_offset = (int *) malloc(sizeof(int) * nnodes);
    L->state_lookup = (int *) malloc(sizeof(int) * IMAX(nstates, 1));
    L->per_node = (int *) malloc(sizeof(int) * nnodes);
    int *mint = (int *) malloc(sizeof(int) * nnodes);
    for (int i = 0; i < nnodes; i++) {
        L->per_node[i] = 0;
        mint[i] = MAXTIME;
    }
    for (int i = 0; i < nstates; i++) {
        int node = states[2 * i], time = states[2 * i + 1];
        L->state_lookup[i] = -1;
        L->per_node[node]++;
        mint[node] = IMIN(mint[node], time);
    }
    int offset = 0;
    for (int i = 0; i < nnodes; i++) {
        L->node_offset[i] = offset - mint[i];
        offset += L->per_node[i];
    }
    for (int i = 0; i < nstates; i++) {
        int j = L->node_offset[states[2 * i]] + states[2 * i + 1];
        L->state_lookup[j] = i;
    }
    free(mint);
}

static void lookup_free(olookup *L)
{
    free(L->node_offset);
    free(L->state_lookup);
    free(L->per_node);
}

static int lookup(const olookup *L, int node, int time)
{
    if (L->per_node[node] == 0)
        return -1;
    int i = L->node_offset[node] + time;
    if (i < 0 || i >= L->nstates)
        return -1;
    int s = L->state_lookup[i];
    if (s == -1)
        return -1;
    if (L->states[2 * s] != node || L->states[2 * s + 1] != time)
        return -1;
    return s;
}

/* --------------------------------------------------------------- lineages */

/* local_tree.cpp:34-69 (external), :82-131 (internal) */
static void count_lineages(const otree *t, int ntimes, int internal,
                           int *nbranches, int *nrecombs, int *ncoals)
{
    for (int i = 0; i < ntimes; i++)
        nbranches[i] = nrecombs[i] = ncoals[i] = 0;

    if (!internal) {
        for (int i = 0; i < t->V; i++) {
            int p = t->parent[i];
            int pa = (p == -1) ? ntimes - 2 : t->age[p];
            for (int j = t->age[i]; j < pa; j++) {
                nbranches[j]++; nrecombs[j]++; ncoals[j]++;
            }
            nrecombs[pa]++;
            ncoals[pa]++;
            if (p == -1)
                nbranches[pa]++;
        }
        nbranches[ntimes - 1] = 1;
        return;
    }

    const int subtree_root = t->c0[t->root];
    const int minage = t->age[subtree_root];
    for (int i = 0; i < t->V; i++) {
        if (i == subtree_root || i == t->root)
            continue;
        int p = t->parent[i];
        int pa = (p == t->root) ? ntimes - 2 : t->age[p];
        for (int j = t->age[i]; j < pa; j++) {
            nbranches[j]++; nrecombs[j]++; ncoals[j]++;
        }
        nrecombs[pa]++;
        ncoals[pa]++;
        if (p == t->root)
            nbranches[pa]++;
    }
    for (int i = 0; i < minage; i++) {
        nbranches[i]--; ncoals[i]--; nrecombs[i]--;
    }
    nbranches[ntimes - 1] = 1;
}

void orc_count_lineages(int nnodes, const int *ptree, const int *ages, int root,
                        int subtree_root, int ntimes, int internal,
                        int *nbranches, int *nrecombs, int *ncoals)
{
    otree t;
    (void) root;
    tree_init(&t, nnodes, ptree, ages, internal ? subtree_root : -1);
    count_lineages(&t, ntimes, internal, nbranches, nrecombs, ncoals);
    tree_free(&t);
}

/* local_tree.cpp:136-155 */
static double get_treelen(const otree *t, const double *times, int use_basal)
{
    double treelen = 0.0;
    for (int i = 0; i < t->V; i++) {
        int p = t->parent[i];
        int age = t->age[i];
        if (p == -1) {
            if (use_basal)
                treelen += times[age + 1] - times[age];
        } else {
            treelen += times[t->age[p]] - times[age];
        }
    }
    return treelen;
}

/* local_tree.cpp:158-176 */
static double get_treelen_internal(const otree *t, const double *times)
{
    double treelen = 0.0;
    for (int i = 0; i < t->V; i++) {
        int p = t->parent[i];
        if (p == t->root || p == -1)
            continue;
        treelen += times[t->age[p]] - times[t->age[i]];
    }
    return treelen;
}

/* local_tree.cpp:179-203 (use_basal = false at the only call site) */
static double get_treelen_branch(const otree *t, const double *times, int node,
                                 int time, double treelen)
{
    double blen = times[time];
    double treelen2 = treelen + blen;
    if (node == t->root)
        treelen2 += blen - times[t->age[t->root]];
    return treelen2;
}

/* local_tree.cpp:206-219 */
static double get_basal_branch(const otree *t, const double *times, int node,
                               int time)
{
    if (node == t->root)
        return times[time + 1] - times[time];
    int rooti = t->age[t->root];
    return times[rooti + 1] - times[rooti];
}

/* ------------------------------------------------------------ transitions */

/* common.h:157-163 */
static double logadd(double lna, double lnb)
{
    if (lna == -INFINITY)
        return lnb;
    if (lnb == -INFINITY)
        return lna;
    return fmax(lna, lnb) + log1p(exp(-fabs(lna - lnb)));
}

enum { TM_D, TM_E, TM_LNB, TM_LNE2, TM_LNNEGG1, TM_G2, TM_G3, TM_LNG4,
       TM_NORECOMBS };

/* trans.cpp:26-115 */
static void calc_transition_probs(const otree *t, const omodel *m,
                                  const int *nbranches, const int *nrecombs,
                                  const int *ncoals, int internal, int minage,
                                  double *const tm[9], int *out_minage)
{
    const int ntimes = m->ntimes;
    const double *times = m->times;
    const double *time_steps = m->time_steps;
    const double rho = m->rho;

    double *coal_rates_alloc = (double *) malloc(sizeof(double) * (2 * ntimes + 1));
    double *coal_rates = coal_rates_alloc + 1;
    coal_rates[-1] = 0.0;
    for (int i = 0; i < 2 * ntimes; i++)
        coal_rates[i] = m->coal_time_steps[i] * nbranches[i / 2] /
            (2.0 * m->popsizes[i / 2]);

    double *C_alloc = (double *) malloc(sizeof(double) * (2 * ntimes + 2));
    double *C = C_alloc + 2;
    C[-2] = 0.0;
    C[-1] = 0.0;
    for (int b = 0; b < 2 * ntimes - 1; b++)
        C[b] = C[b - 1] + coal_rates[b];

    int root_age_index;
    double root_age, treelen;
    if (internal) {
        const int subtree_root = t->c0[t->root];
        const int maintree_root = t->c1[t->root];
        const double subtree_age = times[t->age[subtree_root]];
        root_age_index = t->age[maintree_root];
        root_age = times[root_age_index];
        treelen = get_treelen_internal(t, times) - subtree_age;
        minage = IMAX(minage, t->age[subtree_root]);
    } else {
        root_age_index = t->age[t->root];
        root_age = times[root_age_index];
        treelen = get_treelen(t, times, 0);
    }
    *out_minage = minage;

    for (int b = 0; b < ntimes - 1; b++) {
        double treelen2 = treelen + times[b];
        double treelen2_b;
        if (b > root_age_index) {
            treelen2 += times[b] - root_age;
            treelen2_b = treelen2 + time_steps[b];
        } else {
            treelen2_b = treelen2 + time_steps[root_age_index];
        }
        double term = C[2 * b - 1] + log(
            time_steps[b] * (nbranches[b] + 1.0) / (nrecombs[b] + 1.0));
        if (b == 0)
            tm[TM_LNB][b] = term;
        else
            tm[TM_LNB][b] = logadd(tm[TM_LNB][b - 1], term);
        tm[TM_LNE2][b] = -C[2 * b - 2] +
            (b < ntimes - 2 ?
             log(1 - exp(-coal_rates[2 * b] - coal_rates[2 * b - 1])) : 0.0);
        tm[TM_LNNEGG1][b] = C[2 * b - 1] + log(-time_steps[b] * (
            (nbranches[b] / (nrecombs[b] + 1.0 + (b < root_age_index ? 1 : 0)))
            - (nbranches[b] + 1.0) / (nrecombs[b] + 1.0)));
        tm[TM_G2][b] = (b < ntimes - 2 ? 1.0 - exp(-coal_rates[2 * b]) : 1.0) *
            time_steps[b] * (nbranches[b] + 1.0) / (nrecombs[b] + 1.0);
        tm[TM_G3][b] = (b < ntimes - 2 ? 1.0 - exp(-coal_rates[2 * b]) : 1.0) *
            time_steps[b] *
            (nbranches[b] / (nrecombs[b] + 1.0 + (b < root_age_index ? 1 : 0)));
        tm[TM_LNG4][b] = -C[2 * b - 2] +
            (b < ntimes - 2 ?
             log(1.0 - exp(-coal_rates[2 * b] - coal_rates[2 * b - 1])) : 0.0);
        tm[TM_D][b] = (1.0 - exp(-rho * treelen2)) / treelen2_b;
        tm[TM_E][b] = 1.0 / ncoals[b];
        tm[TM_NORECOMBS][b] = exp(-fmax(rho * treelen2, rho));
    }
    tm[TM_E][ntimes - 2] = 1.0 / ncoals[ntimes - 2];
    for (int k = 0; k < 9; k++)
        tm[k][ntimes - 1] = 0.0; /* unused slot, zeroed for comparisons */

    free(coal_rates_alloc);
    free(C_alloc);
}

/* trans.h:89-128 */
double orc_get_time(const double *const tm[9], int a, int b, int c, int minage,
                    int same_node)
{
    if (a < minage || b < minage)
        return 0.0;
    const double *D = tm[TM_D], *E = tm[TM_E], *lnB = tm[TM_LNB],
        *lnE2 = tm[TM_LNE2], *lnNegG1 = tm[TM_LNNEGG1], *G2 = tm[TM_G2],
        *G3 = tm[TM_G3], *lnG4 = tm[TM_LNG4], *norecombs = tm[TM_NORECOMBS];
    double term1 = D[a] * E[b];
    double minage_term = 0.0;
    if (minage > 0)
        minage_term = exp(lnG4[b] + lnB[minage - 1]);

    if (!same_node) {
        if (a < b) {
            return term1 * (exp(lnE2[b] + lnB[a]) -
                            exp(lnE2[b] + lnNegG1[a]) - minage_term);
        } else if (a == b) {
            return term1 * ((b > 0 ? exp(lnE2[b] + lnB[b - 1]) : 0.0) +
                            G3[b] - minage_term);
        } else {
            return term1 * ((b > 0 ? exp(lnE2[b] + lnB[b - 1]) : 0.0) +
                            G2[b] - minage_term);
        }
    } else {
        double c_term = (c > 0 ? exp(lnG4[b] + lnB[c - 1]) : 0.0);
        if (a < b) {
            return term1 * (2 * (exp(lnE2[b] + lnB[a]) -
                                 exp(lnE2[b] + lnNegG1[a]))
                            - c_term - minage_term);
        } else if (a == b) {
            return term1 * ((2 * ((b > 0 ? exp(lnE2[b] + lnB[b - 1]) : 0.0) +
                                  G3[b])) - c_term - minage_term)
                + norecombs[a];
        } else {
            return term1 * ((2 * ((b > 0 ? exp(lnE2[b] + lnB[b - 1]) : 0.0) +
                                  G2[b])) - c_term - minage_term);
        }
    }
}

/* trans.h:64-83 TransMatrix::get */
static double trans_get(const double *const tm[9], const otree *t,
                        const int *states, int nstates, int internal,
                        int i, int j)
{
    int minage = 0;
    if (internal) {
        if (nstates == 0)
            return 1.0;
        minage = t->age[t->c0[t->root]];
    }
    const int node1 = states[2 * i], a = states[2 * i + 1];
    const int node2 = states[2 * j], b = states[2 * j + 1];
    const int c = t->age[node2];
    return orc_get_time(tm, a, b, c, minage, node1 == node2);
}

/* trans.cpp:785-812 */
static double state_prior(int time, const int *nbranches, const int *ncoals,
                          int minage, const double *popsizes,
                          const double *coal_time_steps, int ntimes)
{
    const int b = time;
    if (b < minage)
        return 0.0;
    double sum = 0.0;
    for (int m = 2 * minage; m < 2 * b - 1; m++)
        sum += coal_time_steps[m] * nbranches[m / 2] / (2.0 * popsizes[m / 2]);
    double p = exp(-sum) / ncoals[b];
    if (b < ntimes - 2) {
        double Z = 0.0;
        if (b > minage)
            Z = coal_time_steps[2 * b - 1] * nbranches[b - 1] /
                (2.0 * popsizes[b - 1]);
        p *= 1.0 - exp(-coal_time_steps[2 * b] * nbranches[b] /
                       (2.0 * popsizes[b]) - Z);
    }
    return p;
}

/* ----------------------------------------------------------------- switch */

typedef struct { int recomb_node, recomb_time, coal_node, coal_time; } ospr;

/* trans.cpp:156-275 */
static void get_deterministic_transitions(
    const otree *last_tree, const otree *tree, const ospr *spr,
    const int *mapping, const int *states1, int nstates1,
    const olookup *state2_lookup, int *next_states, int internal)
{
    for (int i = 0; i < nstates1; i++) {
        const int node1 = states1[2 * i];
        const int time1 = states1[2 * i + 1];

        if ((node1 == spr->coal_node && time1 == spr->coal_time) ||
            (node1 == spr->recomb_node && time1 == spr->recomb_time)) {
            next_states[i] = -1;
        } else if (node1 != spr->recomb_node) {
            int node2;
            int disrupt = 0;
            if (last_tree->c0[node1] == -1) {
                node2 = node1;
            } else {
                const int child1 = last_tree->c0[node1];
                const int child2 = last_tree->c1[node1];
                if (spr->recomb_node == child1) {
                    node2 = mapping[child2];
                    disrupt = 1;
                } else if (spr->recomb_node == child2) {
                    node2 = mapping[child1];
                    disrupt = 1;
                } else {
                    node2 = mapping[node1];
                }
            }

            if ((spr->coal_node == node1 && spr->coal_time < time1) ||
                (mapping[spr->coal_node] == node2 && spr->coal_time < time1) ||
                (disrupt && mapping[spr->coal_node] == node2 &&
                 spr->coal_time <= time1)) {
                node2 = tree->parent[node2];
            }

            if (internal && tree->age[node2] > time1) {
                next_states[i] = -1;
                continue;
            }
            const int p = tree->parent[node2];
            if (p != -1) {
                if (internal && time1 > tree->age[p]) {
                    next_states[i] = -1;
                    continue;
                }
            }
            next_states[i] = lookup(state2_lookup, node2, time1);
        } else {
            if (spr->recomb_time > time1) {
                next_states[i] = lookup(state2_lookup,
                                        mapping[spr->recomb_node], time1);
            } else {
                const int parent = last_tree->parent[spr->recomb_node];
                const int time2 = last_tree->age[parent];
                const int other = (last_tree->c1[parent] == spr->recomb_node ?
                                   last_tree->c0[parent] : last_tree->c1[parent]);
                const int node2 = (other == spr->coal_node ?
                                   tree->parent[mapping[other]] : mapping[other]);
                next_states[i] = lookup(state2_lookup, node2, time2);
            }
        }
    }
}

/* trans.cpp:278-313 */
static void get_recomb_transition_switch(
    const otree *tree, const otree *last_tree, const ospr *spr,
    const int *mapping, const olookup *state2_lookup, int next_states[2])
{
    int parent = last_tree->parent[spr->recomb_node];
    int time2 = last_tree->age[parent];
    int other = tree_sibling(last_tree, spr->recomb_node);
    int node2 = (other == spr->coal_node ?
                 tree->parent[mapping[other]] : mapping[other]);
    next_states[0] = lookup(state2_lookup, mapping[spr->recomb_node],
                            spr->recomb_time);
    next_states[1] = lookup(state2_lookup, node2, time2);
}

/* trans.cpp:317-413 */
static double calc_recomb(const otree *last_tree, const omodel *m,
                          const int *nbranches, const int *nrecombs,
                          const ospr *spr, int state1_node, int state1_time,
                          double last_treelen, int internal)
{
    int a = state1_time;
    const int k = spr->recomb_time;
    double last_treelen_b;
    int root_age;

    if (internal) {
        int subtree_root = last_tree->c0[last_tree->root];
        int maintree_root = last_tree->c1[last_tree->root];
        root_age = last_tree->age[maintree_root];

        if (spr->coal_node == subtree_root) {
            if (a < spr->coal_time)
                return 0.0;
            if (spr->recomb_node == maintree_root) {
                if (state1_node != maintree_root)
                    return 0.0;
            }
        }

        int ptr = spr->coal_node;
        int ptr2 = -1, ptr3 = -1;
        while (ptr != last_tree->root) {
            ptr2 = ptr;
            ptr = last_tree->parent[ptr];
        }
        ptr = spr->recomb_node;
        while (ptr != last_tree->root) {
            ptr3 = ptr;
            ptr = last_tree->parent[ptr];
        }
        if (ptr2 == subtree_root && ptr3 == maintree_root &&
            state1_time == spr->recomb_time) {
            ptr = last_tree->parent[state1_node];
            while (ptr != last_tree->root) {
                if (ptr == spr->recomb_node)
                    return 0.0;
                ptr = last_tree->parent[ptr];
            }
        }

        last_treelen += m->times[a] - m->times[last_tree->age[subtree_root]];
        if (a > root_age) {
            last_treelen += m->times[a] - m->times[root_age];
            last_treelen_b = last_treelen + m->time_steps[a];
        } else {
            last_treelen_b = last_treelen + m->time_steps[root_age];
        }
    } else {
        root_age = last_tree->age[last_tree->root];
        last_treelen = get_treelen_branch(last_tree, m->times, state1_node,
                                          state1_time, last_treelen);
        last_treelen_b = last_treelen + get_basal_branch(
            last_tree, m->times, state1_node, state1_time);
    }

    int nbranches_k = nbranches[k] + (k < a ? 1 : 0);
    int nrecombs_k = nrecombs[k] + (k <= a ? 1 : 0) + (k == a ? 1 : 0) -
        (k >= IMAX(root_age, a) ? 1 : 0);
    double p = nbranches_k * m->time_steps[k] / (nrecombs_k * last_treelen_b) *
        (1.0 - exp(-fmax(m->rho * last_treelen, m->rho)));
    return p;
}

/* trans.cpp:417-440 */
static void calc_recoal_sums(const omodel *m, const int *nbranches,
                             const ospr *spr, int recomb_parent_age,
                             double *sums, double *sums2)
{
    const int k = spr->recomb_time;
    const int j = spr->coal_time;
    double sum = 0.0;
    for (int mm = 2 * k; mm < 2 * j - 1; mm++) {
        int nbranches_m = nbranches[mm / 2] - (mm / 2 < recomb_parent_age ? 1 : 0);
        sum += m->coal_time_steps[mm] * nbranches_m / (2.0 * m->popsizes[mm / 2]);
    }
    *sums = sum;
    sum = 0.0;
    sums2[2 * k] = sum;
    for (int mm = 2 * k; mm < 2 * j - 1; mm++) {
        sum += m->coal_time_steps[mm] / (2.0 * m->popsizes[mm / 2]);
        sums2[mm + 1] = sum;
    }
}

/* trans.cpp:443-502 */
static double calc_recoal(const otree *last_tree, const omodel *m,
                          const int *nbranches, const int *ncoals,
                          const ospr *spr, int state1_time,
                          int recomb_parent_age, int internal)
{
    int a = state1_time;
    const int k = spr->recomb_time;
    const int j = spr->coal_time;

    int nbranches_j = nbranches[j] - (j < recomb_parent_age ? 1 : 0) +
        (j < a ? 1 : 0);
    int ncoals_j = ncoals[j] - (j <= recomb_parent_age ? 1 : 0) -
        (j == recomb_parent_age ? 1 : 0) + (j <= a ? 1 : 0) + (j == a ? 1 : 0);
    int over = 0;
    if (internal) {
        int subtree_root = last_tree->c0[last_tree->root];
        int maintree_root = last_tree->c1[last_tree->root];
        if (spr->recomb_node == maintree_root) {
            if (spr->coal_time >= last_tree->age[subtree_root]) {
                over = 1;
                nbranches_j = 1;
                ncoals_j++;
            }
        }
    }
    double p = 1.0 / ncoals_j;
    if (j < m->ntimes - 2) {
        double Z = 0.0;
        if (j > k) {
            int b1 = nbranches[j - 1] - (j - 1 < recomb_parent_age ? 1 : 0) +
                (j - 1 < a ? 1 : 0);
            if (over)
                b1 = 1;
            Z = m->coal_time_steps[2 * j - 1] * b1 / (2.0 * m->popsizes[j - 1]);
        }
        p *= 1.0 - exp(-m->coal_time_steps[2 * j] * nbranches_j /
                       (2.0 * m->popsizes[j]) - Z);
    }
    return p;
}

/* trans.cpp:505-534 */
static double calc_recomb_recoal(const otree *last_tree, const omodel *m,
                                 const int *nbranches, const int *nrecombs,
                                 const int *ncoals, const ospr *spr,
                                 int state1_node, int state1_time,
                                 int recomb_parent_age, double last_treelen,
                                 int internal)
{
    int a = state1_time;
    const int k = spr->recomb_time;
    const int j = spr->coal_time;
    double p = calc_recomb(last_tree, m, nbranches, nrecombs, spr, state1_node,
                           state1_time, last_treelen, internal);
    double sum = 0.0;
    for (int mm = 2 * k; mm < 2 * j - 1; mm++) {
        int nbranches_m = nbranches[mm / 2] -
            (mm / 2 < recomb_parent_age ? 1 : 0) + (mm / 2 < a ? 1 : 0);
        sum += m->coal_time_steps[mm] * nbranches_m / (2.0 * m->popsizes[mm / 2]);
    }
    p *= exp(-sum);
    p *= calc_recoal(last_tree, m, nbranches, ncoals, spr, state1_time,
                     recomb_parent_age, internal);
    return p;
}

typedef struct {
    int nstates1, nstates2;
    int recoalsrc, recombsrc;
    int *determ;        /* [max(nstates1,1)] */
    double *determprob;
    double *recoalrow;  /* [max(nstates2,1)] */
    double *recombrow;
} oswitch;

/* trans.cpp:538-739.  lineages are those of last_tree. */
static void calc_transition_probs_switch(
    const otree *tree, const otree *last_tree, const ospr *spr,
    const int *mapping, const int *states1, int nstates1, const int *states2,
    int nstates2, const omodel *m, const int *nbranches, const int *nrecombs,
    const int *ncoals, oswitch *sw, int internal)
{
    int recomb_parent_age;
    const int ntimes = m->ntimes;

    sw->nstates1 = nstates1;
    sw->nstates2 = nstates2;
    for (int j = 0; j < IMAX(nstates1, 1); j++) {
        sw->determ[j] = -1;
        sw->determprob[j] = 0.0;
    }
    for (int j = 0; j < IMAX(nstates2, 1); j++) {
        sw->recoalrow[j] = 0.0;
        sw->recombrow[j] = 0.0;
    }

    double last_treelen = internal ?
        get_treelen_internal(last_tree, m->times) :
        get_treelen(last_tree, m->times, 0);

    if (internal) {
        if (nstates1 == 0) {
            if (nstates2 == 0) {
                sw->determ[0] = 0;
                sw->determprob[0] = 1.0;
                sw->recoalsrc = -1;
                sw->recombsrc = -1;
                return;
            }
            const int maintree_root = tree->c1[tree->root];
            for (int j = 0; j < nstates2; j++) {
                if (states2[2 * j] == maintree_root &&
                    states2[2 * j + 1] == spr->coal_time) {
                    sw->determ[0] = j;
                    sw->determprob[0] = 1.0;
                    sw->recoalsrc = -1;
                    sw->recombsrc = -1;
                    return;
                }
            }
            fprintf(stderr, "oracle: switch 0->n: no target state\n");
            abort();
        }
        if (nstates2 == 0) {
            for (int i = 0; i < nstates1; i++)
                sw->determ[i] = 0;
            for (int i = 0; i < nstates1; i++) {
                if (states1[2 * i] == spr->recomb_node &&
                    states1[2 * i + 1] > spr->recomb_time)
                    recomb_parent_age = states1[2 * i + 1];
                else
                    recomb_parent_age = last_tree->age[
                        last_tree->parent[spr->recomb_node]];
                sw->determprob[i] = calc_recomb_recoal(
                    last_tree, m, nbranches, nrecombs, ncoals, spr,
                    states1[2 * i], states1[2 * i + 1], recomb_parent_age,
                    last_treelen, internal);
            }
            sw->recoalsrc = -1;
            sw->recombsrc = -1;
            return;
        }
    }

    olookup L;
    lookup_init(&L, states2, nstates2, tree->V);
    get_deterministic_transitions(last_tree, tree, spr, mapping, states1,
                                  nstates1, &L, sw->determ, internal);

    recomb_parent_age = last_tree->age[last_tree->parent[spr->recomb_node]];
    double sums;
    double *sums2 = (double *) calloc(ntimes * 2 + 1, sizeof(double));
    double *recoals = (double *) malloc(sizeof(double) * ntimes);
    calc_recoal_sums(m, nbranches, spr, recomb_parent_age, &sums, sums2);
    for (int a = 0; a < ntimes; a++)
        recoals[a] = calc_recoal(last_tree, m, nbranches, ncoals, spr, a,
                                 recomb_parent_age, 0 /* sic: trans.cpp:620 */);

    for (int i = 0; i < nstates1; i++) {
        int j = sw->determ[i];
        if (j >= 0) {
            const int node1 = states1[2 * i], time1 = states1[2 * i + 1];
            if (node1 == spr->recomb_node && time1 > spr->recomb_time) {
                recomb_parent_age = time1;
                sw->determprob[i] = calc_recomb_recoal(
                    last_tree, m, nbranches, nrecombs, ncoals, spr, node1,
                    time1, recomb_parent_age, last_treelen, internal);
            } else {
                recomb_parent_age = last_tree->age[
                    last_tree->parent[spr->recomb_node]];
                sw->determprob[i] =
                    calc_recomb(last_tree, m, nbranches, nrecombs, spr, node1,
                                time1, last_treelen, internal) *
                    exp(-sums -
                        sums2[IMAX(IMIN(2 * spr->coal_time - 1, 2 * time1),
                                   2 * spr->recomb_time)]) *
                    recoals[time1];
            }
        }
    }

    int recoalsrc = -1, recombsrc = -1;
    for (int i = 0; i < nstates1; i++) {
        if (states1[2 * i] == spr->recomb_node &&
            states1[2 * i + 1] == spr->recomb_time)
            recombsrc = i;
        else if (states1[2 * i] == spr->coal_node &&
                 states1[2 * i + 1] == spr->coal_time)
            recoalsrc = i;
    }
    sw->recoalsrc = recoalsrc;
    sw->recombsrc = recombsrc;

    if (recombsrc != -1) {
        int nx[2];
        get_recomb_transition_switch(tree, last_tree, spr, mapping, &L, nx);
        int j = nx[0];
        if (j != -1) {
            recomb_parent_age = last_tree->age[
                last_tree->parent[spr->recomb_node]];
            sw->recombrow[j] = calc_recomb_recoal(
                last_tree, m, nbranches, nrecombs, ncoals, spr,
                states1[2 * recombsrc], states1[2 * recombsrc + 1],
                recomb_parent_age, last_treelen, internal);
        }
        j = nx[1];
        if (j != -1) {
            recomb_parent_age = states1[2 * recombsrc + 1];
            sw->recombrow[j] = calc_recomb_recoal(
                last_tree, m, nbranches, nrecombs, ncoals, spr,
                states1[2 * recombsrc], states1[2 * recombsrc + 1],
                recomb_parent_age, last_treelen, internal);
        }
    }

    if (recoalsrc != -1) {
        int node1 = states1[2 * recoalsrc];
        int time1 = states1[2 * recoalsrc + 1];
        int node3;
        int last_parent = last_tree->parent[spr->recomb_node];
        if (last_parent == node1) {
            node3 = mapping[last_tree->c1[last_parent] == spr->recomb_node ?
                            last_tree->c0[last_parent] :
                            last_tree->c1[last_parent]];
        } else {
            node3 = mapping[node1];
        }
        int parent = tree->parent[mapping[spr->recomb_node]];

        for (int j = 0; j < nstates2; j++) {
            const int node2 = states2[2 * j], time2 = states2[2 * j + 1];
            if (!((node2 == mapping[spr->recomb_node] &&
                   time2 >= spr->recomb_time) ||
                  (node2 == node3 && time2 == time1) ||
                  (node2 == parent && time2 == time1)))
                continue;
            recomb_parent_age = last_tree->age[
                last_tree->parent[spr->recomb_node]];
            ospr spr2 = *spr;
            spr2.coal_time = time2;
            sw->recoalrow[j] = calc_recomb_recoal(
                last_tree, m, nbranches, nrecombs, ncoals, &spr2, node1, time1,
                recomb_parent_age, last_treelen, internal);
        }
    }

    free(sums2);
    free(recoals);
    lookup_free(&L);
}

/* trans.h:207-219 */
static double switch_get(const oswitch *sw, int i, int j)
{
    if (i == sw->recoalsrc)
        return sw->recoalrow[j];
    else if (i == sw->recombsrc)
        return sw->recombrow[j];
    else
        return sw->determ[i] == j ? sw->determprob[i] : 0.0;
}

/* -------------------------------------------------------------- emissions */

static int dna2int(unsigned char c)
{
    switch (c) {            /* seq.cpp:15-43 */
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    }
    return -1;
}

/* emit.cpp:86-93 */
static double prob_branch(double t, double mu, int mut)
{
    const double f = 4. / 3.;
    if (!mut)
        return .25 * (1.0 + 3. * exp(-f * mu * t));
    else
        return .25 * (1.0 - exp(-f * mu * t));
}

typedef double lk_row[4];

static void leaf_row(unsigned char c, double *row)
{
    int x = dna2int(c);
    if (c == 'N' || x < 0) {
        /* emit.cpp:162-166; non-ACGT characters other than 'N' are undefined
         * behaviour in the reference (index -1) -- treated as 'N' here */
        row[0] = row[1] = row[2] = row[3] = 1.0;
    } else {
        row[0] = row[1] = row[2] = row[3] = 0.0;
        row[x] = 1.0;
    }
}

/* emit.cpp:620-645 */
static double calc_emit(lk_row *in, lk_row *out, lk_row *in2, int node1,
                        int node2, int maintree_root, const double *nomut,
                        const double *mut)
{
    double emit = 0.0;
    for (int a = 0; a < 4; a++) {
        double p1 = 0.0, p2 = 0.0, p3 = 0.0;
        for (int b = 0; b < 4; b++) {
            if (a == b) {
                p1 += in2[node1][b] * nomut[0];
                p2 += in[node2][b] * nomut[1];
                p3 += out[node2][b] * nomut[2];
            } else {
                p1 += in2[node1][b] * mut[0];
                p2 += in[node2][b] * mut[1];
                p3 += out[node2][b] * mut[2];
            }
        }
        if (node2 != maintree_root)
            emit += p1 * p2 * p3 * .25;
        else
            emit += p1 * p2 * .25;
    }
    return emit;
}

/* emit.cpp:650-845 (phased, no infinite-sites penalty).
 * seqs[j] points at the block's first site of row j; emit is [seqlen][max(S,1)]. */
static void calc_emissions(const int *states, int nstates, const otree *t,
                           const unsigned char *const *seqs, int nseqs,
                           int seqlen, const omodel *m, int internal,
                           double *emit)
{
    const int V = t->V;
    const double mintime = m->mintime;
    const double *times = m->times;
    const int maintree_root = internal ? t->c1[t->root] : t->root;
    const int subtree_root = internal ? t->c0[t->root] : t->root;

    if (internal && nstates == 0) {
        for (int i = 0; i < seqlen; i++)
            emit[i] = 1.0;
        return;
    }

    /* emit.cpp:30-58 */
    char *invariant = (char *) malloc(seqlen);
    char *masked = (char *) malloc(seqlen);
    for (int i = 0; i < seqlen; i++) {
        unsigned char c = seqs[0][i];
        int mut = 0;
        for (int j = 1; j < nseqs; j++)
            if (seqs[j][i] != c) { mut = 1; break; }
        invariant[i] = !mut;
        masked[i] = (c == 'N' && invariant[i]);
    }

    /* emit.cpp:97-116 prob_tree_mutation */
    double *muts = (double *) calloc(V, sizeof(double));
    double *nomuts = (double *) calloc(V, sizeof(double));
    for (int i = 0; i < V; i++) {
        if (i == t->root)
            continue;
        int parent_age = t->age[t->parent[i]];
        if (parent_age == m->removed_root_time)
            continue;
        double bt = fmax(times[parent_age] - times[t->age[i]], mintime);
        muts[i] = prob_branch(bt, m->mu, 1);
        nomuts[i] = prob_branch(bt, m->mu, 0);
    }

    int *order = (int *) malloc(sizeof(int) * V);
    int *queue = (int *) malloc(sizeof(int) * V);
    tree_postorder(t, order);

    lk_row *inner = (lk_row *) malloc(sizeof(lk_row) * V);
    lk_row *outer = (lk_row *) malloc(sizeof(lk_row) * V);
    lk_row inner_sub[1];

    /* tree lengths, emit.cpp:744-774 */
    int top = 0;
    double maintreelen = 0.0;
    queue[top++] = maintree_root;
    while (top > 0) {
        int node = queue[--top];
        if (node != maintree_root)
            maintreelen += fmax(tree_dist(t, node, times), mintime);
        if (!tree_is_leaf(t, node)) {
            queue[top++] = t->c0[node];
            queue[top++] = t->c1[node];
        }
    }
    double subtreelen = 0.0;
    if (internal) {
        queue[top++] = subtree_root;
        while (top > 0) {
            int node = queue[--top];
            if (node != subtree_root)
                subtreelen += fmax(tree_dist(t, node, times), mintime);
            if (!tree_is_leaf(t, node)) {
                queue[top++] = t->c0[node];
                queue[top++] = t->c1[node];
            }
        }
    }

    /* per-state constants, emit.cpp:778-819 */
    double *smut = (double *) malloc(sizeof(double) * 6 * IMAX(nstates, 1));
    double *sinv = (double *) malloc(sizeof(double) * IMAX(nstates, 1));
    const int newleaf = (V + 1) / 2;
    for (int j = 0; j < nstates; j++) {
        int node1 = internal ? subtree_root : 0;
        int node2 = states[2 * j];
        int parent = t->parent[node2];
        double time1 = internal ? times[t->age[node1]] : 0.0;
        double time2 = times[t->age[node2]];
        double parent_time = (parent != -1) ?
            times[IMIN(t->age[parent], m->ntimes - 1)] : 0.0;
        double coal_time = times[states[2 * j + 1]];
        double dist[3];
        dist[0] = fmax(coal_time - time1, mintime);
        dist[1] = fmax(coal_time - time2, mintime);
        dist[2] = fmax(parent_time - coal_time, mintime);
        for (int q = 0; q < 3; q++) {
            smut[6 * j + q] = prob_branch(dist[q], m->mu, 1);
            smut[6 * j + 3 + q] = prob_branch(dist[q], m->mu, 0);
        }
        double treelen;
        if (node2 == maintree_root)
            treelen = maintreelen + subtreelen
                + fmax(coal_time - time1, mintime)
                + fmax(coal_time - times[t->age[maintree_root]], mintime);
        else
            treelen = maintreelen + subtreelen
                + fmax(coal_time - time1, mintime);
        sinv[j] = .25 * exp(-m->mu * fmax(treelen, mintime));
    }

    for (int i = 0; i < seqlen; i++) {
        double *row = emit + (size_t) i * nstates;
        if (masked[i]) {
            for (int j = 0; j < nstates; j++) row[j] = 1.0;
            continue;
        }
        if (invariant[i]) {
            for (int j = 0; j < nstates; j++) row[j] = sinv[j];
            continue;
        }

        /* inner: emit.cpp:151-196, postorder */
        for (int q = 0; q < V; q++) {
            int j = order[q];
            if (tree_is_leaf(t, j)) {
                leaf_row(seqs[j][i], inner[j]);
            } else {
                int c1 = t->c0[j], c2 = t->c1[j];
                for (int a = 0; a < 4; a++) {
                    double p1 = 0.0, p2 = 0.0;
                    for (int b = 0; b < 4; b++) {
                        if (a == b) {
                            p1 += inner[c1][b] * nomuts[c1];
                            p2 += inner[c2][b] * nomuts[c2];
                        } else {
                            p1 += inner[c1][b] * muts[c1];
                            p2 += inner[c2][b] * muts[c2];
                        }
                    }
                    inner[j][a] = p1 * p2;
                }
            }
        }
        /* outer: emit.cpp:200-296, preorder from the maintree root */
        top = 0;
        queue[top++] = maintree_root;
        while (top > 0) {
            int j = queue[--top];
            if (j == maintree_root) {
                outer[j][0] = outer[j][1] = outer[j][2] = outer[j][3] = 1.0;
            } else {
                int sib = tree_sibling(t, j);
                int parent = t->parent[j];
                if (parent != maintree_root) {
                    for (int a = 0; a < 4; a++) {
                        double p1 = 0.0, p2 = 0.0;
                        for (int b = 0; b < 4; b++) {
                            if (a == b) {
                                p1 += inner[sib][b] * nomuts[sib];
                                p2 += outer[parent][b] * nomuts[parent];
                            } else {
                                p1 += inner[sib][b] * muts[sib];
                                p2 += outer[parent][b] * muts[parent];
                            }
                        }
                        outer[j][a] = p1 * p2;
                    }
                } else {
                    for (int a = 0; a < 4; a++) {
                        double p1 = 0.0;
                        for (int b = 0; b < 4; b++) {
                            if (a == b)
                                p1 += inner[sib][b] * nomuts[sib];
                            else
                                p1 += inner[sib][b] * muts[sib];
                        }
                        outer[j][a] = p1;
                    }
                }
            }
            if (!tree_is_leaf(t, j)) {
                queue[top++] = t->c0[j];
                queue[top++] = t->c1[j];
            }
        }
        if (!internal)
            leaf_row(seqs[newleaf][i], inner_sub[0]);

        for (int j = 0; j < nstates; j++) {
            int node1 = internal ? subtree_root : 0;
            row[j] = calc_emit(inner, outer, internal ? inner : inner_sub,
                               node1, states[2 * j], maintree_root,
                               &smut[6 * j + 3], &smut[6 * j]);
        }
    }

    free(invariant); free(masked); free(muts); free(nomuts); free(order);
    free(queue); free(inner); free(outer); free(smut); free(sinv);
}

/* ---------------------------------------------------------------- results */

static int count_states_upper(const orc_problem *p)
{
    /* every node contributes at most ntimes states */
    return p->nnodes * p->ntimes;
}

orc_result *orc_result_new(const orc_problem *p)
{
    orc_result *r = (orc_result *) calloc(1, sizeof(orc_result));
    const int B = p->ntrees, T = p->ntimes;
    r->nstates = (int *) calloc(B, sizeof(int));
    r->state_off = (int64_t *) calloc(B + 1, sizeof(int64_t));
    r->row_off = (int64_t *) calloc(B + 1, sizeof(int64_t));
    r->fw_off = (int64_t *) calloc(B + 1, sizeof(int64_t));
    r->sw1_off = (int64_t *) calloc(B + 1, sizeof(int64_t));
    r->nbranches = (int *) calloc((size_t) B * T, sizeof(int));
    r->nrecombs = (int *) calloc((size_t) B * T, sizeof(int));
    r->ncoals = (int *) calloc((size_t) B * T, sizeof(int));
    for (int k = 0; k < 9; k++)
        r->tm[k] = (double *) calloc((size_t) B * T, sizeof(double));
    r->tm_minage = (int *) calloc(B, sizeof(int));
    r->sw_recombsrc = (int *) calloc(B, sizeof(int));
    r->sw_recoalsrc = (int *) calloc(B, sizeof(int));
    int n = 0;
    for (int b = 0; b < B; b++)
        n += p->blocklens[b];
    r->path = (int *) calloc(n, sizeof(int));
    r->first_bad_site = -1;
    return r;
}

void orc_result_free(orc_result *r)
{
    if (!r) return;
    free(r->nstates); free(r->state_off); free(r->states); free(r->row_off);
    free(r->fw_off); free(r->sw1_off); free(r->nbranches); free(r->nrecombs);
    free(r->ncoals);
    for (int k = 0; k < 9; k++) free(r->tm[k]);
    free(r->tm_minage); free(r->sw_determ); free(r->sw_determprob);
    free(r->sw_recombrow); free(r->sw_recoalrow); free(r->sw_recombsrc);
    free(r->sw_recoalsrc); free(r->emit); free(r->fw); free(r->path);
    free(r);
}

static void model_init(omodel *m, const orc_problem *p)
{
    m->ntimes = p->ntimes;
    m->times = p->times;
    m->popsizes = p->popsizes;
    m->rho = p->rho;
    m->mu = p->mu;
    m->time_steps = (double *) malloc(sizeof(double) * p->ntimes);
    m->coal_time_steps = (double *) malloc(sizeof(double) * 2 * p->ntimes);
    orc_time_steps(p->times, p->ntimes, m->time_steps, m->coal_time_steps);
    m->mintime = p->times[1] * .1;
    m->removed_root_time = p->ntimes + 1;
}

static void model_free(omodel *m)
{
    free(m->time_steps);
    free(m->coal_time_steps);
}

static void problem_tree(const orc_problem *p, int b, otree *t)
{
    const int V = p->nnodes;
    tree_init(t, V, p->ptrees + (size_t) b * V, p->ages + (size_t) b * V,
              (p->internal && p->subtree_roots) ? p->subtree_roots[b] : -1);
}

/* matrices.cpp:8-156: per-block states, emissions, switch, transitions */
void orc_setup(const orc_problem *p, orc_result *r)
{
    const int B = p->ntrees, T = p->ntimes, V = p->nnodes;
    omodel m;
    model_init(&m, p);

    /* pass 1: states */
    int *tmp = (int *) malloc(sizeof(int) * 2 * count_states_upper(p));
    int64_t total = 0;
    int *all = NULL;
    size_t cap = 0;
    int64_t n = 0;
    for (int b = 0; b < B; b++) {
        otree t;
        problem_tree(p, b, &t);
        int S = get_states(&t, T, p->internal, p->minage, tmp);
        if ((size_t) (total + S) * 2 > cap) {
            cap = (size_t) (total + S) * 4 + 1024;
            all = (int *) realloc(all, cap * sizeof(int));
        }
        memcpy(all + 2 * total, tmp, sizeof(int) * 2 * S);
        r->nstates[b] = S;
        r->state_off[b] = total;
        total += S;
        r->row_off[b + 1] = r->row_off[b] + IMAX(S, 1);
        r->fw_off[b + 1] = r->fw_off[b] + (int64_t) IMAX(S, 1) * p->blocklens[b];
        n += p->blocklens[b];
        tree_free(&t);
    }
    /* sw1_off[b] is the offset of block b's determ row (sized by block b-1) */
    r->sw1_off[0] = 0;
    r->sw1_off[1] = 0;
    for (int b = 1; b < B; b++)
        r->sw1_off[b + 1] = r->sw1_off[b] + IMAX(r->nstates[b - 1], 1);
    r->state_off[B] = total;
    r->states = all ? all : (int *) calloc(2, sizeof(int));
    free(tmp);

    r->sw_determ = (int *) calloc(r->sw1_off[B] + 1, sizeof(int));
    r->sw_determprob = (double *) calloc(r->sw1_off[B] + 1, sizeof(double));
    r->sw_recombrow = (double *) calloc(r->row_off[B], sizeof(double));
    r->sw_recoalrow = (double *) calloc(r->row_off[B], sizeof(double));
    r->emit = (double *) calloc(r->fw_off[B], sizeof(double));
    r->fw = (double *) calloc(r->fw_off[B], sizeof(double));

    /* pass 2: matrices */
    const unsigned char **rows = (const unsigned char **)
        malloc(sizeof(unsigned char *) * (p->nleaves + 1));
    int pos = 0; /* offset from start_coord; seqs are indexed from start_coord */
    otree last;
    int have_last = 0;
    for (int b = 0; b < B; b++) {
        otree t;
        problem_tree(p, b, &t);
        const int S = r->nstates[b];
        const int *states = r->states + 2 * r->state_off[b];
        int *nbr = r->nbranches + (size_t) b * T;
        int *nrc = r->nrecombs + (size_t) b * T;
        int *ncl = r->ncoals + (size_t) b * T;

        /* emissions (matrices.cpp:28-51 / :106-124) */
        int nseqs;
        for (int i = 0; i < p->nleaves; i++)
            rows[i] = p->seqs + (size_t) p->seqids[i] * p->seqlen +
                p->start_coord + pos;
        if (p->internal) {
            nseqs = p->nleaves;
        } else {
            rows[p->nleaves] = p->seqs + (size_t) p->new_chrom * p->seqlen +
                p->start_coord + pos;
            nseqs = p->nleaves + 1;
        }
        calc_emissions(states, S, &t, rows, nseqs, p->blocklens[b], &m,
                       p->internal, r->emit + r->fw_off[b]);

        /* switch (matrices.cpp:54-74 / :127-147) */
        r->sw_recombsrc[b] = -1;
        r->sw_recoalsrc[b] = -1;
        if (have_last) {
            const int S1 = r->nstates[b - 1];
            const int *states1 = r->states + 2 * r->state_off[b - 1];
            int *lb = (int *) malloc(sizeof(int) * 3 * T);
            count_lineages(&last, T, p->internal, lb, lb + T, lb + 2 * T);
            ospr spr = { p->sprs[4 * b], p->sprs[4 * b + 1], p->sprs[4 * b + 2],
                         p->sprs[4 * b + 3] };
            oswitch sw;
            sw.determ = r->sw_determ + r->sw1_off[b];
            sw.determprob = r->sw_determprob + r->sw1_off[b];
            sw.recombrow = r->sw_recombrow + r->row_off[b];
            sw.recoalrow = r->sw_recoalrow + r->row_off[b];
            calc_transition_probs_switch(
                &t, &last, &spr, p->mappings + (size_t) b * V, states1, S1,
                states, S, &m, lb, lb + T, lb + 2 * T, &sw, p->internal);
            r->sw_recombsrc[b] = sw.recombsrc;
            r->sw_recoalsrc[b] = sw.recoalsrc;
            free(lb);
            tree_free(&last);
        }

        /* transitions (matrices.cpp:76-82 / :149-155) */
        count_lineages(&t, T, p->internal, nbr, nrc, ncl);
        double *tmv[9];
        for (int k = 0; k < 9; k++)
            tmv[k] = r->tm[k] + (size_t) b * T;
        calc_transition_probs(&t, &m, nbr, nrc, ncl, p->internal, p->minage,
                              tmv, &r->tm_minage[b]);

        last = t;
        have_last = 1;
        pos += p->blocklens[b];
    }
    if (have_last)
        tree_free(&last);
    free(rows);
    model_free(&m);
}

/* sample_thread.cpp:186-296 */
static void forward_block(const otree *t, int ntimes, int blocklen,
                          const int *states, int nstates,
                          const double *const tm[9], int internal,
                          int tm_minage, const double *emit, double *fw,
                          double *logZ)
{
    int minage = tm_minage;
    int maintree_root = 0;
    if (internal) {
        maintree_root = t->c1[t->root];
        if (nstates == 0) {
            for (int i = 1; i < blocklen; i++)
                fw[i] = fw[i - 1];
            return;
        }
    }

    double *tmatrix = (double *) calloc((size_t) ntimes * ntimes, sizeof(double));
    double *tmatrix2 = (double *) calloc((size_t) ntimes * nstates, sizeof(double));
    for (int a = 0; a < ntimes - 1; a++) {
        for (int b = 0; b < ntimes - 1; b++)
            tmatrix[a * ntimes + b] = orc_get_time(tm, a, b, 0, minage, 0);
        for (int k = 0; k < nstates; k++) {
            const int b = states[2 * k + 1];
            const int node2 = states[2 * k];
            const int c = t->age[node2];
            tmatrix2[(size_t) a * nstates + k] =
                orc_get_time(tm, a, b, c, minage, 1) -
                orc_get_time(tm, a, b, 0, minage, 0);
        }
    }

    int maxtime = 0;
    for (int k = 0; k < nstates; k++)
        if (maxtime < states[2 * k + 1])
            maxtime = states[2 * k + 1];

    olookup L;
    lookup_init(&L, states, nstates, t->V);
    int *ages1 = (int *) malloc(sizeof(int) * t->V);
    int *ages2 = (int *) malloc(sizeof(int) * t->V);
    int *indexes = (int *) malloc(sizeof(int) * t->V);
    for (int i = 0; i < t->V; i++) {
        ages1[i] = IMAX(t->age[i], minage);
        indexes[i] = lookup(&L, i, ages1[i]);
        if (internal)
            ages2[i] = (i == maintree_root || i == t->root) ?
                maxtime : t->age[t->parent[i]];
        else
            ages2[i] = (i == t->root) ? maxtime : t->age[t->parent[i]];
    }

    double *tmatrix_fgroups = (double *) malloc(sizeof(double) * ntimes);
    double *fgroups = (double *) malloc(sizeof(double) * ntimes);
    for (int i = 1; i < blocklen; i++) {
        const double *col1 = fw + (size_t) (i - 1) * nstates;
        double *col2 = fw + (size_t) i * nstates;
        const double *emit2 = emit + (size_t) i * nstates;

        for (int a = 0; a < ntimes; a++)
            fgroups[a] = 0.0;
        for (int j = 0; j < nstates; j++)
            fgroups[states[2 * j + 1]] += col1[j];

        for (int b = 0; b < ntimes - 1; b++) {
            double sum = 0.0;
            for (int a = 0; a < ntimes - 1; a++)
                sum += tmatrix[a * ntimes + b] * fgroups[a];
            tmatrix_fgroups[b] = sum;
        }

        double norm = 0.0;
        for (int k = 0; k < nstates; k++) {
            const int b = states[2 * k + 1];
            const int node2 = states[2 * k];
            const int age1 = ages1[node2];
            const int age2 = ages2[node2];
            double sum = tmatrix_fgroups[b];
            const int j1 = indexes[node2];
            for (int j = j1, a = age1; a <= age2; j++, a++)
                sum += tmatrix2[(size_t) a * nstates + k] * col1[j];
            col2[k] = sum * emit2[k];
            norm += col2[k];
        }
        for (int k = 0; k < nstates; k++)
            col2[k] /= norm;
        *logZ += log(norm);
    }

    free(tmatrix); free(tmatrix2); free(ages1); free(ages2); free(indexes);
    free(tmatrix_fgroups); free(fgroups);
    lookup_free(&L);
}

/* sample_thread.cpp:345-389 */
static void forward_switch(const double *col1, double *col2, const oswitch *sw,
                           const double *emit, double *logZ)
{
    const int nstates1 = IMAX(sw->nstates1, 1);
    const int nstates2 = IMAX(sw->nstates2, 1);
    for (int k = 0; k < nstates2; k++)
        col2[k] = 0.0;
    for (int j = 0; j < nstates1; j++) {
        int k = sw->determ[j];
        if (j != sw->recombsrc && j != sw->recoalsrc && k != -1)
            col2[k] += col1[j] * sw->determprob[j];
    }
    double norm = 0.0;
    for (int k = 0; k < nstates2; k++) {
        if (sw->recombsrc != -1 && sw->recombrow[k] > 0.0)
            col2[k] += col1[sw->recombsrc] * sw->recombrow[k];
        if (sw->recoalsrc != -1 && sw->recoalrow[k] > 0.0)
            col2[k] += col1[sw->recoalsrc] * sw->recoalrow[k];
        col2[k] *= emit[k];
        norm += col2[k];
    }
    for (int k = 0; k < nstates2; k++)
        col2[k] /= norm;
    *logZ += log(norm);
}

static void result_switch(const orc_result *r, int b, oswitch *sw)
{
    sw->nstates1 = r->nstates[b - 1];
    sw->nstates2 = r->nstates[b];
    sw->determ = r->sw_determ + r->sw1_off[b];
    sw->determprob = r->sw_determprob + r->sw1_off[b];
    sw->recombrow = r->sw_recombrow + r->row_off[b];
    sw->recoalrow = r->sw_recoalrow + r->row_off[b];
    sw->recombsrc = r->sw_recombsrc[b];
    sw->recoalsrc = r->sw_recoalsrc[b];
}

/* sample_thread.cpp:394-460 */
void orc_forward(const orc_problem *p, orc_result *r, const double *prior)
{
    const int B = p->ntrees, T = p->ntimes;
    omodel m;
    model_init(&m, p);
    r->logZ = 0.0;
    r->first_bad_site = -1;
    int pos = 0;
    for (int b = 0; b < B; b++) {
        otree t;
        problem_tree(p, b, &t);
        const int S = r->nstates[b];
        const int S1 = IMAX(S, 1);
        const int *states = r->states + 2 * r->state_off[b];
        double *fw = r->fw + r->fw_off[b];
        const double *emit = r->emit + r->fw_off[b];
        const double *tmv[9];
        for (int k = 0; k < 9; k++)
            tmv[k] = r->tm[k] + (size_t) b * T;

        if (b == 0) {
            if (prior) {
                memcpy(fw, prior, sizeof(double) * S1);
            } else if (S == 0) {
                fw[0] = 1.0;                       /* trans.cpp:822-825 */
            } else {
                /* StatesModel.minage, not the matrix minage (sample_thread.cpp:428) */
                for (int j = 0; j < S; j++)
                    fw[j] = state_prior(states[2 * j + 1],
                                        r->nbranches + (size_t) b * T,
                                        r->ncoals + (size_t) b * T, p->minage,
                                        m.popsizes, m.coal_time_steps, T);
            }
        } else {
            oswitch sw;
            result_switch(r, b, &sw);
            const double *col1 = r->fw + r->fw_off[b] -
                IMAX(r->nstates[b - 1], 1);
            forward_switch(col1, fw, &sw, emit, &r->logZ);
        }

        double top = fw[0];
        for (int j = 1; j < S1; j++)
            if (fw[j] > top) top = fw[j];
        if (!(top > 0.0) && r->first_bad_site < 0)
            r->first_bad_site = pos;

        forward_block(&t, T, p->blocklens[b], states, S, tmv, p->internal,
                      r->tm_minage[b], emit, fw, &r->logZ);

        const double *lastcol = fw + (size_t) (p->blocklens[b] - 1) * S1;
        top = lastcol[0];
        for (int j = 1; j < S1; j++)
            if (lastcol[j] > top) top = lastcol[j];
        if (!(top > 0.0) && r->first_bad_site < 0)
            r->first_bad_site = pos + p->blocklens[b] - 1;

        pos += p->blocklens[b];
        tree_free(&t);
    }
    model_free(&m);
}

/* common.h:272-290 with frand(total) = rand()/double(RAND_MAX)*total (:106) */
static int sample_weights(const double *weights, int n, int r, int rand_max)
{
    double total = 0.0;
    for (int i = 0; i < n; i++)
        total += weights[i];
    double pick = r / (double) rand_max * total;
    double x = 0.0;
    for (int i = 0; i < n; i++) {
        x += weights[i];
        if (x >= pick)
            return i;
    }
    return n - 1;
}

/* sample_thread.cpp:470-569 */
int orc_traceback(const orc_problem *p, orc_result *r, const int *rand_ints,
                  int rand_max, int last_state)
{
    const int B = p->ntrees, T = p->ntimes;
    int n = 0;
    for (int b = 0; b < B; b++)
        n += p->blocklens[b];
    int *path = r->path;
    int used = 0;
    int pos = n;

    int maxS = 1;
    for (int b = 0; b < B; b++)
        maxS = IMAX(maxS, r->nstates[b]);
    double *A = (double *) malloc(sizeof(double) * maxS);
    double *trans = (double *) malloc(sizeof(double) * maxS);

    if (last_state < 0) {
        const int S1 = IMAX(r->nstates[B - 1], 1);
        const double *col = r->fw + r->fw_off[B] - S1;
        path[pos - 1] = sample_weights(col, S1, rand_ints[used++], rand_max);
    } else {
        path[pos - 1] = last_state;
    }

    for (int b = B - 1; b >= 0; b--) {
        otree t;
        problem_tree(p, b, &t);
        const int S = r->nstates[b];
        const int S1 = IMAX(S, 1);
        const int *states = r->states + 2 * r->state_off[b];
        const double *tmv[9];
        for (int k = 0; k < 9; k++)
            tmv[k] = r->tm[k] + (size_t) b * T;
        const int blocklen = p->blocklens[b];
        pos -= blocklen;
        const double *fw = r->fw + r->fw_off[b];

        /* sample_hmm_posterior, :470-503 */
        int last_k = -1;
        for (int i = blocklen - 2; i >= 0; i--) {
            int k = path[pos + i + 1];
            if (k != last_k) {
                for (int j = 0; j < S1; j++)
                    trans[j] = (S == 0) ? 1.0 :
                        trans_get(tmv, &t, states, S, p->internal, j, k);
                last_k = k;
            }
            for (int j = 0; j < S1; j++)
                A[j] = fw[(size_t) i * S1 + j] * trans[j];
            path[pos + i] = sample_weights(A, S1, rand_ints[used++], rand_max);
        }

        /* sample_hmm_posterior_step, :506-519 */
        if (b > 0) {
            oswitch sw;
            result_switch(r, b, &sw);
            const int n1 = IMAX(sw.nstates1, 1);
            const double *col1 = r->fw + r->fw_off[b] - n1;
            for (int j = 0; j < n1; j++)
                A[j] = col1[j] * switch_get(&sw, j, path[pos]);
            path[pos - 1] = sample_weights(A, n1, rand_ints[used++], rand_max);
        }
        tree_free(&t);
    }
    free(A);
    free(trans);
    return used;
}

int orc_thread_sample(const orc_problem *p, orc_result *r, const int *rand_ints,
                      int rand_max)
{
    orc_setup(p, r);
    orc_forward(p, r, NULL);
    return orc_traceback(p, r, rand_ints, rand_max, -1);
}

/* ------------------------------------------------------ generic dense HMM */

/* common.h:182-202 */
static double logsum(const double *vals, int nvals)
{
    const double threshold = -15;
    if (nvals == 0)
        return 1.0;
    double maxval = vals[0];
    for (int i = 1; i < nvals; i++)
        if (vals[i] > maxval)
            maxval = vals[i];
    double expsum = 0.0;
    for (int i = 0; i < nvals; i++)
        if (vals[i] - maxval > threshold)
            expsum += exp(vals[i] - maxval);
    return maxval + log(expsum);
}

/* hmm.cpp:27-43 */
void orc_hmm_forward_alg(int n, int nstates, const double *trans,
                         const double *emit, double *fw)
{
    double *vec = (double *) malloc(sizeof(double) * nstates);
    for (int i = 1; i < n; i++) {
        const double *col1 = fw + (size_t) (i - 1) * nstates;
        double *col2 = fw + (size_t) i * nstates;
        const double *emit2 = emit + (size_t) i * nstates;
        for (int k = 0; k < nstates; k++) {
            for (int j = 0; j < nstates; j++)
                vec[j] = col1[j] + trans[(size_t) j * nstates + k];
            col2[k] = logsum(vec, nstates) + emit2[k];
        }
    }
    free(vec);
}
