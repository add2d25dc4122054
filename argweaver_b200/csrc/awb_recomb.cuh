// awb_recomb.cuh -- sampling the recombination points of a thread path on the
// device: the step right after the traceback.
//
// Replaces sample_recombinations (recomb.cpp:144-235) with
// recomb_prob_unnormalized (:14-117) and get_possible_recomb (:122-141).  The
// reference walks the path site by site and rebuilds every block's matrices a
// third time to do so; here the per-block tables of the setup kernels are still
// in HBM and the path never leaves the device between traceback and this step.
//
// Random numbers.  The reference draws from libc rand(): one draw per sampled
// waiting time (expovariate, common.h:124) and one per sampled event (sample,
// common.h:272-290) -- a data-dependent number of draws, so they cannot be
// shipped in advance the way the traceback's are.  The kernel therefore runs
// glibc's own generator (the TYPE_3 additive feedback generator behind rand():
// r[i] += r[i-3 mod 31], result = r[i] >> 1; glibc random_r.c) from a snapshot
// of the caller's state and reports how many draws it took; the caller advances
// its libc stream by that many (awb_libc_rand_snapshot / awb_libc_rand_advance in
// awb_api.cu), which leaves the stream exactly where the reference would.
//
// One warp per window: the walk is sequential in the random stream, so every
// lane runs the same scalar code (lane 0 writes the output); the only part with
// any volume -- skipping over the sites at which the path stays put and no
// recombination is due, ~99.9 % of them -- is done 32 sites at a time with a
// coalesced load and a ballot.
#ifndef AWB_RECOMB_CUH
#define AWB_RECOMB_CUH

#include "awb_common.cuh"

struct AwbRng {
    int r[31];
    int f, b;                 // front / rear index: r[f] += r[b]
};

AWB_HD inline unsigned awb_rng_next(AwbRng &g)
{
    const unsigned v = (unsigned) g.r[g.f] + (unsigned) g.r[g.b];
    g.r[g.f] = (int) v;
    if (++g.f >= 31) {
        g.f = 0;
        ++g.b;
    } else if (++g.b >= 31) {
        g.b = 0;
    }
    return v >> 1;
}

// recomb.cpp:14-117
AWB_HD inline double awb_recomb_prob(const AwbModel &m, const int *parent,
                                     const int *age, int root, bool internal,
                                     const int *nbranches, const int *nrecombs,
                                     const int *ncoals, int last_node, int last_time,
                                     int state_time, int rnode, int k)
{
    const int j = state_time;
    int root_time, recomb_parent_age;
    if (internal) {
        // child[0] / child[1] of the virtual root: the caller passes them packed
        // into `root` as (subtree_root << 16 | maintree_root)
        const int subtree_root = root >> 16, maintree_root = root & 0xffff;
        root_time = awb_imax(age[maintree_root], last_time);
        recomb_parent_age = (rnode == subtree_root || parent[rnode] == -1 ||
                             rnode == last_node) ? last_time : age[parent[rnode]];
    } else {
        root_time = awb_imax(age[root], last_time);
        recomb_parent_age = (rnode == -1 || parent[rnode] == -1 ||
                             rnode == last_node) ? last_time : age[parent[rnode]];
    }
    const int nbranches_k = nbranches[k] + (k < last_time ? 1 : 0);
    const int nrecombs_k = nrecombs[k] + (k <= last_time ? 1 : 0) +
        (k == last_time ? 1 : 0) - (k == root_time ? 1 : 0);
    const double precomb = nbranches_k * m.time_steps[k] / nrecombs_k;

    // probability of not coalescing before time j-1
    double sum = 0.0;
    for (int x = k; x < j - 1; x++) {
        const int nb = nbranches[x] + (x < last_time ? 1 : 0) -
            (x < recomb_parent_age ? 1 : 0);
        sum += (m.time_steps[x] * nb / (2.0 * m.popsizes[x]));
    }
    // probability of coalescing at time j
    double pcoal = 1.0;
    const int nbranches_j = nbranches[j] + (j < last_time ? 1 : 0) -
        (j < recomb_parent_age ? 1 : 0);
    if (k == j) {
        if (j < m.ntimes - 2)
            pcoal = 1.0 - exp(-m.coal_time_steps[2 * j] * nbranches_j /
                              (2.0 * m.popsizes[j]));
    } else {
        const int x = j - 1;
        const int nbranches_m = nbranches[x] + (x < last_time ? 1 : 0) -
            (x < recomb_parent_age ? 1 : 0);
        sum += (m.coal_time_steps[2 * x] * nbranches_m / (2.0 * m.popsizes[x]));
        if (j < m.ntimes - 2)
            pcoal = 1.0 - exp(-m.coal_time_steps[2 * j - 1] * nbranches_m /
                              (2.0 * m.popsizes[x]) -
                              m.coal_time_steps[2 * j] * nbranches_j /
                              (2.0 * m.popsizes[j]));
    }
    const int ncoals_j = ncoals[j] - (j <= recomb_parent_age ? 1 : 0) -
        (j == recomb_parent_age ? 1 : 0) + (j <= last_time ? 1 : 0) +
        (j == last_time ? 1 : 0);
    pcoal /= ncoals_j;
    return precomb * exp(-sum) * pcoal;
}

// The walk over one window.  out_pos/out_node/out_time[cap]; info[0] = number of
// recombinations (the true count, also when it exceeds cap), info[1] = draws.
// lane: 0..31 on the device (whole warp calls), -1 for a single host thread
AWB_HD inline void awb_sample_recombs(const AwbChain &ch, AwbRng &rng, int rand_max,
                                      int *out_pos, int *out_node, int *out_time,
                                      int cap, int *info, int lane)
{
    const AwbModel &m = ch.model;
    const int T = m.ntimes, V = ch.nnodes;
    const bool internal = ch.internal != 0;
    const int *path = ch.path;
    int nrec = 0, draws = 0;
    double probs[2 * AWB_MAXT + 2];
    short cnode[2 * AWB_MAXT + 2];
    signed char ctime[2 * AWB_MAXT + 2];

    for (int b = 0; b < ch.ntrees; b++) {
        const int S = ch.nstates[b];
        if (internal && S == 0)
            continue;                               // recomb.cpp:166-168
        const int *parent = ch.ptrees + (size_t) b * V;
        const int *age = ch.ages + (size_t) b * V;
        const int *nbranches = ch.lineages + (size_t) b * 3 * T;
        const int *nrecombs = nbranches + T, *ncoals = nrecombs + T;
        const double *tv = ch.tmvec + (size_t) b * AWB_TM_NVEC * T;
        const long long row0 = ch.row_off[b];
        const int root = ch.root[b];
        int subtree_root = -1, rootarg = root, get_minage = 0;
        if (internal) {
            subtree_root = ch.child0[(size_t) b * V + root];
            rootarg = (subtree_root << 16) | (int) ch.child1[(size_t) b * V + root];
            get_minage = age[subtree_root];         // TransMatrix::get, trans.h:64-83
        }
        // no new recombination at the first site of a block (recomb.cpp:172-175:
        // every block here starts the window or follows a breakpoint)
        const int end = ch.block_start[b + 1];
        int next_recomb = -1;
        for (int i = ch.block_start[b] + 1; i < end; i++) {
            const int cur = path[i], last = path[i - 1];
            if (cur == last) {
                if (i > next_recomb) {
                    // waiting time to the next (invisible) recombination
                    const int a = ch.st_time[row0 + last];
                    const int c = age[ch.st_node[row0 + last]];
                    const double self = awb_get_time(tv, T, a, a, c, get_minage, true);
                    const double rate = 1.0 - (tv[AWB_TM_NORECOMBS * T + a] / self);
                    const double u = awb_rng_next(rng) / double(rand_max);
                    draws++;
                    next_recomb = int(fmin(double(end), i + (-log(u) / rate)));
                }
                if (i < next_recomb) {
                    // nothing is drawn until the next change of state or
                    // next_recomb, whichever comes first
                    int x = i + 1;
#ifdef __CUDA_ARCH__
                    const int lim = awb_imin(end, next_recomb);
                    for (;;) {
                        const int xx = x + lane;
                        const bool stop = xx < lim ? (path[xx] != last) : true;
                        const unsigned sm = __ballot_sync(0xffffffffu, stop);
                        if (sm) {
                            x += __ffs(sm) - 1;
                            break;
                        }
                        x += 32;
                    }
#else
                    while (x < end && x < next_recomb && path[x] == last)
                        x++;
#endif
                    i = x - 1;
                    continue;
                }
            }
            next_recomb = -1;
            const int node = ch.st_node[row0 + cur], time = ch.st_time[row0 + cur];
            const int lnode = ch.st_node[row0 + last], ltime = ch.st_time[row0 + last];
            // get_possible_recomb (recomb.cpp:122-141)
            int nc = 0;
            const int end_time = awb_imin(time, ltime);
            if (node == lnode)
                for (int k = age[node]; k <= end_time; k++) {
                    cnode[nc] = (short) node;
                    ctime[nc++] = (signed char) k;
                }
            if (internal) {
                for (int k = age[subtree_root]; k <= end_time; k++) {
                    cnode[nc] = (short) subtree_root;
                    ctime[nc++] = (signed char) k;
                }
            } else {
                for (int k = 0; k <= end_time; k++) {
                    cnode[nc] = -1;
                    ctime[nc++] = (signed char) k;
                }
            }
            double total = 0.0;
            for (int x = 0; x < nc; x++) {
                probs[x] = awb_recomb_prob(m, parent, age, rootarg, internal, nbranches,
                                           nrecombs, ncoals, lnode, ltime, time,
                                           cnode[x], ctime[x]);
                total += probs[x];
            }
            // sample (common.h:272-290)
            const double pick = awb_rng_next(rng) / double(rand_max) * total;
            draws++;
            int sel = nc - 1;
            double acc = 0.0;
            for (int x = 0; x < nc; x++) {
                acc += probs[x];
                if (acc >= pick) { sel = x; break; }
            }
            if (nrec < cap && lane <= 0) {
                out_pos[nrec] = i;
                out_node[nrec] = cnode[sel];
                out_time[nrec] = ctime[sel];
            }
            nrec++;
        }
    }
    if (lane <= 0) {
        info[0] = nrec;
        info[1] = draws;
    }
}

#endif // AWB_RECOMB_CUH
