// awb_setup.cuh -- per-block setup of the threading HMM.
//
//   awb_block_setup   (K1)  states, lineage counts, tree lengths, the 9 SMC
//                           transition vectors, the (T x T) time matrix, the
//                           same-branch band, invariant-site emissions, priors
//   awb_switch_setup  (K2)  switch matrix at a local-tree breakpoint, as a CSR
//                           gather list over the target states
//
// Both are written as __host__ __device__ workers over one block so the same
// code is unit-tested on the CPU (tests/host_emul.cpp).  On the GPU K1 runs one
// thread per block (awb_api.cu; it is bound by L1 sector throughput -- every
// access of a warp goes to 32 different lines -- which is why counters and loop
// invariants are kept in registers / local memory here) and K2 has a
// warp-per-breakpoint twin (awb_switch_setup_warp).  Together with the time
// matrices they are ~17 % of a step at the bench shape.
//
// Reference behaviour restated here (file:line in mdrasmus/argweaver):
//   states.cpp:53-74,109-166          local_tree.cpp:34-219
//   trans.cpp:13-115 (+ trans.h:89-128) trans.cpp:156-739, 785-831
//   emit.cpp:744-819                  sample_thread.cpp:209-247
#ifndef AWB_SETUP_CUH
#define AWB_SETUP_CUH

#include "awb_common.cuh"

// ---------------------------------------------------------------------------
// K1

AWB_HD inline double awb_state_prior(int b, const int *nbranches,
                                     const int *ncoals, int minage,
                                     const AwbModel &m)
{
    // trans.cpp:785-812
    if (b < minage)
        return 0.0;
    double sum = 0.0;
    for (int k = 2 * minage; k < 2 * b - 1; k++)
        sum += m.coal_time_steps[k] * nbranches[k / 2] /
            (2.0 * m.popsizes[k / 2]);
    double p = exp(-sum) / ncoals[b];
    if (b < m.ntimes - 2) {
        double Z = 0.0;
        if (b > minage)
            Z = m.coal_time_steps[2 * b - 1] * nbranches[b - 1] /
                (2.0 * m.popsizes[b - 1]);
        p *= 1.0 - exp(-m.coal_time_steps[2 * b] * nbranches[b] /
                       (2.0 * m.popsizes[b]) - Z);
    }
    return p;
}

// Returns 0 on success, a positive code on a layout mismatch.
// (T x T) time matrix of block b (sample_thread.cpp:212-216): entries
// first, first+step, ... (one thread per entry on the GPU).  Needs tmvec and
// tm_minage of the block (awb_block_setup).
AWB_HD inline void awb_tmatrix_fill(const AwbChain &ch, int b, int first, int step)
{
    const int T = ch.model.ntimes;
    if (ch.nstates[b] == 0)
        return;
    const double *tv = ch.tmvec + (size_t) b * AWB_TM_NVEC * T;
    const int minage = ch.tm_minage[b];
    double *tm = ch.tmatrix + (size_t) b * T * T;
    for (int x = first; x < (T - 1) * (T - 1); x += step) {
        const int a = x / (T - 1), bb = x - a * (T - 1);
        tm[a * T + bb] = awb_get_time(tv, T, a, bb, 0, minage, false);
    }
}

// VCAP: capacity of the per-node working arrays (>= nnodes).  They are the
// thread's own (local memory on the GPU): the worker walks parent / age /
// children many times, and where the global arrays cost a warp 32 cache lines
// an access (32 blocks, 32 rows), element i of 32 threads' local arrays is ONE
// line -- consecutive blocks are nearly the same tree, so the lanes of a warp
// ask for the same i nearly all the time.  The results (children, state ranges,
// post-order) are copied out once, packed.
template <int VCAP>
AWB_HD inline int awb_block_setup_t(const AwbChain &ch, int b)
{
    const AwbModel &m = ch.model;
    const int T = m.ntimes;
    const int V = ch.nnodes;
    const bool internal = ch.internal != 0;
    short parent[VCAP], c0[VCAP], c1[VCAP], nfirst[VCAP], ncnt[VCAP], order[VCAP];
    signed char age[VCAP];
    {
        const int *gp = ch.ptrees + (size_t) b * V;
        const int *ga = ch.ages + (size_t) b * V;
        // (two nodes a load; ptrees and ages sit at the same offset of arrays
        // aligned alike)
        int i = 0;
        if ((((size_t) gp) & 7) && V > 0) {
            parent[0] = (short) gp[0];
            age[0] = (signed char) ga[0];
            i = 1;
        }
        for (; i + 1 < V; i += 2) {
            int p0, p1, a0, a1;
            awb_load2i(gp + i, p0, p1);
            awb_load2i(ga + i, a0, a1);
            parent[i] = (short) p0;
            parent[i + 1] = (short) p1;
            age[i] = (signed char) a0;
            age[i + 1] = (signed char) a1;
        }
        if (i < V) {
            parent[i] = (short) gp[i];
            age[i] = (signed char) ga[i];
        }
    }
    const long long row0 = ch.row_off[b];
    if (ch.gen_mappings) {
        // make_node_mapping (local_tree.h:767-776): identity, except that the
        // parent of the SPR's recombination node (in the previous tree) is gone
        int *mp = ch.mappings + (size_t) b * V;
        for (int x = 0; x < V; x++)
            mp[x] = x;
        if (b > 0)
            mp[ch.ptrees[(size_t) (b - 1) * V + ch.sprs[4 * (size_t) b]]] = -1;
    }
    const int S = ch.nstates[b];

    // ---- children in node-index order (local_tree.h:188-229)
    int root = -1;
    for (int i = 0; i < V; i++) {
        c0[i] = -1;
        c1[i] = -1;
    }
    for (int i = 0; i < V; i++) {
        const int p = parent[i];
        if (p != -1) {
            if (c0[p] == -1) c0[p] = (short) i;
            else c1[p] = (short) i;
        } else {
            root = i;
        }
    }
    if (internal && ch.subtree_roots) {
        const int want = ch.subtree_roots[b];
        if (want >= 0 && c1[root] == want) {
            const short tmp = c0[root];
            c0[root] = c1[root];
            c1[root] = tmp;
        }
    }
    ch.root[b] = (short) root;

    // ---- post-order (local_tree.h:274-304); ncnt doubles as the visit counter.
    //      Leaves first, a parent as soon as both children are placed: the order
    //      is sorted by height above the leaves, and the level boundaries let
    //      the emission kernel work through a level in parallel (nfirst holds
    //      the heights here; it is filled with its real content further down).
    {
        short *lstart = ch.lstart + (size_t) b * (V + 2);
        int i;
        for (i = 0; i < V; i++)
            ncnt[i] = 0;
        for (i = 0; i < V; i++) {
            if (c0[i] != -1)
                break;
            order[i] = (short) i;
            nfirst[i] = 0;
        }
        int end = i;
        int nlev = 1;
        lstart[0] = 0;
        for (i = 0; i < V && i < end; i++) {
            const int p = parent[order[i]];
            if (p != -1) {
                ncnt[p]++;
                if (ncnt[p] == 2) {
                    const int h = nfirst[order[i]] + 1;   // the later child is the higher
                    nfirst[p] = (short) h;
                    if (h == nlev) {
                        lstart[nlev] = (short) end;
                        nlev++;
                    }
                    order[end++] = (short) p;
                }
            }
        }
        if (end != V)
            return 5;   // not a binary tree with leaves listed first
        lstart[nlev] = (short) V;
        lstart[V + 1] = (short) nlev;
    }

    const int subtree_root = internal ? c0[root] : root;
    const int maintree_root = internal ? c1[root] : root;
    const bool full_tree = internal && age[root] < T;  // no states (states.cpp:116)
    int minage = ch.minage;
    if (internal)
        minage = awb_imax(minage, age[subtree_root]);  // states.cpp:122, trans.cpp:69
    ch.tm_minage[b] = minage;

    // ---- tree lengths for the invariant-site emission (emit.cpp:744-774):
    //      explicit-stack preorder, child[0] pushed first, so child[1]'s
    //      subtree is summed first -- kept for bit-level agreement.
    //      (nfirst is used as the stack; it is filled in later.)
    double maintreelen = 0.0, subtreelen = 0.0;
    {
        int top = 0;
        nfirst[top++] = (short) maintree_root;
        while (top > 0) {
            const int node = nfirst[--top];
            if (node != maintree_root) {
                const int p = parent[node];
                const double d = (p != -1) ?
                    m.times[age[p]] - m.times[age[node]] : 0.0;
                maintreelen += fmax(d, m.mintime);
            }
            if (c0[node] != -1) {
                nfirst[top++] = c0[node];
                nfirst[top++] = c1[node];
            }
        }
        if (internal) {
            nfirst[top++] = (short) subtree_root;
            while (top > 0) {
                const int node = nfirst[--top];
                if (node != subtree_root) {
                    const int p = parent[node];
                    const double d = (p != -1) ?
                        m.times[age[p]] - m.times[age[node]] : 0.0;
                    subtreelen += fmax(d, m.mintime);
                }
                if (c0[node] != -1) {
                    nfirst[top++] = c0[node];
                    nfirst[top++] = c1[node];
                }
            }
        }
    }

    // ---- which nodes carry states (states.cpp:124-143): ncnt = -1 marks ignored
    for (int i = 0; i < V; i++)
        ncnt[i] = 0;
    if (internal) {
        ncnt[root] = -1;
        int top = 0;
        nfirst[top++] = (short) subtree_root;
        while (top > 0) {
            const int node = nfirst[--top];
            ncnt[node] = -1;
            if (c0[node] != -1) {
                nfirst[top++] = c0[node];
                nfirst[top++] = c1[node];
            }
        }
    }

    // ---- invariant-site emission (emit.cpp:786-819): it depends on the
    //      coalescence time and on whether the branch is the main tree's root
    //      branch, 2(T-1) values; the states loop below writes them per state
    double ie[2][AWB_MAXT];
    {
        const double time1 = internal ? m.times[age[subtree_root]] : 0.0;
        for (int bt = 0; bt < T - 1; bt++) {
            const double coal_time = m.times[bt];
            const double tl2 = maintreelen + subtreelen + fmax(coal_time - time1, m.mintime);
            const double tl2r = tl2 + fmax(coal_time - m.times[age[maintree_root]], m.mintime);
            ie[0][bt] = .25 * exp(-m.mu * fmax(tl2, m.mintime));
            ie[1][bt] = .25 * exp(-m.mu * fmax(tl2r, m.mintime));
        }
    }

    // ---- states, node-major (states.cpp:53-74 / :146-165)
    int ns = 0;
    {
        // (the state tables are written through packers: 8 bytes a store)
        AwbPacker<short> pk_node;
        AwbPacker<signed char> pk_time, pk_age;
        pk_node.init(ch.st_node, row0);
        pk_time.init(ch.st_time, row0);
        pk_age.init(ch.st_age, row0);
        double ie_held = 0.0;       // inv_emit of an even row waiting for its neighbour
        for (int i = 0; i < V; i++) {
            if (full_tree || ncnt[i] < 0) {
                nfirst[i] = -1;
                ncnt[i] = 0;
                continue;
            }
            const int p = parent[i];
            const int ai = age[i];
            const int lo = awb_imax(ai, minage);
            const int hi = (p == -1 || (internal && p == root)) ? T - 2 : age[p];
            const int cnt = hi - lo + 1;
            if (cnt <= 0) {
                nfirst[i] = -1;
                ncnt[i] = 0;
                continue;
            }
            nfirst[i] = (short) ns;
            ncnt[i] = (short) cnt;
            if (ns + cnt <= S) {
                const double *iev = ie[i == maintree_root ? 1 : 0];
                long long k = row0 + ns;
                for (int t = lo; t <= hi; t++, k++) {
                    pk_node.put((unsigned) i);
                    pk_time.put((unsigned) t);
                    pk_age.put((unsigned) ai);
                    if (k & 1) {
                        if (k > row0)
                            awb_store2(ch.inv_emit + k - 1, ie_held, iev[t]);
                        else
                            ch.inv_emit[k] = iev[t];
                    } else {
                        ie_held = iev[t];
                    }
                }
            }
            ns += cnt;
        }
        pk_node.flush();
        pk_time.flush();
        pk_age.flush();
        if (ns <= S && ns > 0 && ((row0 + ns) & 1))
            ch.inv_emit[row0 + ns - 1] = ie_held;      // an even last row
    }
    if (ns != S)
        return 1;   // host layout and device enumeration disagree
    {
        // the node tables the later kernels read
        AwbPacker<short> p0, p1, pf, pc, po;
        const long long o = (long long) b * V;
        p0.init(ch.child0, o);
        p1.init(ch.child1, o);
        pf.init(ch.node_first, o);
        pc.init(ch.node_cnt, o);
        po.init(ch.order, o);
        for (int i = 0; i < V; i++) {
            p0.put((unsigned) c0[i]);
            p1.put((unsigned) c1[i]);
            pf.put((unsigned) nfirst[i]);
            pc.put((unsigned) ncnt[i]);
            po.put((unsigned) order[i]);
        }
        p0.flush();
        p1.flush();
        pf.flush();
        pc.flush();
        po.flush();
    }
    if (S == 0) {
        ch.st_node[row0] = -1;
        ch.st_time[row0] = -1;
    }

    // ---- lineage counts (local_tree.cpp:34-69 / :82-131)
    //      counted in thread-local arrays (one thread per block on the GPU: the
    //      increments would be scattered read-modify-writes in global memory)
    //      A branch from age a to its parent's age pa adds one to nbranches on
    //      [a, pa) -- [a, pa] for the top branch -- and one to nrecombs and
    //      ncoals on [a, pa]: difference arrays, then prefix sums.
    int nbranches[AWB_MAXT + 1], nrecombs[AWB_MAXT + 1], ncoals[AWB_MAXT + 1];
    for (int i = 0; i <= T; i++)
        nbranches[i] = nrecombs[i] = 0;
    for (int i = 0; i < V; i++) {
        if (internal && (i == subtree_root || i == root))
            continue;
        const int p = parent[i];
        const bool top = internal ? (p == root) : (p == -1);
        const int pa = top ? T - 2 : age[p];
        const int a = awb_imin(age[i], pa);
        nbranches[a]++;
        nbranches[top ? pa + 1 : pa]--;
        nrecombs[a]++;
        nrecombs[pa + 1]--;
    }
    {
        int rb = 0, rr = 0;
        for (int i = 0; i < T; i++) {
            rb += nbranches[i];
            rr += nrecombs[i];
            nbranches[i] = rb;
            nrecombs[i] = rr;
            ncoals[i] = rr;
        }
    }
    if (internal) {
        const int sa = age[subtree_root];
        for (int i = 0; i < sa; i++) {
            nbranches[i]--;
            ncoals[i]--;
            nrecombs[i]--;
        }
    }
    nbranches[T - 1] = 1;
    {
        AwbPacker<int> pb, pr, pc;
        const long long lg = (long long) b * 3 * T;
        pb.init(ch.lineages, lg);
        pr.init(ch.lineages, lg + T);
        pc.init(ch.lineages, lg + 2 * T);
        for (int i = 0; i < T; i++) {
            pb.put((unsigned) nbranches[i]);
            pr.put((unsigned) nrecombs[i]);
            pc.put((unsigned) ncoals[i]);
        }
        pb.flush();
        pr.flush();
        pc.flush();
    }

    // ---- tree length (local_tree.cpp:136-176), summed in node order
    double treelen = 0.0;
    for (int i = 0; i < V; i++) {
        const int p = parent[i];
        if (p == -1 || (internal && p == root))
            continue;
        treelen += m.times[age[p]] - m.times[age[i]];
    }
    ch.treelen[b] = treelen;

    // ---- transition vectors (trans.cpp:26-115)
    double *tv = ch.tmvec + (size_t) b * AWB_TM_NVEC * T;
    {
        int root_age_index;
        double root_age, tl;
        if (internal) {
            root_age_index = age[maintree_root];
            root_age = m.times[root_age_index];
            tl = treelen - m.times[age[subtree_root]];
        } else {
            root_age_index = age[root];
            root_age = m.times[root_age_index];
            tl = treelen;
        }
        // running cumulative coalescent rates C[2b-2], C[2b-1] (trans.cpp:44-54)
        double Cm2 = 0.0, Cm1 = 0.0;       // C[2b-2], C[2b-1] for b = 0
        double cr_m1 = 0.0;                // coal_rates[2b-1]
        double lnb_prev = 0.0;             // tv[LNB][bb - 1]
        double *lin = ch.lin + (size_t) b * 7 * T;
        const int NROW = AWB_TM_NVEC + 7;  // rows of tv, then of lin
        double cur[AWB_TM_NVEC + 7], prv[AWB_TM_NVEC + 7];
        for (int bb = 0; bb < T - 1; bb++) {
            const double cr0 = m.coal_time_steps[2 * bb] * nbranches[bb] /
                (2.0 * m.popsizes[bb]);                       // coal_rates[2b]
            const double cr1 = m.coal_time_steps[2 * bb + 1] * nbranches[bb] /
                (2.0 * m.popsizes[bb]);                       // coal_rates[2b+1]
            double treelen2 = tl + m.times[bb];
            double treelen2_b;
            if (bb > root_age_index) {
                treelen2 += m.times[bb] - root_age;
                treelen2_b = treelen2 + m.time_steps[bb];
            } else {
                treelen2_b = treelen2 + m.time_steps[root_age_index];
            }
            const double nb = nbranches[bb], nr = nrecombs[bb];
            const double term = Cm1 + log(m.time_steps[bb] * (nb + 1.0) / (nr + 1.0));
            const double lnb = (bb == 0) ? term : awb_logadd(lnb_prev, term);
            const double le2 = -Cm2 +
                (bb < T - 2 ? log(1 - exp(-cr0 - cr_m1)) : 0.0);
            const int below = (bb < root_age_index) ? 1 : 0;
            const double lnnegg1 = Cm1 + log(-m.time_steps[bb] * (
                (nb / (nr + 1.0 + below)) - (nb + 1.0) / (nr + 1.0)));
            const double g = (bb < T - 2 ? 1.0 - exp(-cr0) : 1.0);
            const double g2 = g * m.time_steps[bb] * (nb + 1.0) / (nr + 1.0);
            const double g3 = g * m.time_steps[bb] * (nb / (nr + 1.0 + below));
            const double Dv = (1.0 - exp(-m.rho * treelen2)) / treelen2_b;
            const double Ev = 1.0 / ncoals[bb];
            const double nrc = exp(-fmax(m.rho * treelen2, m.rho));
            cur[AWB_TM_LNB] = lnb;
            cur[AWB_TM_LNE2] = le2;
            cur[AWB_TM_LNNEGG1] = lnnegg1;
            cur[AWB_TM_G2] = g2;
            cur[AWB_TM_G3] = g3;
            cur[AWB_TM_LNG4] = le2;             // (trans.cpp:93: the same expression)
            cur[AWB_TM_D] = Dv;
            cur[AWB_TM_E] = Ev;
            cur[AWB_TM_NORECOMBS] = nrc;
            // the linear-domain vectors of the fast forward kernel, from the
            // values at hand (the stored ones need not be read back)
            {
                const double Bx = exp(lnb);
                const double pre = (bb > 0) ? exp(le2 + lnb_prev) : 0.0;
                cur[AWB_TM_NVEC + 0] = Dv;
                cur[AWB_TM_NVEC + 1] = Bx - exp(lnnegg1);
                cur[AWB_TM_NVEC + 2] = Bx;
                cur[AWB_TM_NVEC + 3] = Ev * exp(le2);
                cur[AWB_TM_NVEC + 4] = Ev * (pre + g3);
                cur[AWB_TM_NVEC + 5] = Ev * (pre + g2);
                cur[AWB_TM_NVEC + 6] = nrc;
            }
            // two time points a store
            if (bb & 1) {
#pragma unroll
                for (int k = 0; k < NROW; k++)
                    awb_store_pair((k < AWB_TM_NVEC ? tv + k * T : lin + (k - AWB_TM_NVEC) * T) +
                                   bb - 1, prv[k], cur[k]);
            } else {
#pragma unroll
                for (int k = 0; k < NROW; k++)
                    prv[k] = cur[k];
            }
            lnb_prev = lnb;
            // advance: C[2b] = C[2b-1] + cr0 ; C[2b+1] = C[2b] + cr1
            const double C2b = Cm1 + cr0;
            Cm2 = C2b;
            Cm1 = C2b + cr1;
            cr_m1 = cr1;
        }
        // the last time point holds zeros; with it, the row of an even T - 2
#pragma unroll
        for (int k = 0; k < NROW; k++) {
            double *dst = (k < AWB_TM_NVEC ? tv + k * T : lin + (k - AWB_TM_NVEC) * T);
            if ((T - 1) & 1)
                awb_store_pair(dst + T - 2, prv[k], 0.0);
            else
                dst[T - 1] = 0.0;
        }
    }

    // ---- time-major permutation, row starts, partial slots
    unsigned short rowstart[AWB_MAXT + 1];      // (kept here too: read below)
    unsigned short *pstart = ch.pstart + (size_t) b * (T + 1);
    {
        // states per time row (difference array over the branches), row starts,
        // then one pass over the branches fills the time-major permutation and
        // its inverse; within a row the states keep node order
        int rowpos[AWB_MAXT + 1];
        for (int t = 0; t <= T; t++)
            rowpos[t] = 0;
        if (S > 0) {
            for (int i = 0; i < V; i++) {
                const int cnt = ncnt[i];
                if (cnt <= 0) continue;
                const int lo = awb_imax(age[i], minage);
                rowpos[lo]++;
                rowpos[lo + cnt]--;
            }
        }
        int q = 0, run = 0;
        for (int t = 0; t < T; t++) {
            run += rowpos[t];
            rowstart[t] = (unsigned short) q;
            rowpos[t] = q;
            if (t < T - 1)
                q += run;
        }
        rowstart[T] = (unsigned short) q;
        {
            AwbPacker<unsigned short> pk;
            pk.init(ch.rowstart, (long long) b * (T + 1));
            for (int t = 0; t <= T; t++)
                pk.put(rowstart[t]);
            pk.flush();
        }

        // ---- the F-scribes' view of the column (awb_forward_fast.cuh).  The 64
        // scribe lanes are shared out over the T-1 time rows, nl lanes for a row
        // of w states with nl = ceil(w / CH) and CH the smallest (even) count
        // that makes all rows fit (a row stays inside one warp).  In shared
        // memory every row is PADDED to nl*CH slots, the lanes of a row read it
        // interleaved (lane i: slots i, i+nl, ...; consecutive lanes, consecutive
        // words, no bank conflicts), and EVERY lane reads exactly CH slots -- the
        // padding holds zeros -- so the summation loop has no clamps or masks.
        // iperm[state] is the state's padded slot.
        int zbase[AWB_MAXT];
        {
            int wrow[AWB_MAXT];
            for (int t = 0; t < T - 1; t++)
                wrow[t] = (S == 0) ? (t == 0 ? 1 : 0) : rowstart[t + 1] - rowstart[t];
            int CH;
            if (awb_scribe_plan(wrow, T - 1, CH) > ch.zcap)
                return 7;       // the host layout sized the buffer with the same code
            ch.sc_ch[b] = (unsigned short) CH;
            // (the four tables are written once, lane by lane, through packers;
            // a lane without a row: start 0, count 0, row 255, stride 1)
            AwbPacker<unsigned short> pks, pkc;
            AwbPacker<unsigned char> pkr, pkt;
            pks.init(ch.sc_start, (long long) b * AWB_NSCRIBE);
            pkc.init(ch.sc_cnt, (long long) b * AWB_NSCRIBE);
            pkr.init(ch.sc_row, (long long) b * AWB_NSCRIBE);
            pkt.init(ch.sc_stride, (long long) b * AWB_NSCRIBE);
            int l = 0, zb = 0;
            for (int t = 0; t < T - 1; t++) {
                // every row gets at least one lane: an empty row sums zeros
                const int w = wrow[t];
                const int nl = w == 0 ? 1 : (w + CH - 1) / CH;
                if ((l & 31) + nl > 32) {
                    for (; l & 31; l++) {
                        pks.put(0u); pkc.put(0u); pkr.put(255u); pkt.put(1u);
                    }
                }
                zbase[t] = zb - (S == 0 ? 0 : rowstart[t]);   // padded slot = zbase + time-major position
                for (int i = 0; i < nl; i++) {
                    pks.put((unsigned) (zb + i));
                    pkc.put((unsigned) (w > i ? (w - i + nl - 1) / nl : 0));   // real slots
                    pkr.put((unsigned) t);
                    pkt.put((unsigned) nl);
                    l++;
                }
                zb += nl * CH;
            }
            for (; l < AWB_NSCRIBE; l++) {
                pks.put(0u); pkc.put(0u); pkr.put(255u); pkt.put(1u);
            }
            pks.flush(); pkc.flush(); pkr.flush(); pkt.flush();
        }

        if (S > 0) {
            // (the states come in order, st = 0, 1, ..: iperm goes through a
            // packer; perm is the generic forward kernel's)
            AwbPacker<unsigned short> pk_iperm;
            pk_iperm.init(ch.iperm, row0);
            const bool want_perm = ch.need_band != 0;
            if (!want_perm) {
                // one running counter per row: the next padded slot
                for (int t = 0; t < T - 1; t++)
                    zbase[t] += rowpos[t];
                for (int i = 0; i < V; i++) {
                    const int cnt = ncnt[i];
                    if (cnt <= 0) continue;
                    const int lo = awb_imax(age[i], minage);
                    for (int x = 0; x < cnt; x++)
                        pk_iperm.put((unsigned) zbase[lo + x]++);
                }
            } else {
                for (int i = 0; i < V; i++) {
                    const int cnt = ncnt[i];
                    if (cnt <= 0) continue;
                    const int lo = awb_imax(age[i], minage);
                    const int nf = nfirst[i];
                    for (int x = 0; x < cnt; x++) {
                        const int tt = lo + x;
                        const int pos = rowpos[tt]++;
                        ch.perm[row0 + pos] = (unsigned short) (nf + x);
                        pk_iperm.put((unsigned) (zbase[tt] + pos));
                    }
                }
            }
            pk_iperm.flush();
        }
        if (S == 0)
            ch.perm[row0] = 0;
    }
    // (the generic forward kernel's slot tables)
    if (ch.need_band) {
        // partial slots: a new slot starts at each warp boundary and each new row
        int slot = -1;
        int t = 0;
        for (int qq = 0; qq < S; qq++) {
            bool newrow = false;
            while (t < T && rowstart[t + 1] <= qq) t++;   // row of qq
            if (qq == rowstart[t]) newrow = true;
            if (newrow || (qq & 31) == 0) {
                slot++;
                if (slot < ch.slotcap)
                    ch.slotrow[(size_t) b * ch.slotcap + slot] = (unsigned char) t;
            }
            ch.pslot[row0 + qq] = (unsigned short) slot;
        }
        if (S == 0)
            ch.pslot[row0] = 0;
        // first slot of each row (rows without states get the next row's slot)
        int cur = 0;
        for (int tt = 0; tt <= T; tt++) {
            if (tt < T && rowstart[tt] < S)
                cur = ch.pslot[row0 + rowstart[tt]];
            else
                cur = slot + 1;
            pstart[tt] = (unsigned short) cur;
        }
    }

    // ---- tables of the fast forward kernel (awb_forward_fast.cuh)
    {
        // node-major thread map; a branch never straddles a warp
        const long long tr0 = ch.trow_off[b];
        const int NSb = (int) (ch.trow_off[b + 1] - tr0);
        if (S == 0) {
            AwbPacker<unsigned short> pk;
            pk.init(ch.tmap, tr0);
            pk.put(0u);
            for (int t = 1; t < NSb; t++)
                pk.put(0xFFFFu);
            pk.flush();
            ch.iperm[row0] = 0;
            ch.st_age[row0] = 0;
        } else {
            // first-fit-decreasing packing when it fits the reserved slots (it
            // nearly always does, and is tighter); else node order, which is
            // what the host reserved (awb_count_states)
            AwbRowLists<VCAP> rl;
            if (awb_pack_place<VCAP>(ncnt, V, &rl) > NSb) {
                if (awb_pack_place_in_order<VCAP>(ncnt, V, &rl) > NSb)
                    return 6;
            }
            awb_pack_emit<VCAP>(rl, ncnt, nfirst, ch.tmap, tr0, NSb);
        }
    }

    if (S == 0) {
        ch.inv_emit[row0] = 1.0;
        if (ch.need_band) {
            ch.band_j1[row0] = 0;
            ch.band_len[row0] = 0;
            ch.band_boff[row0] = 0;
        }
        if (b == 0)
            ch.fw[ch.fw_off[0]] = 1.0;        // trans.cpp:822-825
        return 0;
    }

    // ---- (T x T) time matrix (sample_thread.cpp:212-216): on the GPU a kernel
    //      of its own fills it, one thread per entry (awb_tmatrix_fill)
#if !defined(__CUDA_ARCH__)
    awb_tmatrix_fill(ch, b, 0, 1);
#endif

    // ---- same-branch band (tmatrix2, sample_thread.cpp:218-225, :282-286),
    //      invariant-site emission (emit.cpp:786-819)
    double *band = ch.band + ch.band_off[b];
    int boff = 0;
    for (int k = 0; ch.need_band && k < S; k++) {
        const int node = ch.st_node[row0 + k];
        const int bt = ch.st_time[row0 + k];
        const int c = age[node];
        const int lo = awb_imax(c, minage);
        const int len = ncnt[node];
        ch.band_j1[row0 + k] = (unsigned short) nfirst[node];
        ch.band_len[row0 + k] = (unsigned char) len;
        ch.band_boff[row0 + k] = boff;
        for (int i = 0; i < len; i++) {
            const int a = lo + i;
            band[boff + i] = awb_get_time(tv, T, a, bt, c, minage, true) -
                awb_get_time(tv, T, a, bt, 0, minage, false);
        }
        boff += len;
    }
    if (ch.need_band && (long long) boff != ch.band_off[b + 1] - ch.band_off[b])
        return 2;

    // ---- prior column of the first block (sample_thread.cpp:425-429)
    if (b == 0) {
        double *col = ch.fw + ch.fw_off[0];
        for (int k = 0; k < S; k++)
            col[k] = awb_state_prior(ch.st_time[row0 + k], nbranches, ncoals,
                                     ch.minage, m);
    }
    return 0;
}

// working-array capacities the kernels are built for (the stack of a thread is
// sized by the largest array, and the driver reserves it for every thread that
// can be resident)
#define AWB_K1_VCAP_SMALL 128
#define AWB_K1_VCAP_MID 256

AWB_HD inline int awb_block_setup(const AwbChain &ch, int b)
{
    if (ch.nnodes <= AWB_K1_VCAP_SMALL)
        return awb_block_setup_t<AWB_K1_VCAP_SMALL>(ch, b);
    if (ch.nnodes <= AWB_K1_VCAP_MID)
        return awb_block_setup_t<AWB_K1_VCAP_MID>(ch, b);
    return awb_block_setup_t<AWB_MAXV>(ch, b);
}

// ---------------------------------------------------------------------------
// K2

struct AwbTreeView {
    const int *parent;
    const int *age;
    const short *c0, *c1;
    const short *nfirst, *ncnt;
    int root;
    int minage;
};

// NodeStateLookup::lookup (states.h:109-123): index of state (node,time) or -1
AWB_HD inline int awb_lookup(const AwbTreeView &t, int node, int time)
{
    const int cnt = t.ncnt[node];
    if (cnt <= 0)
        return -1;
    const int lo = awb_imax(t.age[node], t.minage);
    const int idx = time - lo;
    if (idx < 0 || idx >= cnt)
        return -1;
    return t.nfirst[node] + idx;
}

struct AwbSpr { int recomb_node, recomb_time, coal_node, coal_time; };

struct AwbLineages { const int *nbranches, *nrecombs, *ncoals; };

// The two ancestor walks of calc_recomb (trans.cpp:345-356) depend only on the
// SPR, not on the state: child-of-root ancestors of coal_node and recomb_node.
struct AwbRecombCtx { int ptr2, ptr3; };

AWB_HD inline AwbRecombCtx awb_recomb_ctx(const AwbTreeView &lt, const AwbSpr &spr,
                                          bool internal)
{
    AwbRecombCtx c = { -1, -1 };
    if (!internal)
        return c;
    int ptr = spr.coal_node;
    while (ptr != lt.root) {
        c.ptr2 = ptr;
        ptr = lt.parent[ptr];
    }
    ptr = spr.recomb_node;
    while (ptr != lt.root) {
        c.ptr3 = ptr;
        ptr = lt.parent[ptr];
    }
    return c;
}

// trans.cpp:317-413
AWB_HD inline double awb_calc_recomb(const AwbTreeView &lt, const AwbModel &m,
                                     const AwbLineages &L, const AwbSpr &spr,
                                     int s_node, int s_time,
                                     double last_treelen, bool internal,
                                     const AwbRecombCtx *rcx = 0)
{
    const int a = s_time;
    const int k = spr.recomb_time;
    double last_treelen_b;
    int root_age;

    if (internal) {
        const int subtree_root = lt.c0[lt.root];
        const int maintree_root = lt.c1[lt.root];
        root_age = lt.age[maintree_root];

        if (spr.coal_node == subtree_root) {
            if (a < spr.coal_time)
                return 0.0;
            if (spr.recomb_node == maintree_root && s_node != maintree_root)
                return 0.0;
        }
        const AwbRecombCtx cx = rcx ? *rcx : awb_recomb_ctx(lt, spr, internal);
        const int ptr2 = cx.ptr2, ptr3 = cx.ptr3;
        int ptr;
        if (ptr2 == subtree_root && ptr3 == maintree_root &&
            s_time == spr.recomb_time) {
            ptr = lt.parent[s_node];
            while (ptr != lt.root) {
                if (ptr == spr.recomb_node)
                    return 0.0;
                ptr = lt.parent[ptr];
            }
        }
        last_treelen += m.times[a] - m.times[lt.age[subtree_root]];
        if (a > root_age) {
            last_treelen += m.times[a] - m.times[root_age];
            last_treelen_b = last_treelen + m.time_steps[a];
        } else {
            last_treelen_b = last_treelen + m.time_steps[root_age];
        }
    } else {
        root_age = lt.age[lt.root];
        // get_treelen_branch / get_basal_branch (local_tree.cpp:179-219)
        const double blen = m.times[a];
        double treelen2 = last_treelen + blen;
        double root_time;
        if (s_node == lt.root) {
            treelen2 += blen - m.times[lt.age[lt.root]];
            root_time = m.times[a + 1] - m.times[a];
        } else {
            const int rooti = lt.age[lt.root];
            root_time = m.times[rooti + 1] - m.times[rooti];
        }
        last_treelen = treelen2;
        last_treelen_b = treelen2 + root_time;
    }

    const int nbranches_k = L.nbranches[k] + (k < a ? 1 : 0);
    const int nrecombs_k = L.nrecombs[k] + (k <= a ? 1 : 0) + (k == a ? 1 : 0) -
        (k >= awb_imax(root_age, a) ? 1 : 0);
    return nbranches_k * m.time_steps[k] / (nrecombs_k * last_treelen_b) *
        (1.0 - exp(-fmax(m.rho * last_treelen, m.rho)));
}

// trans.cpp:443-502
AWB_HD inline double awb_calc_recoal(const AwbTreeView &lt, const AwbModel &m,
                                     const AwbLineages &L, const AwbSpr &spr,
                                     int a, int recomb_parent_age, bool internal)
{
    const int k = spr.recomb_time;
    const int j = spr.coal_time;
    int nbranches_j = L.nbranches[j] - (j < recomb_parent_age ? 1 : 0) +
        (j < a ? 1 : 0);
    int ncoals_j = L.ncoals[j] - (j <= recomb_parent_age ? 1 : 0) -
        (j == recomb_parent_age ? 1 : 0) + (j <= a ? 1 : 0) + (j == a ? 1 : 0);
    bool over = false;
    if (internal) {
        const int subtree_root = lt.c0[lt.root];
        const int maintree_root = lt.c1[lt.root];
        if (spr.recomb_node == maintree_root &&
            spr.coal_time >= lt.age[subtree_root]) {
            over = true;
            nbranches_j = 1;
            ncoals_j++;
        }
    }
    double p = 1.0 / ncoals_j;
    if (j < m.ntimes - 2) {
        double Z = 0.0;
        if (j > k) {
            int b1 = L.nbranches[j - 1] - (j - 1 < recomb_parent_age ? 1 : 0) +
                (j - 1 < a ? 1 : 0);
            if (over)
                b1 = 1;
            Z = m.coal_time_steps[2 * j - 1] * b1 / (2.0 * m.popsizes[j - 1]);
        }
        p *= 1.0 - exp(-m.coal_time_steps[2 * j] * nbranches_j /
                       (2.0 * m.popsizes[j]) - Z);
    }
    return p;
}

// trans.cpp:505-534
AWB_HD inline double awb_calc_recomb_recoal(
    const AwbTreeView &lt, const AwbModel &m, const AwbLineages &L,
    const AwbSpr &spr, int s_node, int s_time, int recomb_parent_age,
    double last_treelen, bool internal, const AwbRecombCtx *rcx = 0)
{
    const int a = s_time;
    const int k = spr.recomb_time;
    const int j = spr.coal_time;
    double p = awb_calc_recomb(lt, m, L, spr, s_node, s_time, last_treelen,
                               internal, rcx);
    double sum = 0.0;
    for (int mm = 2 * k; mm < 2 * j - 1; mm++) {
        const int nbranches_m = L.nbranches[mm / 2] -
            (mm / 2 < recomb_parent_age ? 1 : 0) + (mm / 2 < a ? 1 : 0);
        sum += m.coal_time_steps[mm] * nbranches_m / (2.0 * m.popsizes[mm / 2]);
    }
    p *= exp(-sum);
    p *= awb_calc_recoal(lt, m, L, spr, s_time, recomb_parent_age, internal);
    return p;
}

// trans.cpp:156-275, one source state
AWB_HD inline int awb_determ_one(const AwbTreeView &lt, const AwbTreeView &t,
                                 const AwbSpr &spr, const int *mapping,
                                 int node1, int time1, bool internal)
{
#define AWB_MAP(x) (mapping[(x)])
    if ((node1 == spr.coal_node && time1 == spr.coal_time) ||
        (node1 == spr.recomb_node && time1 == spr.recomb_time))
        return -1;

    if (node1 != spr.recomb_node) {
        int node2;
        bool disrupt = false;
        if (lt.c0[node1] == -1) {
            node2 = node1;
        } else {
            const int child1 = lt.c0[node1], child2 = lt.c1[node1];
            if (spr.recomb_node == child1) {
                node2 = AWB_MAP(child2);
                disrupt = true;
            } else if (spr.recomb_node == child2) {
                node2 = AWB_MAP(child1);
                disrupt = true;
            } else {
                node2 = AWB_MAP(node1);
            }
        }
        if ((spr.coal_node == node1 && spr.coal_time < time1) ||
            (AWB_MAP(spr.coal_node) == node2 && spr.coal_time < time1) ||
            (disrupt && AWB_MAP(spr.coal_node) == node2 &&
             spr.coal_time <= time1))
            node2 = t.parent[node2];

        if (internal && t.age[node2] > time1)
            return -1;
        const int p = t.parent[node2];
        if (p != -1 && internal && time1 > t.age[p])
            return -1;
        return awb_lookup(t, node2, time1);
    }

    if (spr.recomb_time > time1)
        return awb_lookup(t, AWB_MAP(spr.recomb_node), time1);

    const int parent = lt.parent[spr.recomb_node];
    const int time2 = lt.age[parent];
    const int other = (lt.c1[parent] == spr.recomb_node ? lt.c0[parent] :
                       lt.c1[parent]);
    const int node2 = (other == spr.coal_node ? t.parent[AWB_MAP(other)] :
                       AWB_MAP(other));
    return awb_lookup(t, node2, time2);
}

AWB_HD inline void awb_tree_view(const AwbChain &ch, int b, AwbTreeView &t)
{
    const int V = ch.nnodes;
    t.parent = ch.ptrees + (size_t) b * V;
    t.age = ch.ages + (size_t) b * V;
    t.c0 = ch.child0 + (size_t) b * V;
    t.c1 = ch.child1 + (size_t) b * V;
    t.nfirst = ch.node_first + (size_t) b * V;
    t.ncnt = ch.node_cnt + (size_t) b * V;
    t.root = ch.root[b];
    t.minage = ch.tm_minage[b];
}

// Build the switch matrix of block b (b >= 1) as a CSR gather list:
// target k (state order of block b) <- sources (state order of block b-1).
// Entry order per target = the reference's accumulation order
// (sample_thread.cpp:357-373): deterministic sources ascending, then the
// recombination source, then the re-coalescence source.
AWB_HD inline int awb_switch_setup(const AwbChain &ch, int b)
{
    const AwbModel &m = ch.model;
    const int T = m.ntimes;
    const bool internal = ch.internal != 0;
    const int S1 = ch.nstates[b - 1], S2 = ch.nstates[b];
    const int n1 = awb_imax(S1, 1), n2 = awb_imax(S2, 1);
    const long long r1 = ch.row_off[b - 1], r2 = ch.row_off[b];
    const long long e0 = ch.ent_off[b];
    const int ecap = (int) (ch.ent_off[b + 1] - e0);

    AwbTreeView lt, t;
    awb_tree_view(ch, b - 1, lt);
    awb_tree_view(ch, b, t);
    AwbSpr spr = { ch.sprs[4 * b], ch.sprs[4 * b + 1], ch.sprs[4 * b + 2],
                   ch.sprs[4 * b + 3] };
    AwbLineages L;
    L.nbranches = ch.lineages + (size_t) (b - 1) * 3 * T;
    L.nrecombs = L.nbranches + T;
    L.ncoals = L.nrecombs + T;
    const double last_treelen = ch.treelen[b - 1];

    unsigned short *cnt = ch.sw_cnt + r2;
    unsigned short *start = ch.sw_start + r2;
    unsigned short *esrc = ch.sw_src + e0;
    double *eprob = ch.sw_prob + e0;
    for (int k = 0; k < n2; k++)
        cnt[k] = 0;

    int *dbg_determ = ch.sw_determ + ch.sw1_off[b];   // also the pass-1 scratch
    double *dbg_dprob = ch.keep_debug ? ch.sw_determprob + ch.sw1_off[b] : 0;
    double *dbg_rrow = ch.keep_debug ? ch.sw_recombrow + r2 : 0;
    double *dbg_crow = ch.keep_debug ? ch.sw_recoalrow + r2 : 0;
    for (int j = 0; j < n1; j++)
        dbg_determ[j] = -1;
    if (ch.keep_debug) {
        for (int j = 0; j < n1; j++) dbg_dprob[j] = 0.0;
        for (int k = 0; k < n2; k++) { dbg_rrow[k] = 0.0; dbg_crow[k] = 0.0; }
        ch.sw_recombsrc[b] = -1;
        ch.sw_recoalsrc[b] = -1;
    }

    // ---- internal-mode corner cases (trans.cpp:554-603)
    if (internal && S1 == 0) {
        int target = 0;
        if (S2 > 0) {
            const int maintree_root = t.c1[t.root];
            target = awb_lookup(t, maintree_root, spr.coal_time);
            if (target < 0)
                return 3;
        }
        for (int k = 0; k < n2; k++) {
            start[k] = (unsigned short) (k > target ? 1 : 0);
            cnt[k] = (unsigned short) (k == target ? 1 : 0);
        }
        esrc[0] = 0;
        eprob[0] = 1.0;
        dbg_determ[0] = target;
        if (ch.keep_debug) dbg_dprob[0] = 1.0;
        return 0;
    }
    if (internal && S2 == 0) {
        if (S1 > ecap)
            return 4;
        for (int i = 0; i < S1; i++) {
            const int node1 = ch.st_node[r1 + i], time1 = ch.st_time[r1 + i];
            int rpa;
            if (node1 == spr.recomb_node && time1 > spr.recomb_time)
                rpa = time1;
            else
                rpa = lt.age[lt.parent[spr.recomb_node]];
            const double p = awb_calc_recomb_recoal(lt, m, L, spr, node1, time1,
                                                    rpa, last_treelen, internal);
            esrc[i] = (unsigned short) i;
            eprob[i] = p;
            dbg_determ[i] = 0;
            if (ch.keep_debug) dbg_dprob[i] = p;
        }
        start[0] = 0;
        cnt[0] = (unsigned short) S1;
        return 0;
    }

    const int *mapping = ch.mappings + (size_t) b * ch.nnodes;
    const int broken = lt.parent[spr.recomb_node];
    const int recomb_parent_age0 = lt.age[broken];

    // ---- per-time tables (trans.cpp:616-621, 417-440)
    double sums = 0.0;
    double sums2[2 * AWB_MAXT + 1];
    double recoals[AWB_MAXT];
    {
        const int k = spr.recomb_time, j = spr.coal_time;
        double sum = 0.0;
        for (int mm = 2 * k; mm < 2 * j - 1; mm++) {
            const int nbm = L.nbranches[mm / 2] -
                (mm / 2 < recomb_parent_age0 ? 1 : 0);
            sum += m.coal_time_steps[mm] * nbm / (2.0 * m.popsizes[mm / 2]);
        }
        sums = sum;
        sum = 0.0;
        for (int mm = 0; mm < 2 * T + 1; mm++) sums2[mm] = 0.0;
        sums2[2 * k] = sum;
        for (int mm = 2 * k; mm < 2 * j - 1; mm++) {
            sum += m.coal_time_steps[mm] / (2.0 * m.popsizes[mm / 2]);
            sums2[mm + 1] = sum;
        }
        for (int a = 0; a < T; a++)
            recoals[a] = awb_calc_recoal(lt, m, L, spr, a, recomb_parent_age0,
                                         false /* sic, trans.cpp:620 */);
    }

    // ---- pass 1: classify sources, count entries per target
    int recombsrc = -1, recoalsrc = -1;
    for (int i = 0; i < S1; i++) {
        const int node1 = ch.st_node[r1 + i], time1 = ch.st_time[r1 + i];
        if (node1 == spr.recomb_node && time1 == spr.recomb_time)
            recombsrc = i;
        else if (node1 == spr.coal_node && time1 == spr.coal_time)
            recoalsrc = i;
        const int d = awb_determ_one(lt, t, spr, mapping, node1, time1, internal);
        dbg_determ[i] = d;
        if (d >= 0)
            cnt[d]++;
    }

    // recombination row (trans.cpp:665-693): "stay" and "escape"
    int rk[2] = { -1, -1 };
    double rv[2] = { 0.0, 0.0 };
    if (recombsrc != -1) {
        const int parent = lt.parent[spr.recomb_node];
        const int time2 = lt.age[parent];
        const int other = (lt.c0[parent] == spr.recomb_node ? lt.c1[parent] :
                           lt.c0[parent]);
        const int node2 = (other == spr.coal_node ? t.parent[AWB_MAP(other)] :
                           AWB_MAP(other));
        rk[0] = awb_lookup(t, AWB_MAP(spr.recomb_node), spr.recomb_time);
        rk[1] = awb_lookup(t, node2, time2);
        const int sn = ch.st_node[r1 + recombsrc], stime = ch.st_time[r1 + recombsrc];
        if (rk[0] != -1)
            rv[0] = awb_calc_recomb_recoal(lt, m, L, spr, sn, stime,
                                           recomb_parent_age0, last_treelen,
                                           internal);
        if (rk[1] != -1)
            rv[1] = awb_calc_recomb_recoal(lt, m, L, spr, sn, stime, stime,
                                           last_treelen, internal);
        if (rk[0] != -1 && rk[0] == rk[1]) {
            // the second assignment overwrites the first (trans.cpp:678,688)
            rv[0] = rv[1];
            rk[1] = -1;
        }
        for (int q = 0; q < 2; q++)
            if (rk[q] != -1 && rv[q] > 0.0)
                cnt[rk[q]]++;
    }

    // re-coalescence row (trans.cpp:697-738)
    int node3 = -1, cparent = -1, ctime1 = -1, cnode1 = -1;
    if (recoalsrc != -1) {
        cnode1 = ch.st_node[r1 + recoalsrc];
        ctime1 = ch.st_time[r1 + recoalsrc];
        if (broken == cnode1)
            node3 = AWB_MAP(lt.c1[broken] == spr.recomb_node ? lt.c0[broken] :
                            lt.c1[broken]);
        else
            node3 = AWB_MAP(cnode1);
        cparent = t.parent[AWB_MAP(spr.recomb_node)];
    }

    // re-coalescence row entries, computed once (at most T+3 of them)
    int nck = 0;
    int ckk[AWB_MAXT + 4];
    double ckv[AWB_MAXT + 4];
    if (recoalsrc != -1) {
        for (int k = 0; k < S2; k++) {
            const int node2 = ch.st_node[r2 + k], time2 = ch.st_time[r2 + k];
            if (!((node2 == AWB_MAP(spr.recomb_node) && time2 >= spr.recomb_time) ||
                  (node2 == node3 && time2 == ctime1) ||
                  (node2 == cparent && time2 == ctime1)))
                continue;
            AwbSpr spr2 = spr;
            spr2.coal_time = time2;
            const double p = awb_calc_recomb_recoal(
                lt, m, L, spr2, cnode1, ctime1, recomb_parent_age0,
                last_treelen, internal);
            if (ch.keep_debug) dbg_crow[k] = p;
            if (p > 0.0 && nck < AWB_MAXT + 4) {
                ckk[nck] = k;
                ckv[nck] = p;
                nck++;
                cnt[k]++;
            }
        }
    }
    int total = 0;
    for (int k = 0; k < n2; k++) {
        start[k] = (unsigned short) total;
        total += cnt[k];
        cnt[k] = 0;
    }
    if (total > ecap)
        return 4;

    // ---- pass 2: deterministic entries in ascending source order
    for (int i = 0; i < S1; i++) {
        const int node1 = ch.st_node[r1 + i], time1 = ch.st_time[r1 + i];
        const int d = dbg_determ[i];
        if (d < 0)
            continue;
        double p;
        if (node1 == spr.recomb_node && time1 > spr.recomb_time) {
            p = awb_calc_recomb_recoal(lt, m, L, spr, node1, time1, time1,
                                       last_treelen, internal);
        } else {
            const int idx = awb_imax(awb_imin(2 * spr.coal_time - 1, 2 * time1),
                                     2 * spr.recomb_time);
            p = awb_calc_recomb(lt, m, L, spr, node1, time1, last_treelen,
                                internal) *
                exp(-sums - sums2[idx]) * recoals[time1];
        }
        if (ch.keep_debug) dbg_dprob[i] = p;
        const int pos = start[d] + cnt[d]++;
        esrc[pos] = (unsigned short) i;
        eprob[pos] = p;
    }
    // recombination source
    for (int q = 0; q < 2; q++) {
        if (rk[q] != -1) {
            if (ch.keep_debug) dbg_rrow[rk[q]] = rv[q];
            if (rv[q] > 0.0) {
                const int pos = start[rk[q]] + cnt[rk[q]]++;
                esrc[pos] = (unsigned short) recombsrc;
                eprob[pos] = rv[q];
            }
        }
    }
    // re-coalescence source
    for (int q = 0; q < nck; q++) {
        const int pos = start[ckk[q]] + cnt[ckk[q]]++;
        esrc[pos] = (unsigned short) recoalsrc;
        eprob[pos] = ckv[q];
    }
    if (ch.keep_debug) {
        ch.sw_recombsrc[b] = recombsrc;
        ch.sw_recoalsrc[b] = recoalsrc;
    }
#undef AWB_MAP
    return 0;
}

#if defined(__CUDACC__)
// ---------------------------------------------------------------------------
// K2, warp-cooperative: one warp builds the switch CSR of one breakpoint.
// Same result as awb_switch_setup above (entry order per target: deterministic
// sources ascending, recombination source, re-coalescence source); sources are
// classified 32 at a time, targets counted with __match_any groups (a stable
// counting sort, so the order does not depend on scheduling).
//
// scratch per warp (shared memory): cnt[maxS] start[maxS] (u16) |
// sums2[2T+1] recoals[T] ckv[T+4] (f64) | ckk[T+4] (i32)
__host__ __device__ inline size_t awb_sw_warp_scratch_bytes(int maxS, int T)
{
    size_t n = (((size_t) (maxS > 0 ? maxS : 1) * 4) + 15) & ~(size_t) 15;
    n += (size_t) (2 * T + 1 + T + T + 4) * sizeof(double);
    n += (((size_t) (T + 4) * sizeof(int)) + 15) & ~(size_t) 15;
    return (n + 15) & ~(size_t) 15;
}

__device__ inline int awb_switch_setup_warp(const AwbChain &ch, int b, int lane,
                                            unsigned char *scr)
{
    const unsigned FULL = 0xffffffffu;
    const AwbModel &m = ch.model;
    const int T = m.ntimes;
    const bool internal = ch.internal != 0;
    const int S1 = ch.nstates[b - 1], S2 = ch.nstates[b];
    const int n1 = awb_imax(S1, 1), n2 = awb_imax(S2, 1);
    const long long r1 = ch.row_off[b - 1], r2 = ch.row_off[b];
    const long long e0 = ch.ent_off[b];
    const int ecap = (int) (ch.ent_off[b + 1] - e0);
    const int capS = ch.maxS > 0 ? ch.maxS : 1;

    unsigned short *cntS = (unsigned short *) scr;
    unsigned short *startS = cntS + capS;
    double *sums2 = (double *) (scr + ((((size_t) capS * 4) + 15) & ~(size_t) 15));
    double *recoals = sums2 + (2 * T + 1);
    double *ckv = recoals + T;
    int *ckk = (int *) (ckv + (T + 4));

    AwbTreeView lt, t;
    awb_tree_view(ch, b - 1, lt);
    awb_tree_view(ch, b, t);
    AwbSpr spr = { ch.sprs[4 * b], ch.sprs[4 * b + 1], ch.sprs[4 * b + 2],
                   ch.sprs[4 * b + 3] };
    AwbLineages L;
    L.nbranches = ch.lineages + (size_t) (b - 1) * 3 * T;
    L.nrecombs = L.nbranches + T;
    L.ncoals = L.nrecombs + T;
    const double last_treelen = ch.treelen[b - 1];

    unsigned short *cnt = ch.sw_cnt + r2;
    unsigned short *start = ch.sw_start + r2;
    unsigned short *esrc = ch.sw_src + e0;
    double *eprob = ch.sw_prob + e0;
    int *determ = ch.sw_determ + ch.sw1_off[b];
    double *dbg_dprob = ch.keep_debug ? ch.sw_determprob + ch.sw1_off[b] : 0;
    double *dbg_rrow = ch.keep_debug ? ch.sw_recombrow + r2 : 0;
    double *dbg_crow = ch.keep_debug ? ch.sw_recoalrow + r2 : 0;

    for (int k = lane; k < n2; k += 32)
        cntS[k] = 0;
    if (ch.keep_debug) {
        for (int j = lane; j < n1; j += 32) dbg_dprob[j] = 0.0;
        for (int k = lane; k < n2; k += 32) { dbg_rrow[k] = 0.0; dbg_crow[k] = 0.0; }
        if (lane == 0) {
            ch.sw_recombsrc[b] = -1;
            ch.sw_recoalsrc[b] = -1;
        }
    }

    // ---- internal-mode corner cases (trans.cpp:554-603)
    if (internal && S1 == 0) {
        int target = 0;
        if (S2 > 0) {
            const int maintree_root = t.c1[t.root];
            target = awb_lookup(t, maintree_root, spr.coal_time);
            if (target < 0)
                return 3;
        }
        for (int k = lane; k < n2; k += 32) {
            start[k] = (unsigned short) (k > target ? 1 : 0);
            cnt[k] = (unsigned short) (k == target ? 1 : 0);
        }
        if (lane == 0) {
            esrc[0] = 0;
            eprob[0] = 1.0;
            determ[0] = target;
            if (ch.keep_debug) dbg_dprob[0] = 1.0;
        }
        return 0;
    }
    const AwbRecombCtx rcx = awb_recomb_ctx(lt, spr, internal);
    if (internal && S2 == 0) {
        if (S1 > ecap)
            return 4;
        for (int i = lane; i < S1; i += 32) {
            const int node1 = ch.st_node[r1 + i], time1 = ch.st_time[r1 + i];
            int rpa;
            if (node1 == spr.recomb_node && time1 > spr.recomb_time)
                rpa = time1;
            else
                rpa = lt.age[lt.parent[spr.recomb_node]];
            const double p = awb_calc_recomb_recoal(lt, m, L, spr, node1, time1,
                                                    rpa, last_treelen, internal, &rcx);
            esrc[i] = (unsigned short) i;
            eprob[i] = p;
            determ[i] = 0;
            if (ch.keep_debug) dbg_dprob[i] = p;
        }
        if (lane == 0) {
            start[0] = 0;
            cnt[0] = (unsigned short) S1;
        }
        return 0;
    }

#define AWB_MAP(x) (mapping[(x)])
    const int *mapping = ch.mappings + (size_t) b * ch.nnodes;
    const int broken = lt.parent[spr.recomb_node];
    const int recomb_parent_age0 = lt.age[broken];

    // ---- per-time tables (trans.cpp:616-621, 417-440)
    double sums = 0.0;
    {
        const int k = spr.recomb_time, j = spr.coal_time;
        for (int mm = lane; mm < 2 * T + 1; mm += 32)
            sums2[mm] = 0.0;
        __syncwarp();
        double sum = 0.0;
        for (int mm = 2 * k; mm < 2 * j - 1; mm++) {
            const int nbm = L.nbranches[mm / 2] -
                (mm / 2 < recomb_parent_age0 ? 1 : 0);
            sum += m.coal_time_steps[mm] * nbm / (2.0 * m.popsizes[mm / 2]);
        }
        sums = sum;
        if (lane == 0) {
            sum = 0.0;
            sums2[2 * k] = sum;
            for (int mm = 2 * k; mm < 2 * j - 1; mm++) {
                sum += m.coal_time_steps[mm] / (2.0 * m.popsizes[mm / 2]);
                sums2[mm + 1] = sum;
            }
        }
        for (int a = lane; a < T; a += 32)
            recoals[a] = awb_calc_recoal(lt, m, L, spr, a, recomb_parent_age0,
                                         false /* sic, trans.cpp:620 */);
        __syncwarp();
    }

    // ---- pass 1: classify sources, count deterministic entries per target
    int recombsrc = -1, recoalsrc = -1;
    for (int base = 0; base < S1; base += 32) {
        const int i = base + lane;
        int d = -1;
        if (i < S1) {
            const int node1 = ch.st_node[r1 + i], time1 = ch.st_time[r1 + i];
            if (node1 == spr.recomb_node && time1 == spr.recomb_time)
                recombsrc = i;
            else if (node1 == spr.coal_node && time1 == spr.coal_time)
                recoalsrc = i;
            d = awb_determ_one(lt, t, spr, mapping, node1, time1, internal);
            determ[i] = d;
        }
        const unsigned act = __ballot_sync(FULL, d >= 0);
        if (d >= 0) {
            const unsigned g = __match_any_sync(act, d);
            if (lane == __ffs(g) - 1)
                cntS[d] = (unsigned short) (cntS[d] + __popc(g));
        }
        __syncwarp();
    }
    recombsrc = __reduce_max_sync(FULL, recombsrc);
    recoalsrc = __reduce_max_sync(FULL, recoalsrc);

    // recombination row (trans.cpp:665-693): "stay" and "escape" (uniform)
    int rk[2] = { -1, -1 };
    double rv[2] = { 0.0, 0.0 };
    if (recombsrc != -1) {
        const int parent = lt.parent[spr.recomb_node];
        const int time2 = lt.age[parent];
        const int other = (lt.c0[parent] == spr.recomb_node ? lt.c1[parent] :
                           lt.c0[parent]);
        const int node2 = (other == spr.coal_node ? t.parent[AWB_MAP(other)] :
                           AWB_MAP(other));
        rk[0] = awb_lookup(t, AWB_MAP(spr.recomb_node), spr.recomb_time);
        rk[1] = awb_lookup(t, node2, time2);
        const int sn = ch.st_node[r1 + recombsrc], stime = ch.st_time[r1 + recombsrc];
        if (rk[0] != -1)
            rv[0] = awb_calc_recomb_recoal(lt, m, L, spr, sn, stime,
                                           recomb_parent_age0, last_treelen,
                                           internal, &rcx);
        if (rk[1] != -1)
            rv[1] = awb_calc_recomb_recoal(lt, m, L, spr, sn, stime, stime,
                                           last_treelen, internal, &rcx);
        if (rk[0] != -1 && rk[0] == rk[1]) {
            rv[0] = rv[1];              // trans.cpp:678,688: second assignment wins
            rk[1] = -1;
        }
        if (lane == 0)
            for (int q = 0; q < 2; q++)
                if (rk[q] != -1 && rv[q] > 0.0)
                    cntS[rk[q]]++;
    }
    __syncwarp();

    // re-coalescence row (trans.cpp:697-738): candidate targets tested 32 at a time
    int nck = 0;
    if (recoalsrc != -1) {
        const int cnode1 = ch.st_node[r1 + recoalsrc];
        const int ctime1 = ch.st_time[r1 + recoalsrc];
        int node3;
        if (broken == cnode1)
            node3 = AWB_MAP(lt.c1[broken] == spr.recomb_node ? lt.c0[broken] :
                            lt.c1[broken]);
        else
            node3 = AWB_MAP(cnode1);
        const int cparent = t.parent[AWB_MAP(spr.recomb_node)];
        const int rmap = AWB_MAP(spr.recomb_node);
        for (int base = 0; base < S2; base += 32) {
            const int k = base + lane;
            bool hit = false;
            double p = 0.0;
            if (k < S2) {
                const int node2 = ch.st_node[r2 + k], time2 = ch.st_time[r2 + k];
                hit = (node2 == rmap && time2 >= spr.recomb_time) ||
                    (node2 == node3 && time2 == ctime1) ||
                    (node2 == cparent && time2 == ctime1);
                if (hit) {
                    AwbSpr spr2 = spr;
                    spr2.coal_time = time2;
                    p = awb_calc_recomb_recoal(lt, m, L, spr2, cnode1, ctime1,
                                               recomb_parent_age0, last_treelen,
                                               internal, &rcx);
                    if (ch.keep_debug) dbg_crow[k] = p;
                }
            }
            const bool keep = hit && p > 0.0;
            const unsigned mk = __ballot_sync(FULL, keep);
            if (keep) {
                const int pos = nck + __popc(mk & ((1u << lane) - 1u));
                if (pos < T + 4) {
                    ckk[pos] = k;
                    ckv[pos] = p;
                    cntS[k]++;          // one candidate per target: no conflict
                }
            }
            nck += __popc(mk);
        }
        nck = awb_imin(nck, T + 4);
    }
    __syncwarp();

    // ---- exclusive scan of the counts -> start[]
    int total = 0;
    for (int base = 0; base < n2; base += 32) {
        const int k = base + lane;
        const int c = (k < n2) ? cntS[k] : 0;
        int x = c;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const int y = __shfl_up_sync(FULL, x, dd);
            if (lane >= dd)
                x += y;
        }
        if (k < n2) {
            const unsigned short st = (unsigned short) (total + x - c);
            startS[k] = st;
            start[k] = st;
            cntS[k] = 0;
        }
        total += __shfl_sync(FULL, x, 31);
    }
    if (total > ecap)
        return 4;
    __syncwarp();

    // ---- pass 2: deterministic entries, ascending source order per target
    for (int base = 0; base < S1; base += 32) {
        const int i = base + lane;
        int d = -1;
        double p = 0.0;
        if (i < S1) {
            d = determ[i];
            if (d >= 0) {
                const int node1 = ch.st_node[r1 + i], time1 = ch.st_time[r1 + i];
                if (node1 == spr.recomb_node && time1 > spr.recomb_time) {
                    p = awb_calc_recomb_recoal(lt, m, L, spr, node1, time1, time1,
                                               last_treelen, internal, &rcx);
                } else {
                    const int idx = awb_imax(awb_imin(2 * spr.coal_time - 1, 2 * time1),
                                             2 * spr.recomb_time);
                    p = awb_calc_recomb(lt, m, L, spr, node1, time1, last_treelen,
                                        internal, &rcx) *
                        exp(-sums - sums2[idx]) * recoals[time1];
                }
                if (ch.keep_debug) dbg_dprob[i] = p;
            }
        }
        const unsigned act = __ballot_sync(FULL, d >= 0);
        if (d >= 0) {
            const unsigned g = __match_any_sync(act, d);
            const int pos = startS[d] + cntS[d] + __popc(g & ((1u << lane) - 1u));
            esrc[pos] = (unsigned short) i;
            eprob[pos] = p;
            __syncwarp(act);
            if (lane == __ffs(g) - 1)
                cntS[d] = (unsigned short) (cntS[d] + __popc(g));
        }
        __syncwarp();
    }

    // recombination source, then re-coalescence source
    if (lane == 0) {
        for (int q = 0; q < 2; q++) {
            if (rk[q] != -1) {
                if (ch.keep_debug) dbg_rrow[rk[q]] = rv[q];
                if (rv[q] > 0.0) {
                    const int pos = startS[rk[q]] + cntS[rk[q]]++;
                    esrc[pos] = (unsigned short) recombsrc;
                    eprob[pos] = rv[q];
                }
            }
        }
        for (int q = 0; q < nck; q++) {
            const int pos = startS[ckk[q]] + cntS[ckk[q]]++;
            esrc[pos] = (unsigned short) recoalsrc;
            eprob[pos] = ckv[q];
        }
        if (ch.keep_debug) {
            ch.sw_recombsrc[b] = recombsrc;
            ch.sw_recoalsrc[b] = recoalsrc;
        }
    }
    __syncwarp();
    for (int k = lane; k < n2; k += 32)
        cnt[k] = cntS[k];
#undef AWB_MAP
    return 0;
}
#endif // __CUDACC__


#endif // AWB_SETUP_CUH
