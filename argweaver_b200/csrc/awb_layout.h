// awb_layout.h -- host-side (pure C++) layout of one problem's arrays.
//
// Integer-only host logic: counts the HMM states of every local tree
// (reference states.cpp:53-74,109-166), derives the offsets of the per-state
// rows, the forward-table slabs, the same-branch bands and the switch CSR
// lists, validates the inputs, and assigns every device array a position in
// one arena.  Used by awb_api.cu (arena in HBM) and by the CPU emulation
// harness tests/host_emul.cpp (arena in host memory).
#ifndef AWB_LAYOUT_H
#define AWB_LAYOUT_H

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "argweaver_b200.h"
#include "awb_common.cuh"

struct AwbCopy {            // one host -> arena copy of an input array
    size_t dst_off;
    const void *src;
    size_t bytes;
    int own;                // see awb_layout_build ("input copies")
};

struct AwbLayout {
    int B, V, T, n, nrows;
    int maxS, maxband;
    int maxNS;                           // forward kernel: padded threads per block
    int zcap;                            // forward kernel: padded time-major column (slots)
    int maxcnt;                          // longest branch (states)
    int keep_debug;
    double states_sites;                 // sum blocklen * nstates
    std::vector<int> nstates, block_start, rowidx;
    int gen_mappings;                    // no mappings given: K1 makes them
    // the linear-domain transition vectors (lin: exp(lnB) ...) may overflow for
    // this problem (cumulative coalescent rate beyond the double exponent
    // range, e.g. tiny population sizes): the batch then runs the generic
    // forward kernel and the traceback uses the closed-form transitions
    int lin_unsafe;
    bool packed;                    // the alignment comes as variant columns (var_cols)
    size_t o_varpos;
    // Jukes-Cantor branch probabilities by time-index pair (emission kernel):
    // ptab[(y*T + x)*2 + {0: mutation, 1: none}] for a branch from time y up to
    // time x, and p0tab[x*2 + ..] for a branch from time 0.0 up to time x
    std::vector<double> ptab;
    size_t o_ptab;
    std::vector<long long> row_off, fw_off, band_off, ent_off, sw1_off, trow_off;
    AwbModel model;

    // arena
    size_t total_bytes;
    size_t total_bytes_noslots;     // checkpointed table: without the segment tables
    size_t slots_at;                // checkpointed table: where they start
    int nslots;                     // checkpointed table: resident segment tables
    size_t bytes_before_band;       // arena size when the band is left out
    // checkpointed table (seg_start.size() == nseg + 1, block indices)
    int ckpt, nseg;
    std::vector<int> seg_start;
    long long seg_doubles;          // largest segment table (+ one extra row)
    int seg_sites;                  // most sites in a segment (+ 1)
    size_t o_seg_start, o_ckptcol;
    size_t o_mappings, o_slotrow, o_trow_off, o_tmap, o_iperm, o_st_age, o_lin,
        o_sc_start, o_sc_cnt, o_sc_row, o_sc_stride, o_sc_ch;
    size_t o_ptrees, o_ages, o_sprs, o_blocklens, o_subtree_roots, o_rowidx,
        o_seqs, o_block_start, o_nstates, o_row_off, o_fw_off, o_band_off,
        o_ent_off, o_sw1_off, o_st_node, o_st_time, o_perm, o_pslot, o_band_j1,
        o_band_len, o_band_boff, o_inv_emit, o_band, o_tmatrix, o_tmvec,
        o_rowstart, o_pstart, o_node_first, o_node_cnt, o_child0, o_child1,
        o_order, o_lstart, o_root, o_lineages, o_treelen, o_tm_minage, o_sw_start,
        o_sw_cnt, o_sw_src, o_sw_prob, o_sw_determ, o_sw_determprob,
        o_sw_recombrow, o_sw_recoalrow, o_sw_recombsrc, o_sw_recoalsrc, o_kind,
        o_fw, o_path, o_rand, o_logz, o_status, o_sink, o_fsum;
    bool has_subtree_roots;
    std::vector<AwbCopy> copies;         // inputs taken straight from the caller
};

// model.h:322-333, model.cpp:9-23
inline void awb_model_fill(AwbModel &m, const awb_problem &p)
{
    memset(&m, 0, sizeof(m));
    const int T = p.ntimes;
    m.ntimes = T;
    m.removed_root_time = T + 1;
    m.rho = p.rho;
    m.mu = p.mu;
    for (int i = 0; i < T; i++) {
        m.times[i] = p.times[i];
        m.popsizes[i] = p.popsizes[i];
    }
    m.mintime = p.times[1] * .1;
    for (int i = 0; i < T - 1; i++)
        m.time_steps[i] = p.times[i + 1] - p.times[i];
    m.time_steps[T - 1] = INFINITY;
    std::vector<double> t2(2 * T + 1, p.times[T - 1]);
    for (int i = 0; i < T - 1; i++) {
        t2[2 * i] = p.times[i];
        t2[2 * i + 1] = sqrt((p.times[i + 1] + 1.0) * (p.times[i] + 1.0));
    }
    for (int i = 0; i < 2 * T - 2; i++)
        m.coal_time_steps[i] = t2[i + 1] - t2[i];
    m.coal_time_steps[2 * T - 2] = INFINITY;
    m.coal_time_steps[2 * T - 1] = INFINITY;
}

inline std::vector<short> &awb_pack_scratch()
{
    static thread_local std::vector<short> v;
    return v;
}


// A parent chain that never reaches the root can only run through nodes of
// EQUAL age (a parent is never younger than its child, checked by the callers):
// from every node whose parent has its own age, walk up while the age stays the
// same; more than V steps means a cycle.
inline bool awb_parent_cycle(const int *parent, const int *age, int V)
{
    for (int i = 0; i < V; i++) {
        int x = i, steps = 0;
        while (parent[x] >= 0 && parent[x] < V && age[parent[x]] == age[x]) {
            x = parent[x];
            if (++steps > V)
                return true;
        }
    }
    return false;
}

// Number of states of one tree and the sum of squared branch state counts.
// Returns false on a malformed tree.
inline bool awb_count_states_checked(const awb_problem &p, int b, std::vector<int> &c0,
                             std::vector<int> &c1, std::vector<int> &stack,
                             std::vector<char> &ignore, int &S, int &band,
                             int &tpos, int &maxcnt, std::string &err,
                             int *wrow = NULL)
{
    const int V = p.nnodes, T = p.ntimes;
    const int *parent = p.ptrees + (size_t) b * V;
    const int *age = p.ages + (size_t) b * V;
    int root = -1, nroots = 0;
    const bool internal = p.internal != 0;
    int *c0p = c0.data(), *c1p = c1.data();
    memset(c0p, 0xFF, (size_t) V * sizeof(int));       // -1
    memset(c1p, 0xFF, (size_t) V * sizeof(int));
    memset(ignore.data(), 0, (size_t) V);
    // one pass: children in node-index order (no data-dependent branches: which
    // child slot is free is a coin flip to the branch predictor) and the checks
    unsigned three = 0, young = 0, badage = 0;
    for (int i = 0; i < V; i++) {
        const int pa = parent[i], a = age[i];
        if (pa == -1) {
            root = i;
            nroots++;
            const bool vroot = internal && a == T + 1;
            if (!vroot && (a < 0 || a > T - 2)) badage = 1;
            continue;
        }
        if (pa < 0 || pa >= V || pa == i) {
            err = "tree " + std::to_string(b) + ": bad parent index";
            return false;
        }
        badage |= (unsigned) (a < 0) | (unsigned) (a > T - 2);
        young |= (unsigned) (age[pa] < a);
        const int f = c0p[pa];
        const bool has0 = f != -1;
        three |= (unsigned) (has0 & (c1p[pa] != -1));
        c0p[pa] = has0 ? f : i;
        c1p[pa] = has0 ? i : -1;
    }
    if (three) {
        err = "tree " + std::to_string(b) + ": node with three children";
        return false;
    }
    if (nroots != 1) {
        err = "tree " + std::to_string(b) + ": expected exactly one root";
        return false;
    }
    if (badage) {
        err = "tree " + std::to_string(b) + ": node age out of range";
        return false;
    }
    if (young) {
        err = "tree " + std::to_string(b) + ": parent younger than child";
        return false;
    }
    // a binary tree with the leaves listed first: nodes [0, nleaves) have no
    // children, every other node has exactly two (K1's level order, the subtree
    // walk below and the emission passes rely on it)
    for (int i = 0; i < V; i++) {
        const bool leaf = i < p.nleaves;
        if (leaf ? (c0p[i] != -1) : (c0p[i] == -1 || c1p[i] == -1)) {
            err = "tree " + std::to_string(b) +
                ": not a binary tree with the leaves listed first";
            return false;
        }
    }
    if (awb_parent_cycle(parent, age, V)) {
        err = "tree " + std::to_string(b) + ": parent cycle";
        return false;
    }
    S = 0;
    band = 0;
    tpos = 1;
    if (wrow) {
        for (int t = 0; t < T; t++) wrow[t] = 0;
        wrow[0] = 1;                        // (a block without states: one dummy slot)
    }
    int minage = p.minage;
    if (internal) {
        if (V < 3 || c0[root] == -1) {
            err = "internal mode needs a root with two children";
            return false;
        }
        if (age[root] < T)
            return true;                    // fully specified tree: no states
        int sub = c0[root];
        if (p.subtree_roots && p.subtree_roots[b] >= 0) {
            sub = p.subtree_roots[b];
            if (sub != c0[root] && sub != c1[root]) {
                err = "tree " + std::to_string(b) +
                    ": subtree_root is not a child of the root";
                return false;
            }
        }
        if (age[sub] > minage) minage = age[sub];
        ignore[root] = 1;
        int top = 0;
        stack[top++] = sub;
        while (top > 0) {
            const int node = stack[--top];
            ignore[node] = 1;
            if (c0[node] != -1) {
                stack[top++] = c0[node];
                stack[top++] = c1[node];
            }
        }
    }
    std::vector<short> &bcnt = awb_pack_scratch();
    bcnt.assign(V, 0);
    if (wrow) wrow[0] = 0;

    // tpos: thread slots of the forward kernel when the branches are laid out in
    // node order, none straddling a warp ("next fit") -- the capacity reserved
    // for this block's thread map; K1 packs tighter when it can (see below)
    tpos = 0;
    for (int i = 0; i < V; i++) {
        if (ignore[i]) continue;
        const int pa = parent[i];
        const int lo = age[i] > minage ? age[i] : minage;
        const int hi = (pa == -1 || (internal && pa == root)) ? T - 2 : age[pa];
        const int cnt = hi - lo + 1;
        if (cnt > 0) {
            S += cnt;
            band += cnt * cnt;
            bcnt[i] = (short) cnt;
            if (wrow) {             // states per time row, as differences
                wrow[lo]++;
                wrow[hi + 1]--;
            }
            if (cnt > maxcnt) maxcnt = cnt;
            if (cnt > 32) {
                // (two whole warps' worth of slots, starting at an even one: the
                // forward kernel keeps such a branch in two register sets of
                // one warp)
                tpos = ((tpos + 63) & ~63) + 64;
            } else {
                if ((tpos & 31) + cnt > 32)
                    tpos = (tpos + 31) & ~31;
                tpos += cnt;
            }
        }
    }
    if (S == 0)
        tpos = 1;
    if (wrow) {
        for (int t = 1; t < T; t++) wrow[t] += wrow[t - 1];
        if (S == 0) wrow[0] = 1;
    }
    return true;
}

// The same, with the common case (a new leaf is threaded: every branch carries
// states) done in ONE branch-light pass over the nodes -- the layout of a
// genome-scale batch visits ~5*10^8 nodes on the host.  Anything unusual (an
// invalid tree included) is left to the checked version above, which also
// words the error.
inline bool awb_count_states(const awb_problem &p, int b, std::vector<int> &c0,
                             std::vector<int> &c1, std::vector<int> &stack,
                             std::vector<char> &ignore, int &S, int &band,
                             int &tpos, int &maxcnt, std::string &err,
                             int *wrow)
{
    const int V = p.nnodes, T = p.ntimes;
    if (!p.internal) {
        const int *parent = p.ptrees + (size_t) b * V;
        const int *age = p.ages + (size_t) b * V;
        const int minage = p.minage;
        std::vector<short> &bcnt = awb_pack_scratch();
        bcnt.resize(V);
        short *bc = bcnt.data();
        unsigned char nch[AWB_MAXV];
        memset(nch, 0, (size_t) V);
        for (int t = 0; t < T; t++) wrow[t] = 0;
        int S_ = 0, band_ = 0, tp = 0, mc = maxcnt, nroots = 0;
        unsigned bad = 0;
        bool odd = false;
        for (int i = 0; i < V; i++) {
            const int pa = parent[i], a = age[i];
            const bool isroot = pa == -1;
            const int pidx = isroot ? i : pa;
            if ((unsigned) pidx >= (unsigned) V) { odd = true; break; }
            const int pa_age = age[pidx];
            if (((unsigned) a > (unsigned) (T - 2)) | ((unsigned) pa_age > (unsigned) (T - 2))) {
                odd = true;
                break;
            }
            nroots += isroot;
            nch[pidx] = (unsigned char) (nch[pidx] + !isroot);
            bad |= (unsigned) (!isroot & (pa == i)) | (unsigned) (pa_age < a);
            const int lo = a > minage ? a : minage;
            const int hi = isroot ? T - 2 : pa_age;
            const int cnt = hi - lo + 1;
            if (cnt <= 0 || cnt > 32) { odd = true; break; }
            S_ += cnt;
            band_ += cnt * cnt;
            bc[i] = (short) cnt;
            wrow[lo]++;
            wrow[hi + 1]--;
            mc = cnt > mc ? cnt : mc;
            const int tp32 = (tp + 31) & ~31;
            tp = ((tp & 31) + cnt > 32 ? tp32 : tp) + cnt;
        }
        if (!odd) {
            // leaves first, every other node with exactly two children
            const int nl = p.nleaves;
            for (int i = 0; i < V; i++)
                bad |= (unsigned) (nch[i] != (i < nl ? 0 : 2));
            if (!bad && awb_parent_cycle(parent, age, V))
                bad = 1;
        }
        if (!odd && !bad && nroots == 1) {
            S = S_;
            band = band_;
            tpos = tp;
            maxcnt = mc;
            for (int t = 1; t < T; t++) wrow[t] += wrow[t - 1];
            return true;
        }
    }
    return awb_count_states_checked(p, b, c0, c1, stack, ignore, S, band, tpos,
                                    maxcnt, err, wrow);
}

inline size_t awb_align(size_t x) { return (x + 255) & ~(size_t) 255; }

// Checkpointed table: the segment tables sit at the end of the window's
// sub-arena -- `nslots` of them (fw, then fsum).  One is needed; the batch gives
// more when the device has room, and the last `nslots` segments of the forward
// pass then stay resident for the traceback instead of being rebuilt
// (awb_api.cu batch_bind).  Returns the window's arena bytes.
inline size_t awb_layout_slot_bytes(const AwbLayout &L)
{
    const int T = L.T;
    return awb_align((size_t) L.seg_doubles * sizeof(double)) +
        awb_align((size_t) L.seg_sites * (T > 1 ? T - 1 : 1) * sizeof(double));
}

inline size_t awb_layout_place_slots(AwbLayout &L, int nslots, bool with_band)
{
    const int T = L.T;
    const size_t base = with_band ? L.total_bytes_noslots : L.bytes_before_band;
    L.nslots = nslots;
    L.o_fw = base;
    L.o_fsum = base + awb_align((size_t) nslots * L.seg_doubles * sizeof(double));
    L.slots_at = base;
    return L.o_fsum +
        awb_align((size_t) nslots * L.seg_sites * (T > 1 ? T - 1 : 1) * sizeof(double));
}

// Build the layout.  Returns false and sets err on invalid input.
// seg_cap > 0 selects the checkpointed table: the forward table is not kept
// whole; the window is cut into segments of whole blocks whose tables hold at
// most seg_cap doubles, and only one segment's table is resident at a time
// (see awb_api.cu).
inline bool awb_layout_build(const awb_problem &p, int keep_debug, AwbLayout &L,
                             std::string &err, long long seg_cap = 0)
{
    const int B = p.ntrees, V = p.nnodes, T = p.ntimes;
    if (T < 3 || T > AWB_MAXT) { err = "ntimes must be in [3, 64]"; return false; }
    if (V < 1 || V > AWB_MAXV) { err = "nnodes must be in [1, 1024]"; return false; }
    if (B < 1) { err = "ntrees must be >= 1"; return false; }
    if (!p.times || !p.popsizes || (!p.seqs && !p.var_cols) || !p.seqids || !p.ptrees ||
        !p.ages || !p.sprs || !p.blocklens) {
        err = "null input array";
        return false;
    }
    L.packed = p.var_cols != 0;
    if (L.packed) {
        if (p.nvar < 0 || (p.nvar > 0 && !p.var_pos)) { err = "bad variant columns"; return false; }
        for (int i = 0; i < p.nvar; i++)
            if (p.var_pos[i] < 0 || p.var_pos[i] >= p.seqlen ||
                (i > 0 && p.var_pos[i] <= p.var_pos[i - 1])) {
                err = "variant positions must be ascending, unique and inside the sequence";
                return false;
            }
    }
    if (p.nleaves != (V + 1) / 2) { err = "nleaves != (nnodes+1)/2"; return false; }
    if (p.unphased) {
        const int nr = p.nleaves + (p.internal ? 0 : 1);
        if (p.phase_row1 < 0 || p.phase_row1 >= nr || p.phase_row2 < 0 ||
            p.phase_row2 >= nr || p.phase_row1 == p.phase_row2) {
            err = "phase_row1 / phase_row2 must be two different rows of the leaf order";
            return false;
        }
        if (p.infsites_penalty > 0.0 && p.infsites_penalty < 1.0) {
            err = "unphased emissions are not combined with the infinite-sites penalty";
            return false;
        }
    }
    for (int i = 0; i + 1 < T; i++)
        if (!(p.times[i + 1] > p.times[i])) { err = "times must increase"; return false; }
    L.B = B; L.V = V; L.T = T;
    L.keep_debug = keep_debug;
    awb_model_fill(L.model, p);

    // rows compared for invariance (matrices.cpp:30-34 / :108-112)
    L.rowidx.clear();
    for (int i = 0; i < p.nleaves; i++) {
        if (p.seqids[i] < 0 || p.seqids[i] >= p.nseqs) { err = "bad seqid"; return false; }
        L.rowidx.push_back(p.seqids[i]);
    }
    if (!p.internal) {
        if (p.new_chrom < 0 || p.new_chrom >= p.nseqs) { err = "bad new_chrom"; return false; }
        L.rowidx.push_back(p.new_chrom);
    }
    L.nrows = (int) L.rowidx.size();

    L.nstates.assign(B, 0);
    L.block_start.assign(B + 1, 0);
    L.row_off.assign(B + 1, 0);
    L.fw_off.assign(B + 1, 0);
    L.band_off.assign(B + 1, 0);
    L.ent_off.assign(B + 1, 0);
    L.sw1_off.assign(B + 1, 0);
    L.trow_off.assign(B + 1, 0);
    L.maxS = 1;
    L.maxNS = 32;
    L.zcap = 64;
    L.maxcnt = 1;
    L.maxband = 0;
    L.states_sites = 0;
    std::vector<int> c0(V), c1(V), stack(V + 2);
    std::vector<char> ignore(V);
    // Upper bound on lnB (trans.cpp:44-71) = cumulative coalescent rate + log of
    // a time step: K1's linear-domain vectors hold exp(lnB), so beyond ~700
    // they overflow.  First with every lineage present at every time; only when
    // that fails, block by block from the states per time row (>= the lineages
    // of that time).
    double rate[AWB_MAXT];
    double gbound = 0.0, maxstep = 0.0;
    for (int t = 0; t < T - 1; t++) {
        rate[t] = (L.model.coal_time_steps[2 * t] + L.model.coal_time_steps[2 * t + 1]) /
            (2.0 * p.popsizes[t]);
        gbound += rate[t] * (p.nleaves + 1);
        if (L.model.time_steps[t] > maxstep) maxstep = L.model.time_steps[t];
    }
    const double lnslack = log(maxstep * (V + 2.0)) + log((double) T);
    const double lnlimit = 680.0;
    const bool block_bound = !(gbound + lnslack <= lnlimit);
    L.lin_unsafe = 0;
    for (int b = 0; b < B; b++) {
        int S = 0, band = 0, tpos = 1;
        int wrow[AWB_MAXT + 1];
        if (!awb_count_states(p, b, c0, c1, stack, ignore, S, band, tpos,
                              L.maxcnt, err, wrow))
            return false;
        if (block_bound && !L.lin_unsafe && S > 0) {
            double c = 0.0;
            for (int t = 0; t < T - 1; t++)
                c += rate[t] * (wrow[t] > 0 ? wrow[t] : p.nleaves + 1);
            if (!(c + lnslack <= lnlimit))
                L.lin_unsafe = 1;
        }
        {
            // the scribes' padded column (awb_scribe_plan): the slots per lane
            // are at most CHub = the even count for which the rows surely fit
            // (lanes <= S/CH + rows + a warp-boundary gap), hence the column is
            // at most S + rows*CHub; the exact plan only for a block whose
            // bound is large and could raise the maximum
            const int rows = T - 1;
            int maxw = 1;
            for (int t = 0; t < rows; t++)
                if (wrow[t] > maxw) maxw = wrow[t];
            const int spare = AWB_NSCRIBE - rows - 1;
            long long chub = spare > 0 ? (S + 2 * maxw + spare - 1) / spare : maxw;
            if (chub * 32 < maxw) chub = (maxw + 31) / 32;
            if (chub > maxw) chub = maxw;
            chub = (chub + 1) & ~1ll;
            if (chub < 2) chub = 2;
            const long long zub = S + rows * chub;
            if (zub > L.zcap) {
                if (zub <= 2 * (long long) ((tpos + 31) & ~31) + 64) {
                    L.zcap = (int) zub;     // a modest buffer: the bound will do
                } else {
                    int CH;
                    const int z = awb_scribe_plan(wrow, rows, CH);
                    if (CH > 65535) { err = "block too wide for the forward kernel"; return false; }
                    if (z > L.zcap) L.zcap = z;
                }
            }
        }
        const int NSb = (tpos + 31) & ~31;
        L.trow_off[b + 1] = L.trow_off[b] + NSb;
        if (NSb > L.maxNS) {
            // K1 packs first-fit-decreasing when that is tighter; the exact
            // count only matters for a block that could raise the maximum
            const int ffd = S > 0 ?
                awb_pack_branches(awb_pack_scratch().data(), V) : 32;
            const int used = ffd < NSb ? ffd : NSb;
            if (used > L.maxNS) L.maxNS = used;
        }
        if (S > AWB_MAXS) {
            err = "block with more than 2048 states is not supported by this build";
            return false;
        }
        if (p.blocklens[b] < 1) { err = "blocklen must be >= 1"; return false; }
        if (!p.internal && S == 0) { err = "external block without states"; return false; }
        const int S1 = S > 0 ? S : 1;
        L.nstates[b] = S;
        L.block_start[b + 1] = L.block_start[b] + p.blocklens[b];
        L.row_off[b + 1] = L.row_off[b] + S1;
        L.fw_off[b + 1] = L.fw_off[b] + (long long) S1 * p.blocklens[b];
        L.band_off[b + 1] = L.band_off[b] + band;
        const int prevS1 = b > 0 ? (L.nstates[b - 1] > 0 ? L.nstates[b - 1] : 1) : 0;
        L.ent_off[b + 1] = L.ent_off[b] + (b > 0 ? prevS1 + T + 4 : 0);
        L.sw1_off[b + 1] = L.sw1_off[b] + (b > 0 ? prevS1 : 0);
        if (S1 > L.maxS) L.maxS = S1;
        if (band > L.maxband) L.maxband = band;
        L.states_sites += (double) S * p.blocklens[b];

        if (b > 0) {
            const int *spr = p.sprs + 4 * (size_t) b;
            const int *lp = p.ptrees + (size_t) (b - 1) * V;
            if (spr[0] < 0 || spr[0] >= V || spr[2] < 0 || spr[2] >= V ||
                spr[1] < 0 || spr[3] < spr[1] || spr[3] > T - 1 || lp[spr[0]] < 0) {
                err = "tree " + std::to_string(b) + ": invalid SPR";
                return false;
            }
            if (p.mappings) {
                const int broken = lp[spr[0]];
                const int *mp = p.mappings + (size_t) b * V;
                for (int x = 0; x < V; x++) {
                    if (mp[x] < -1 || mp[x] >= V || (mp[x] == -1) != (x == broken)) {
                        err = "tree " + std::to_string(b) +
                            ": invalid node mapping (only the broken node maps to -1)";
                        return false;
                    }
                }
            }
        }
    }
    L.n = L.block_start[B];
    if (p.start_coord < 0 || p.start_coord + L.n > p.seqlen) {
        err = "blocks exceed the sequence length";
        return false;
    }
    L.has_subtree_roots = p.internal && p.subtree_roots;

    // node mappings: the caller's, or identity except the broken node
    // (make_node_mapping, local_tree.h:767-776), which K1 writes on the device
    L.gen_mappings = p.mappings ? 0 : 1;

    // ---- segments of the checkpointed table
    L.ckpt = seg_cap > 0 ? 1 : 0;
    L.nseg = 0;
    L.seg_start.clear();
    L.seg_doubles = 0;
    L.seg_sites = 0;
    if (L.ckpt) {
        // the fewest segments the capacity allows, then cut by SITES (a
        // segment's run time is proportional to its sites, and every launch
        // waits for the slowest window), never exceeding the capacity
        const long long total = L.fw_off[B] + L.maxS;
        long long want = (total + seg_cap - 1) / seg_cap;
        if (want < 1) want = 1;
        const int target = (int) ((L.n + want - 1) / want);
        int b0 = 0;
        while (b0 < B) {
            int b1 = b0 + 1;
            // a segment holds blocks [b0, b1) plus the first row of block b1
            while (b1 < B &&
                   (L.fw_off[b1 + 1] - L.fw_off[b0]) + L.maxS <= seg_cap &&
                   L.block_start[b1] - L.block_start[b0] < target)
                b1++;
            L.seg_start.push_back(b0);
            const long long d = (L.fw_off[b1] - L.fw_off[b0]) + L.maxS;
            const int ns = (L.block_start[b1] - L.block_start[b0]) + 1;
            if (d > L.seg_doubles) L.seg_doubles = d;
            if (ns > L.seg_sites) L.seg_sites = ns;
            b0 = b1;
        }
        L.nseg = (int) L.seg_start.size();
        L.seg_start.push_back(B);
    }

    // ---- arena
    size_t off = 0;
    const size_t rows = (size_t) L.row_off[B];
    const size_t BV = (size_t) B * V;
#define AWB_PLACE(name, bytes) do { L.name = off; off = awb_align(off + (bytes)); } while (0)
    AWB_PLACE(o_ptrees, BV * sizeof(int));
    AWB_PLACE(o_ages, BV * sizeof(int));
    AWB_PLACE(o_mappings, BV * sizeof(int));
    AWB_PLACE(o_seqs, L.packed ? (size_t) p.nvar * p.nseqs : (size_t) p.nseqs * p.seqlen);
    AWB_PLACE(o_varpos, L.packed ? (size_t) p.nvar * sizeof(int) : 0);
    // the small inputs and the arrays made above, next to each other and in the
    // order of L.copies: they go up as ONE copy from a pinned staging buffer
    // (awb_api.cu)
    AWB_PLACE(o_sprs, (size_t) B * 4 * sizeof(int));
    AWB_PLACE(o_blocklens, (size_t) B * sizeof(int));
    AWB_PLACE(o_subtree_roots, (size_t) B * sizeof(int));
    AWB_PLACE(o_rowidx, (size_t) L.nrows * sizeof(int));
    AWB_PLACE(o_block_start, (size_t) (B + 1) * sizeof(int));
    AWB_PLACE(o_nstates, (size_t) B * sizeof(int));
    AWB_PLACE(o_row_off, (size_t) (B + 1) * sizeof(long long));
    AWB_PLACE(o_fw_off, (size_t) (B + 1) * sizeof(long long));
    AWB_PLACE(o_band_off, (size_t) (B + 1) * sizeof(long long));
    AWB_PLACE(o_ent_off, (size_t) (B + 1) * sizeof(long long));
    AWB_PLACE(o_sw1_off, (size_t) (B + 1) * sizeof(long long));
    AWB_PLACE(o_trow_off, (size_t) (B + 1) * sizeof(long long));
    AWB_PLACE(o_ptab, (size_t) (T * T + T) * 2 * sizeof(double));
    if (L.ckpt)
        AWB_PLACE(o_seg_start, (size_t) (L.nseg + 1) * sizeof(int));
    else
        L.o_seg_start = 0;
    AWB_PLACE(o_tmap, (size_t) L.trow_off[B] * sizeof(short));
    AWB_PLACE(o_iperm, rows * sizeof(short));
    AWB_PLACE(o_st_age, rows);
    AWB_PLACE(o_lin, (size_t) B * 7 * T * sizeof(double));
    AWB_PLACE(o_sc_start, (size_t) B * AWB_NSCRIBE * sizeof(short));
    AWB_PLACE(o_sc_cnt, (size_t) B * AWB_NSCRIBE * sizeof(short));
    AWB_PLACE(o_sc_row, (size_t) B * AWB_NSCRIBE);
    AWB_PLACE(o_sc_stride, (size_t) B * AWB_NSCRIBE);
    AWB_PLACE(o_sc_ch, (size_t) B * 2);
    AWB_PLACE(o_st_node, rows * sizeof(short));
    AWB_PLACE(o_st_time, rows);
    AWB_PLACE(o_perm, rows * sizeof(short));
    AWB_PLACE(o_inv_emit, rows * sizeof(double));
    AWB_PLACE(o_tmatrix, (size_t) B * T * T * sizeof(double));
    AWB_PLACE(o_tmvec, (size_t) B * AWB_TM_NVEC * T * sizeof(double));
    AWB_PLACE(o_rowstart, (size_t) B * (T + 1) * sizeof(short));
    AWB_PLACE(o_pstart, (size_t) B * (T + 1) * sizeof(short));
    AWB_PLACE(o_slotrow, (size_t) B * (T + 32));
    AWB_PLACE(o_node_first, BV * sizeof(short));
    AWB_PLACE(o_node_cnt, BV * sizeof(short));
    AWB_PLACE(o_child0, BV * sizeof(short));
    AWB_PLACE(o_child1, BV * sizeof(short));
    AWB_PLACE(o_order, BV * sizeof(short));
    AWB_PLACE(o_lstart, (size_t) B * (V + 2) * sizeof(short));
    AWB_PLACE(o_root, (size_t) B * sizeof(short));
    AWB_PLACE(o_lineages, (size_t) B * 3 * T * sizeof(int));
    AWB_PLACE(o_treelen, (size_t) B * sizeof(double));
    AWB_PLACE(o_tm_minage, (size_t) B * sizeof(int));
    AWB_PLACE(o_sw_start, rows * sizeof(short));
    AWB_PLACE(o_sw_cnt, rows * sizeof(short));
    AWB_PLACE(o_sw_src, (size_t) L.ent_off[B] * sizeof(short) + 8);
    AWB_PLACE(o_sw_prob, (size_t) L.ent_off[B] * sizeof(double) + 8);
    AWB_PLACE(o_sw_determ, (size_t) L.sw1_off[B] * sizeof(int) + 8);
    if (keep_debug) {
        AWB_PLACE(o_sw_determprob, (size_t) L.sw1_off[B] * sizeof(double) + 8);
        AWB_PLACE(o_sw_recombrow, rows * sizeof(double));
        AWB_PLACE(o_sw_recoalrow, rows * sizeof(double));
        AWB_PLACE(o_sw_recombsrc, (size_t) B * sizeof(int));
        AWB_PLACE(o_sw_recoalsrc, (size_t) B * sizeof(int));
    } else {
        L.o_sw_determprob = L.o_sw_recombrow = 0;
        L.o_sw_recoalrow = L.o_sw_recombsrc = L.o_sw_recoalsrc = 0;
    }
    AWB_PLACE(o_kind, (size_t) L.n + 8);     // the forward kernel reads a few sites ahead
    if (L.ckpt) {
        // the first column of every segment; the segment tables themselves
        // (fw, fsum) are placed behind everything else by awb_layout_place_slots
        AWB_PLACE(o_ckptcol, (size_t) (L.nseg + 1) * L.maxS * sizeof(double));
        L.o_fw = L.o_fsum = 0;
    } else {
        AWB_PLACE(o_fw, (size_t) L.fw_off[B] * sizeof(double));
        // per-site, per-time sums of the stored forward column (traceback)
        AWB_PLACE(o_fsum, (size_t) L.n * (T > 1 ? T - 1 : 1) * sizeof(double));
        L.o_ckptcol = 0;
    }
    AWB_PLACE(o_path, (size_t) L.n * sizeof(int));
    AWB_PLACE(o_rand, (size_t) L.n * sizeof(int));
    AWB_PLACE(o_logz, sizeof(double));
    AWB_PLACE(o_sink, (size_t) 1024 * sizeof(double));   // forward kernel: discarded stores
    AWB_PLACE(o_status, sizeof(int));
    // the tmatrix2 band and the per-state slot tables are only read by the
    // generic forward kernel: they come last so that a batch on the fast path
    // can leave them out of its arena
    L.bytes_before_band = off;
    AWB_PLACE(o_pslot, rows * sizeof(short));
    AWB_PLACE(o_band_j1, rows * sizeof(short));
    AWB_PLACE(o_band_len, rows);
    AWB_PLACE(o_band_boff, rows * sizeof(int));
    AWB_PLACE(o_band, (size_t) L.band_off[B] * sizeof(double) + 8);
#undef AWB_PLACE
    L.total_bytes = L.total_bytes_noslots = off;
    L.nslots = 1;
    L.slots_at = 0;
    if (L.ckpt)
        L.total_bytes = awb_layout_place_slots(L, 1, true);

    // ---- branch-probability table of the emission kernel (emit.cpp:86-116)
    L.ptab.assign((size_t) (T * T + T) * 2, 0.0);
    for (int y = 0; y < T; y++)
        for (int x = 0; x < T; x++) {
            const double t = fmax(p.times[x] - p.times[y], L.model.mintime);
            L.ptab[((size_t) y * T + x) * 2 + 0] = awb_prob_branch(t, L.model.mu, true);
            L.ptab[((size_t) y * T + x) * 2 + 1] = awb_prob_branch(t, L.model.mu, false);
        }
    for (int x = 0; x < T; x++) {
        const double t = fmax(p.times[x] - 0.0, L.model.mintime);
        L.ptab[((size_t) T * T + x) * 2 + 0] = awb_prob_branch(t, L.model.mu, true);
        L.ptab[((size_t) T * T + x) * 2 + 1] = awb_prob_branch(t, L.model.mu, false);
    }

    // ---- input copies
    // (own: 0 = a large caller array, copied from where it is; 1 = a small
    // caller array, 2 = made by the layout -- both staged when the batch has a
    // pinned staging buffer; the staged ones in arena order)
    L.copies.clear();
    L.copies.push_back({ L.o_ptrees, p.ptrees, BV * sizeof(int), 0 });
    L.copies.push_back({ L.o_ages, p.ages, BV * sizeof(int), 0 });
    if (p.mappings)
        L.copies.push_back({ L.o_mappings, p.mappings, BV * sizeof(int), 0 });
    if (L.packed) {
        if (p.nvar > 0) {
            L.copies.push_back({ L.o_seqs, p.var_cols, (size_t) p.nvar * p.nseqs, 0 });
            L.copies.push_back({ L.o_varpos, p.var_pos, (size_t) p.nvar * sizeof(int), 0 });
        }
    } else {
        L.copies.push_back({ L.o_seqs, p.seqs, (size_t) p.nseqs * p.seqlen, 0 });
    }
    L.copies.push_back({ L.o_sprs, p.sprs, (size_t) B * 4 * sizeof(int), 1 });
    L.copies.push_back({ L.o_blocklens, p.blocklens, (size_t) B * sizeof(int), 1 });
    if (L.has_subtree_roots)
        L.copies.push_back({ L.o_subtree_roots, p.subtree_roots, (size_t) B * sizeof(int), 1 });
    L.copies.push_back({ L.o_rowidx, L.rowidx.data(), (size_t) L.nrows * sizeof(int), 2 });
    L.copies.push_back({ L.o_block_start, L.block_start.data(), (size_t) (B + 1) * sizeof(int), 2 });
    L.copies.push_back({ L.o_nstates, L.nstates.data(), (size_t) B * sizeof(int), 2 });
    L.copies.push_back({ L.o_row_off, L.row_off.data(), (size_t) (B + 1) * sizeof(long long), 2 });
    L.copies.push_back({ L.o_fw_off, L.fw_off.data(), (size_t) (B + 1) * sizeof(long long), 2 });
    L.copies.push_back({ L.o_band_off, L.band_off.data(), (size_t) (B + 1) * sizeof(long long), 2 });
    L.copies.push_back({ L.o_ent_off, L.ent_off.data(), (size_t) (B + 1) * sizeof(long long), 2 });
    L.copies.push_back({ L.o_sw1_off, L.sw1_off.data(), (size_t) (B + 1) * sizeof(long long), 2 });
    L.copies.push_back({ L.o_trow_off, L.trow_off.data(), (size_t) (B + 1) * sizeof(long long), 2 });
    L.copies.push_back({ L.o_ptab, L.ptab.data(), L.ptab.size() * sizeof(double), 2 });
    if (L.ckpt)
        L.copies.push_back({ L.o_seg_start, L.seg_start.data(),
                             (size_t) (L.nseg + 1) * sizeof(int), 2 });
    return true;
}

// Point an AwbChain at an arena (device or host base pointer).
inline void awb_layout_bind(const AwbLayout &L, const awb_problem &p, char *base,
                            AwbChain &ch)
{
    memset(&ch, 0, sizeof(ch));
    ch.model = L.model;
    ch.internal = p.internal;
    ch.minage = p.minage;
    ch.nleaves = p.nleaves;
    ch.nrows = L.nrows;
    ch.nseqs = p.nseqs;
    ch.seqlen = p.seqlen;
    ch.ntrees = L.B;
    ch.nnodes = L.V;
    ch.nsites = L.n;
    ch.start_coord = p.start_coord;
    ch.maxS = L.maxS;
    ch.maxband = L.maxband;
    ch.maxNS = L.maxNS;
    ch.zcap = L.zcap;
    ch.maxcnt = L.maxcnt;
    ch.keep_debug = L.keep_debug;
    ch.need_band = 1;
    ch.gen_mappings = L.gen_mappings;
    ch.ckpt = L.ckpt;
    ch.nseg = L.nseg;
    ch.nslots = L.nslots;
    ch.nrot = 1;                        // (the batch sets it: awb_api.cu)
    ch.seg_doubles = L.seg_doubles;
    ch.seg_sites = L.seg_sites;
    ch.seg_start = L.ckpt ? (const int *) (base + L.o_seg_start) : 0;
    ch.ckptcol = L.ckpt ? (double *) (base + L.o_ckptcol) : 0;
    ch.last_state = -1;
#define AWB_P(type, field, off) ch.field = (type) (base + L.off)
    AWB_P(const int *, ptrees, o_ptrees);
    AWB_P(const int *, ages, o_ages);
    AWB_P(const int *, sprs, o_sprs);
    AWB_P(int *, mappings, o_mappings);
    AWB_P(const int *, blocklens, o_blocklens);
    ch.subtree_roots = L.has_subtree_roots ?
        (const int *) (base + L.o_subtree_roots) : 0;
    AWB_P(const int *, rowidx, o_rowidx);
    if (L.packed) {
        ch.seqs = 0;
        ch.nvar = p.nvar;
        AWB_P(const unsigned char *, var_cols, o_seqs);
        AWB_P(const int *, var_pos, o_varpos);
    } else {
        AWB_P(const unsigned char *, seqs, o_seqs);
    }
    ch.phase_row1 = ch.phase_row2 = -1;
    if (p.unphased) {
        ch.phase_row1 = p.phase_row1;
        ch.phase_row2 = p.phase_row2;
    }
    ch.default_char = p.default_char ? p.default_char : 'A';
    ch.infsites_penalty = (p.infsites_penalty > 0.0 && p.infsites_penalty < 1.0) ?
        p.infsites_penalty : 1.0;
    AWB_P(const int *, block_start, o_block_start);
    AWB_P(const int *, nstates, o_nstates);
    AWB_P(const long long *, row_off, o_row_off);
    AWB_P(const long long *, fw_off, o_fw_off);
    AWB_P(const long long *, band_off, o_band_off);
    AWB_P(const long long *, ent_off, o_ent_off);
    AWB_P(const long long *, sw1_off, o_sw1_off);
    AWB_P(const long long *, trow_off, o_trow_off);
    AWB_P(unsigned short *, tmap, o_tmap);
    AWB_P(unsigned short *, iperm, o_iperm);
    AWB_P(signed char *, st_age, o_st_age);
    AWB_P(double *, lin, o_lin);
    AWB_P(const double *, ptab, o_ptab);
    AWB_P(unsigned short *, sc_start, o_sc_start);
    AWB_P(unsigned short *, sc_cnt, o_sc_cnt);
    AWB_P(unsigned char *, sc_row, o_sc_row);
    AWB_P(unsigned char *, sc_stride, o_sc_stride);
    AWB_P(unsigned short *, sc_ch, o_sc_ch);
    AWB_P(short *, st_node, o_st_node);
    AWB_P(signed char *, st_time, o_st_time);
    AWB_P(unsigned short *, perm, o_perm);
    AWB_P(unsigned short *, pslot, o_pslot);
    AWB_P(unsigned short *, band_j1, o_band_j1);
    AWB_P(unsigned char *, band_len, o_band_len);
    AWB_P(int *, band_boff, o_band_boff);
    AWB_P(double *, inv_emit, o_inv_emit);
    AWB_P(double *, band, o_band);
    AWB_P(double *, tmatrix, o_tmatrix);
    AWB_P(double *, tmvec, o_tmvec);
    AWB_P(unsigned short *, rowstart, o_rowstart);
    AWB_P(unsigned short *, pstart, o_pstart);
    AWB_P(unsigned char *, slotrow, o_slotrow);
    ch.slotcap = L.T + 32;
    AWB_P(short *, node_first, o_node_first);
    AWB_P(short *, node_cnt, o_node_cnt);
    AWB_P(short *, child0, o_child0);
    AWB_P(short *, child1, o_child1);
    AWB_P(short *, order, o_order);
    AWB_P(short *, lstart, o_lstart);
    AWB_P(short *, root, o_root);
    AWB_P(int *, lineages, o_lineages);
    AWB_P(double *, treelen, o_treelen);
    AWB_P(int *, tm_minage, o_tm_minage);
    AWB_P(unsigned short *, sw_start, o_sw_start);
    AWB_P(unsigned short *, sw_cnt, o_sw_cnt);
    AWB_P(unsigned short *, sw_src, o_sw_src);
    AWB_P(double *, sw_prob, o_sw_prob);
    AWB_P(int *, sw_determ, o_sw_determ);
    if (L.keep_debug) {
        AWB_P(double *, sw_determprob, o_sw_determprob);
        AWB_P(double *, sw_recombrow, o_sw_recombrow);
        AWB_P(double *, sw_recoalrow, o_sw_recoalrow);
        AWB_P(int *, sw_recombsrc, o_sw_recombsrc);
        AWB_P(int *, sw_recoalsrc, o_sw_recoalsrc);
    }
    AWB_P(unsigned char *, kind, o_kind);
    AWB_P(double *, fw, o_fw);
    AWB_P(int *, path, o_path);
    AWB_P(const int *, rand_ints, o_rand);
    AWB_P(double *, logz, o_logz);
    AWB_P(double *, sink, o_sink);
    AWB_P(double *, fsum, o_fsum);
    AWB_P(int *, status, o_status);
#undef AWB_P
}

// Named access to arena arrays (tests / AWB_KEEP_DEBUG).  Returns false if the
// name is unknown or the array is not kept.
inline bool awb_layout_find(const AwbLayout &L, const char *name, size_t &off,
                            size_t &bytes)
{
    const size_t rows = (size_t) L.row_off[L.B];
    const size_t BV = (size_t) L.B * L.V;
    const size_t B = L.B, T = L.T;
    struct Ent { const char *name; size_t off; size_t bytes; bool dbg; };
    const Ent table[] = {
        { "st_node", L.o_st_node, rows * 2, false },
        { "st_time", L.o_st_time, rows, false },
        { "perm", L.o_perm, rows * 2, false },
        { "pslot", L.o_pslot, rows * 2, false },
        { "band_j1", L.o_band_j1, rows * 2, false },
        { "band_len", L.o_band_len, rows, false },
        { "band_boff", L.o_band_boff, rows * 4, false },
        { "inv_emit", L.o_inv_emit, rows * 8, false },
        { "band", L.o_band, (size_t) L.band_off[L.B] * 8, false },
        { "tmatrix", L.o_tmatrix, B * T * T * 8, false },
        { "tmvec", L.o_tmvec, B * AWB_TM_NVEC * T * 8, false },
        { "rowstart", L.o_rowstart, B * (T + 1) * 2, false },
        { "pstart", L.o_pstart, B * (T + 1) * 2, false },
        { "tmap", L.o_tmap, (size_t) L.trow_off[L.B] * 2, false },
        { "iperm", L.o_iperm, rows * 2, false },
        { "lin", L.o_lin, B * 7 * T * 8, false },
        { "slotrow", L.o_slotrow, B * (T + 32), false },
        { "node_first", L.o_node_first, BV * 2, false },
        { "node_cnt", L.o_node_cnt, BV * 2, false },
        { "child0", L.o_child0, BV * 2, false },
        { "child1", L.o_child1, BV * 2, false },
        { "order", L.o_order, BV * 2, false },
        { "root", L.o_root, B * 2, false },
        { "lineages", L.o_lineages, B * 3 * T * 4, false },
        { "treelen", L.o_treelen, B * 8, false },
        { "tm_minage", L.o_tm_minage, B * 4, false },
        { "sw_start", L.o_sw_start, rows * 2, false },
        { "sw_cnt", L.o_sw_cnt, rows * 2, false },
        { "sw_src", L.o_sw_src, (size_t) L.ent_off[L.B] * 2, false },
        { "sw_prob", L.o_sw_prob, (size_t) L.ent_off[L.B] * 8, false },
        { "sw_determ", L.o_sw_determ, (size_t) L.sw1_off[L.B] * 4, false },
        { "sw_determprob", L.o_sw_determprob, (size_t) L.sw1_off[L.B] * 8, true },
        { "sw_recombrow", L.o_sw_recombrow, rows * 8, true },
        { "sw_recoalrow", L.o_sw_recoalrow, rows * 8, true },
        { "sw_recombsrc", L.o_sw_recombsrc, B * 4, true },
        { "sw_recoalsrc", L.o_sw_recoalsrc, B * 4, true },
        { "kind", L.o_kind, (size_t) L.n, false },
        { "fw", L.o_fw, (size_t) L.fw_off[L.B] * 8, false },
        { "path", L.o_path, (size_t) L.n * 4, false },
        { "fsum", L.o_fsum, (size_t) L.n * (T > 1 ? T - 1 : 1) * 8, false },
        { "sink", L.o_sink, 1024 * 8, false },
        { "ent_off", L.o_ent_off, (B + 1) * 8, false },
        { "band_off", L.o_band_off, (B + 1) * 8, false },
    };
    for (size_t i = 0; i < sizeof(table) / sizeof(table[0]); i++) {
        if (strcmp(table[i].name, name) == 0) {
            if (table[i].dbg && !L.keep_debug)
                return false;
            off = table[i].off;
            bytes = table[i].bytes;
            return true;
        }
    }
    return false;
}

#endif // AWB_LAYOUT_H
