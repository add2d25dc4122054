// awb_emit.cuh -- emissions of the threading HMM (reference emit.cpp:650-869),
// with the infinite-sites penalty (:457-589, :848-862) and the integration
// over the two phasings of one individual (:705-742, :834-842).
//
// Invariant sites (~97 % of compressed sites) share one per-state constant per
// block (inv_emit, computed by awb_block_setup); masked sites emit 1.  Only
// VARIANT sites need Felsenstein pruning: awb_emit_site computes the inner
// (post-order) and outer (pre-order) partial likelihoods of one site and the
// per-state emission with the new branch attached, and stores the row into
// the forward table's own slab for that site (fw[i][*]); the forward kernel
// reads it there before overwriting the row with the forward column, so the
// emissions never occupy separate HBM.
//
// awb_emit_site is lane-cooperative: `nlanes` workers (a warp on the GPU, 1 on
// the host emulation) call it with the same arguments and their own `lane`.
#ifndef AWB_EMIT_CUH
#define AWB_EMIT_CUH

#include "awb_common.cuh"

#if defined(__CUDA_ARCH__)
#define AWB_LANESYNC() __syncwarp()
#else
#define AWB_LANESYNC() ((void) 0)
#endif

// emit.cpp:30-58 (find_invariant_sites, find_masked_sites) for one site
AWB_HD inline void awb_site_kind(const AwbChain &ch, int i, int vc = -1)
{
    const size_t col = (size_t) ch.start_coord + i;
    const unsigned char c = awb_seq_at(ch, ch.rowidx[0], col, vc);
    bool mut = false;
    for (int r = 1; r < ch.nrows; r++) {
        if (awb_seq_at(ch, ch.rowidx[r], col, vc) != c) {
            mut = true;
            break;
        }
    }
    ch.kind[i] = mut ? AWB_SITE_VARIANT :
        (c == 'N' ? AWB_SITE_MASKED : AWB_SITE_INVARIANT);
}

AWB_HD inline void awb_leaf_row(unsigned char c, double *row)
{
    int x = -1;                       // seq.cpp:15-43 (dna2int)
    switch (c) {
    case 'A': case 'a': x = 0; break;
    case 'C': case 'c': x = 1; break;
    case 'G': case 'g': x = 2; break;
    case 'T': case 't': x = 3; break;
    }
    if (x < 0) {                      // 'N' (emit.cpp:162-166); other symbols as N
        row[0] = row[1] = row[2] = row[3] = 1.0;
    } else {
        row[0] = row[1] = row[2] = row[3] = 0.0;
        row[x] = 1.0;
    }
}

// Binary search: block containing site i (block_start is [B+1], ascending)
AWB_HD inline int awb_find_block(const AwbChain &ch, int i)
{
    int lo = 0, hi = ch.ntrees - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ch.block_start[mid] <= i) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// scratch per worker group: inner[4V], outer[4V], mut[V], nomut[V] doubles and
// inmain[V] bytes
// ... and, for the infinite-sites rule, the parsimony sets sets[V], anc[V] bytes
AWB_HD inline size_t awb_emit_scratch_bytes(int V)
{
    return (size_t) V * (10 * sizeof(double)) + ((3 * (size_t) V + 7) & ~(size_t) 7);
}

AWB_HD inline int awb_base_bit(unsigned char c)
{
    switch (c) {                      // 1 << dna2int[c]; 0 for 'N' (emit.cpp:916-920)
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 4;
    case 'T': case 't': return 8;
    }
    return 0;
}

AWB_HD inline int awb_lanes_or(int x)
{
#if defined(__CUDA_ARCH__)
    return (int) __reduce_or_sync(0xffffffffu, (unsigned) x);
#else
    return x;
#endif
}

// tree arrays of block b: parent/age (int), child0/child1/order (short).  The
// CUDA kernel stages them in shared memory (the pruning loops walk them with
// dependent accesses); the host emulation passes the global arrays.
// swap: the rows of the two haplotypes of the unphased individual change
// places (the other phasing of a heterozygous site).  only_state >= 0: evaluate
// that state alone and return its emission (all lanes); nothing is stored.
AWB_HD inline double awb_emit_site_phase(const AwbChain &ch, int i, int b, int lane,
                                         int nlanes, unsigned char *scratch,
                                         const int *parent, const int *age,
                                         const short *c0, const short *c1,
                                         const short *order, const short *lstart,
                                         long long fwbias, int swap, int only_state,
                                         int vc_hint = -1)
{
    const AwbModel &m = ch.model;
    const int V = ch.nnodes;
    const int T = m.ntimes;
    const int S = ch.nstates[b];
    if (S == 0)
        return 1.0;                     // emit.cpp:665-669
    // row of the alignment that leaf j (or the new leaf, j = nrows - 1) reads
    const int pr1 = swap ? ch.phase_row1 : -1, pr2 = swap ? ch.phase_row2 : -1;
#define AWB_ROW_OF(j) ch.rowidx[(j) == pr1 ? pr2 : ((j) == pr2 ? pr1 : (j))]
    const bool internal = ch.internal != 0;
    const int root = ch.root[b];
    const int maintree_root = internal ? c1[root] : root;
    const int subtree_root = internal ? c0[root] : root;
    const size_t col = (size_t) ch.start_coord + i;
    // (vc_hint: the caller already knows the site's variant column)
    const int vc = ch.seqs ? -1 : (vc_hint >= 0 ? vc_hint : awb_var_find(ch, (long long) col));

    double *inner = (double *) scratch;
    double *outer = inner + 4 * (size_t) V;
    double *mut = outer + 4 * (size_t) V;
    double *nomut = mut + V;
    unsigned char *inmain = (unsigned char *) (nomut + V);

    // branch mutation probabilities (emit.cpp:97-116) and leaf rows (:159-173)
    for (int j = lane; j < V; j += nlanes) {
        double mu_j = 0.0, nomu_j = 0.0;
        if (j != root) {
            const int pa = age[parent[j]];
            if (pa != m.removed_root_time) {
                // prob_branch(max(times[pa] - times[age[j]], mintime)), tabulated
                const double *pt = ch.ptab + ((size_t) age[j] * T + pa) * 2;
                mu_j = pt[0];
                nomu_j = pt[1];
            }
        }
        mut[j] = mu_j;
        nomut[j] = nomu_j;
        inmain[j] = 0;
        if (c0[j] == -1)
            awb_leaf_row(awb_seq_at(ch, AWB_ROW_OF(j), col, vc), inner + 4 * j);
    }
    AWB_LANESYNC();

    // `order` is sorted by height above the leaves and lstart[] gives the level
    // boundaries (awb_block_setup): the nodes of a level only depend on lower
    // levels (inner pass) or higher levels (outer pass), so a level is worked
    // through in parallel, one lane per (node, base).
    const int nlev = lstart[V + 1];

    // inner partials, children before parents (emit.cpp:175-194)
    for (int L = 1; L < nlev; L++) {
        const int q0 = lstart[L], cntL = lstart[L + 1] - q0;
        for (int idx = lane; idx < 4 * cntL; idx += nlanes) {
            const int j = order[q0 + (idx >> 2)];
            const int a = idx & 3;
            const int k1 = c0[j], k2 = c1[j];
            // Jukes-Cantor: sum_x v[x] P(a,x) = mut * sum_x v[x] + (nomut - mut) * v[a]
            const double *v1 = inner + 4 * k1, *v2 = inner + 4 * k2;
            const double p1 = fma(nomut[k1] - mut[k1], v1[a],
                                  mut[k1] * ((v1[0] + v1[1]) + (v1[2] + v1[3])));
            const double p2 = fma(nomut[k2] - mut[k2], v2[a],
                                  mut[k2] * ((v2[0] + v2[1]) + (v2[2] + v2[3])));
            inner[4 * j + a] = p1 * p2;
        }
        AWB_LANESYNC();
    }

    // outer partials, parents before children (emit.cpp:200-296); only nodes
    // of the main tree
    for (int L = nlev - 1; L >= 0; L--) {
        const int q0 = lstart[L], cntL = lstart[L + 1] - q0;
        for (int idx = lane; idx < 4 * cntL; idx += nlanes) {
            const int j = order[q0 + (idx >> 2)];
            const int a = idx & 3;
            bool in;
            if (j == maintree_root) {
                in = true;
                outer[4 * j + a] = 1.0;
            } else {
                const int p = parent[j];
                in = (p != -1) && inmain[p];
                if (in) {
                    const int sib = (c0[p] == j) ? c1[p] : c0[p];
                    const double *v1 = inner + 4 * sib, *v2 = outer + 4 * p;
                    const double p1 = fma(nomut[sib] - mut[sib], v1[a],
                                          mut[sib] * ((v1[0] + v1[1]) + (v1[2] + v1[3])));
                    const double p2 = fma(nomut[p] - mut[p], v2[a],
                                          mut[p] * ((v2[0] + v2[1]) + (v2[2] + v2[3])));
                    outer[4 * j + a] = (p != maintree_root) ? p1 * p2 : p1;
                }
            }
            if (a == 0)
                inmain[j] = in ? 1 : 0;
        }
        AWB_LANESYNC();
    }

    // ---- infinite-sites rule (get_infinite_sites_states, emit.cpp:457-589):
    // unweighted parsimony sets of the site (parsimony_ancestral_set,
    // emit.cpp:898-945), level by level like the partials; a state is valid when
    // the threaded lineage's base set meets the set of its node or of the
    // node's parent.  Invalid states have their emission multiplied by the
    // penalty (emit.cpp:848-862).
    const bool infsites = ch.infsites_penalty < 1.0;
    unsigned char *sets = inmain + V;
    unsigned char *anc = sets + V;
    int cset = 0;
    bool allvalid = true;
    if (infsites) {
        for (int L = 0; L < nlev; L++) {
            const int q0 = lstart[L], cntL = lstart[L + 1] - q0;
            for (int idx = lane; idx < cntL; idx += nlanes) {
                const int j = order[q0 + idx];
                if (c0[j] == -1) {
                    sets[j] = (unsigned char)
                        awb_base_bit(awb_seq_at(ch, AWB_ROW_OF(j), col, vc));
                } else {
                    const int l = sets[c0[j]], r = sets[c1[j]];
                    sets[j] = (unsigned char) ((l & r) ? (l & r) : (l | r));
                }
            }
            AWB_LANESYNC();
        }
        // bases of the leaves on either side
        int mainset = 0, subset = 0;
        for (int j = lane; j < V; j += nlanes) {
            if (c0[j] == -1) {
                if (inmain[j]) mainset |= sets[j]; else subset |= sets[j];
            }
        }
        mainset = awb_lanes_or(mainset);
        subset = awb_lanes_or(subset);
        if (internal) {
            cset = sets[subtree_root];
        } else {
            cset = awb_base_bit(awb_seq_at(ch, AWB_ROW_OF(ch.nrows - 1), col, vc));
            subset = cset;
        }
        // distinct bases on the two sides: the lineage can go anywhere
        allvalid = !(subset & mainset);
        if (!allvalid) {
            for (int L = nlev - 1; L >= 0; L--) {
                const int q0 = lstart[L], cntL = lstart[L + 1] - q0;
                for (int idx = lane; idx < cntL; idx += nlanes) {
                    const int j = order[q0 + idx];
                    if (!inmain[j])
                        continue;
                    if (j == maintree_root) {
                        anc[j] = sets[j];
                    } else {
                        const int ps = anc[parent[j]] & sets[j];
                        anc[j] = (unsigned char) (ps ? ps : sets[j]);
                    }
                }
                AWB_LANESYNC();
            }
        }
    }

    // the branch being threaded: new leaf (external) or the subtree root
    double in2[4];
    if (internal) {
        for (int x = 0; x < 4; x++)
            in2[x] = inner[4 * subtree_root + x];
    } else {
        awb_leaf_row(awb_seq_at(ch, AWB_ROW_OF(ch.nrows - 1), col, vc), in2);
    }
    // branch of the threaded lineage below the coalescence point: from the
    // subtree root's time (internal) or from time 0.0 (external) up to the
    // state's time; tabulated like every other branch probability
    const double *tab0 = internal ?
        ch.ptab + (size_t) age[subtree_root] * T * 2 : ch.ptab + (size_t) T * T * 2;

    const double in2sum = (in2[0] + in2[1]) + (in2[2] + in2[3]);

    // per-state emission (emit.cpp:778-805, calc_emit :620-645)
    const long long row0 = ch.row_off[b];
    double *out = ch.fw + (ch.fw_off[b] - fwbias) +
        (long long) (i - ch.block_start[b]) * S;
    double ret = 0.0;
    const int kbeg = only_state >= 0 ? only_state : lane;
    const int kend = only_state >= 0 ? only_state + 1 : S;
    for (int k = kbeg; k < kend; k += nlanes) {
        const int node2 = ch.st_node[row0 + k];
        const int p = parent[node2];
        const int bt = ch.st_time[row0 + k];
        // d0 = coal - time1, d1 = coal - times[age[node2]], d2 = parent - coal
        // (each floored at mintime; the root branch's d2 is not used)
        const double *t0 = tab0 + (size_t) bt * 2;
        const double *t1 = ch.ptab + ((size_t) age[node2] * T + bt) * 2;
        const double *t2 = ch.ptab +
            ((size_t) bt * T + (p != -1 ? awb_imin(age[p], T - 1) : bt)) * 2;
        const double mu0 = t0[0], no0 = t0[1];
        const double mu1 = t1[0], no1 = t1[1];
        const double mu2 = t2[0], no2 = t2[1];
        const double *in_n = inner + 4 * node2;
        const double *out_n = outer + 4 * node2;
        // (Jukes-Cantor as above: one sum per vector, one FMA per base)
        const double s1 = mu0 * in2sum;
        const double s2 = mu1 * ((in_n[0] + in_n[1]) + (in_n[2] + in_n[3]));
        const double s3 = mu2 * ((out_n[0] + out_n[1]) + (out_n[2] + out_n[3]));
        const double d0 = no0 - mu0, d1 = no1 - mu1, d2 = no2 - mu2;
        const bool rootbr = node2 == maintree_root;
        double emit = 0.0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const double p1 = fma(d0, in2[a], s1);
            const double p2 = fma(d1, in_n[a], s2);
            const double p3 = rootbr ? 1.0 : fma(d2, out_n[a], s3);
            emit += p1 * p2 * p3 * .25;
        }
        if (infsites && !allvalid) {
            const bool valid = (cset & anc[node2]) ||
                (node2 != maintree_root && (cset & anc[p]));
            if (!valid)
                emit *= ch.infsites_penalty;
        }
        if (only_state >= 0)
            ret = emit;
        else if (swap)
            out[k] = (out[k] + emit) * 0.5;     // emit.cpp:840-841
        else
            out[k] = emit;
    }
#undef AWB_ROW_OF
    return ret;
}

// is site i heterozygous in the unphased individual (emit.cpp:716-717)?
AWB_HD inline bool awb_site_het(const AwbChain &ch, int i)
{
    if (ch.phase_row1 < 0 || ch.phase_row2 < 0)
        return false;
    const size_t col = (size_t) ch.start_coord + i;
    const int vc = ch.seqs ? -1 : awb_var_find(ch, (long long) col);
    return awb_seq_at(ch, ch.rowidx[ch.phase_row1], col, vc) !=
        awb_seq_at(ch, ch.rowidx[ch.phase_row2], col, vc);
}

AWB_HD inline void awb_emit_site(const AwbChain &ch, int i, int b, int lane,
                                 int nlanes, unsigned char *scratch,
                                 const int *parent, const int *age,
                                 const short *c0, const short *c1,
                                 const short *order, const short *lstart,
                                 long long fwbias = 0, int vc_hint = -1)
{
    awb_emit_site_phase(ch, i, b, lane, nlanes, scratch, parent, age, c0, c1, order,
                        lstart, fwbias, 0, -1, vc_hint);
    if (awb_site_het(ch, i)) {
        // the other phasing; the row becomes the mean of the two
        AWB_LANESYNC();
        awb_emit_site_phase(ch, i, b, lane, nlanes, scratch, parent, age, c0, c1, order,
                            lstart, fwbias, 1, -1, vc_hint);
    }
}

// P(the data's phasing | state) at a heterozygous site: e1 / (e1 + e2), what
// calc_emissions hands to PhaseProbs::add (emit.cpp:838-839)
AWB_HD inline double awb_phase_prob(const AwbChain &ch, int i, int b, int state, int lane,
                                    int nlanes, unsigned char *scratch,
                                    const int *parent, const int *age,
                                    const short *c0, const short *c1,
                                    const short *order, const short *lstart)
{
    const double e1 = awb_emit_site_phase(ch, i, b, lane, nlanes, scratch, parent, age,
                                          c0, c1, order, lstart, 0, 0, state);
    AWB_LANESYNC();
    const double e2 = awb_emit_site_phase(ch, i, b, lane, nlanes, scratch, parent, age,
                                          c0, c1, order, lstart, 0, 1, state);
    return e1 / (e1 + e2);
}

#endif // AWB_EMIT_CUH
