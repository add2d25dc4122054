// awb_common.cuh -- shared device/host structures of the B200 threading-HMM path.
//
// Data layout in HBM (one AwbChain per independent problem, all arrays flat):
//
//   per block b (local tree), B = ntrees:
//     nstates[b]                true state count S_b (0 allowed in internal mode)
//     row_off[b]                offset of the block's per-state rows; a row has
//                               max(S_b,1) entries ("zero-state blocks are size 1",
//                               reference sample_thread.cpp:349-351)
//     fw_off[b]                 offset (doubles) of the block's slab of the
//                               forward table: fw[fw_off[b] + (i-start_b)*S1 + j]
//                               with j in the reference's state order
//   per state (node-major = reference order, index row_off[b]+j):
//     st_node, st_time, inv_emit, band_j1/band_len/band_boff, sw_start/sw_cnt
//   per state in time-major order (index row_off[b]+q): perm (-> j), pslot
//
// The forward kernel runs one CTA per chain with one thread per state in
// TIME-MAJOR order, so that the per-time group sums of the SMC transition are
// contiguous lane segments (warp-shuffle segmented reduction).
#ifndef AWB_COMMON_CUH
#define AWB_COMMON_CUH

#include <string.h>

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define AWB_HD __host__ __device__
#else
#define AWB_HD
#endif

#define AWB_MAXT 64      // == AWB_MAX_NTIMES
#define AWB_MAXV 1024    // == AWB_MAX_NNODES
#define AWB_MAXS 2048    // == AWB_MAX_NSTATES
#ifndef AWB_NSCRIBE
#define AWB_NSCRIBE 64    // F-scribe lanes of the fast forward kernel
#endif

enum { AWB_TM_D = 0, AWB_TM_E, AWB_TM_LNB, AWB_TM_LNE2, AWB_TM_LNNEGG1,
       AWB_TM_G2, AWB_TM_G3, AWB_TM_LNG4, AWB_TM_NORECOMBS, AWB_TM_NVEC };

enum { AWB_SITE_INVARIANT = 0, AWB_SITE_VARIANT = 1, AWB_SITE_MASKED = 2 };

// ArgModel time grid (reference model.h:41-354)
struct AwbModel {
    int ntimes;
    int removed_root_time;             // ntimes + 1 (model.h:207)
    double rho, mu, mintime;           // mintime = 0.1 * times[1] (model.h:211)
    double times[AWB_MAXT];
    double time_steps[AWB_MAXT];       // last = inf (model.h:322-333)
    double popsizes[AWB_MAXT];
    double coal_time_steps[2 * AWB_MAXT]; // model.cpp:9-23
};

struct AwbChain {
    AwbModel model;
    int internal, minage;
    int nleaves, nrows, nseqs, seqlen;
    int ntrees, nnodes, nsites, start_coord;
    int maxS;                 // max over blocks of max(S_b, 1)
    int maxband;              // max over blocks of the band size (doubles)
    int maxNS;                // max over blocks of the padded thread count
    int zcap;                 // max over blocks of the scribes' padded column (awb_scribe_plan)
    int maxcnt;               // longest branch (states)
    int keep_debug;
    int need_band;            // compute the tmatrix2 band (generic forward kernel only)
    int gen_mappings;         // no node mappings given: identity except the broken node
    // checkpointed table: `fw` / `fsum` hold ONE segment (blocks
    // [seg_start[s], seg_start[s+1]) plus the first row of the next block);
    // ckptcol[s] is the stored first column of segment s (s >= 1)
    int ckpt, nseg;
    int phase_row1, phase_row2;   // unphased individual: its two rows (leaf order; -1: off)
    int nrot;                 // the second pass rebuilds this many segments side by side
    int nslots;               // segment tables in fw / fsum (the last nslots segments
                              //   of the forward pass stay resident for the traceback)
    int seg_sites;            // sites per segment table (fsum stride between tables)
    long long seg_doubles;    // doubles per segment table
    const int *seg_start;     // [nseg + 1]
    double *ckptcol;          // [nseg + 1][maxS]

    // ---- inputs
    const int *ptrees;        // [B][V]
    const int *ages;          // [B][V]
    const int *sprs;          // [B][4]
    int *mappings;            // [B][V] old node -> new node (-1: broken), caller's or K1's
    const double *ptab;       // [(T*T + T)][2] branch probabilities by time-index pair
    const int *blocklens;     // [B]
    const int *subtree_roots; // [B] (internal) or NULL
    const int *rowidx;        // [nrows] rows of seqs compared for invariance:
                              //   leaves' seqids, then new_chrom (external)
    const unsigned char *seqs; // [nseqs][seqlen], or NULL when the alignment is given
                              //   as its variant columns only:
    int nvar;                 // variant columns
    const int *var_pos;       // [nvar] ascending column indices of the dense rows
    const unsigned char *var_cols; // [nvar][nseqs]
    int default_char;         // every other column of every row
    double infsites_penalty;  // < 1: infinite-sites penalty on (0 / 1: off)

    // ---- layout (host computed)
    const int *block_start;   // [B+1] site offset of each block (0-based in chain)
    const int *nstates;       // [B]
    const long long *row_off; // [B+1]
    const long long *fw_off;  // [B+1]
    const long long *band_off;// [B+1]
    const long long *ent_off; // [B+1]
    const long long *sw1_off; // [B+1] (debug arrays of the switch matrix)
    const long long *trow_off;// [B+1] offsets of the forward kernel's thread map

    // ---- block setup outputs (K1)
    short *st_node;           // [rows]
    signed char *st_time;     // [rows]
    unsigned short *perm;     // [rows] time-major q -> node-major j
    unsigned short *pslot;    // [rows] partial-sum slot of time-major q
    unsigned short *band_j1;  // [rows] first state of this state's branch
    unsigned char *band_len;  // [rows] number of states on the branch
    int *band_boff;           // [rows] offset of this state's coefficients in the block band
    double *inv_emit;         // [rows] invariant-site emission
    double *band;             // [band_off[B]] same-branch coefficients (tmatrix2)
    double *tmatrix;          // [B][T*T]  tmatrix[a*T+b]
    double *tmvec;            // [B][9][T]
    unsigned short *rowstart; // [B][T+1] time-major start of each time row
    unsigned short *pstart;   // [B][T+1] first partial slot of each time row
    unsigned char *slotrow;   // [B][slotcap] time row of each partial slot
    unsigned short *tmap;     // [trow_off] forward thread -> state (0xFFFF idle);
                              //   node-major, a branch never straddles a warp
    unsigned short *iperm;    // [rows] state -> time-major slot (inverse of perm)
    signed char *st_age;      // [rows] age of the state's node
    double *lin;              // [B][7][T] linear-domain transition vectors:
                              //   D, h=B-NegG1, B, E*E2, E*(pre+G3), E*(pre+G2), norecombs
    unsigned short *sc_start; // [B][AWB_NSCRIBE] scribe lane -> first time-major slot
    unsigned short *sc_cnt;   // [B][AWB_NSCRIBE] scribe lane -> number of slots
    unsigned char *sc_row;    // [B][AWB_NSCRIBE] scribe lane -> time row (255 idle)
    unsigned char *sc_stride; // [B][AWB_NSCRIBE] lanes sharing the row: lane i sums slots i, i+stride, ...
    unsigned short *sc_ch;    // [B] slots every scribe lane sums (rows are zero-padded)
    int slotcap;              // ntimes + 32
    short *node_first;        // [B][V] first state index of node (or -1)
    short *node_cnt;          // [B][V]
    short *child0, *child1;   // [B][V]
    short *order;             // [B][V] post-order (local_tree.h:274-304)
    short *lstart;            // [B][V+2] order[] is sorted by height above the leaves:
                              //   level L = order[lstart[L] .. lstart[L+1]); [V+1] = #levels
    short *root;              // [B]
    int *lineages;            // [B][3][T] nbranches, nrecombs, ncoals
    double *treelen;          // [B] get_treelen / get_treelen_internal
    int *tm_minage;           // [B]

    // ---- switch setup outputs (K2), CSR over targets of block b
    unsigned short *sw_start; // [rows]
    unsigned short *sw_cnt;   // [rows]
    unsigned short *sw_src;   // [ent_off[B]] source state (previous block order)
    double *sw_prob;          // [ent_off[B]]
    // debug copies in the reference's representation (keep_debug)
    int *sw_determ;           // [sw1_off[B]]
    double *sw_determprob;    // [sw1_off[B]]
    double *sw_recombrow;     // [rows]
    double *sw_recoalrow;     // [rows]
    int *sw_recombsrc;        // [B]
    int *sw_recoalsrc;        // [B]

    // ---- sites, tables, results
    unsigned char *kind;      // [nsites]
    double *fw;               // [fw_off[B]]
    int *path;                // [nsites]
    const int *rand_ints;     // [nsites]
    double *logz;             // [1]
    double *sink;             // [1024] per-thread dump slot of the forward kernel
    double *fsum;             // [n][T-1] per-time sums of the stored forward columns
    int *status;              // [1] first bad site or -1
    int last_state;           // traceback: -1 = sample last column
};

// ------------------------------------------------------------------ sequences

// index of dense column `col` among the variant columns, or -1
AWB_HD inline int awb_var_find(const AwbChain &ch, long long col)
{
    int lo = 0, hi = ch.nvar - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int v = ch.var_pos[mid];
        if (v == col) return mid;
        if (v < col) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

// character of sequence `row` at dense column `col`; vc = awb_var_find(col) when
// the alignment is packed (ch.seqs == NULL)
AWB_HD inline unsigned char awb_seq_at(const AwbChain &ch, int row, size_t col, int vc)
{
    if (ch.seqs)
        return ch.seqs[(size_t) row * ch.seqlen + col];
    return vc >= 0 ? ch.var_cols[(size_t) vc * ch.nseqs + row] :
        (unsigned char) ch.default_char;
}

// ------------------------------------------------------------------ math

// common.h:157-163
AWB_HD inline double awb_logadd(double lna, double lnb)
{
    if (lna == -INFINITY) return lnb;
    if (lnb == -INFINITY) return lna;
    return fmax(lna, lnb) + log1p(exp(-fabs(lna - lnb)));
}

// The part of a chain one kernel launch works on.  Whole table: everything.
// Checkpointed table (ch.ckpt): segment s = blocks [b0, b1) plus, when there is
// a next block, its first site ("extra": it yields the first column of segment
// s+1 and lets the traceback step through the breakpoint between the segments).
// fw / fsum then hold that segment only; fwbias / site0 turn whole-table
// offsets into offsets of the segment's table.
struct AwbSeg {
    int b0, b1, extra, site0, nsites;
    long long fwbias;       // fw table of the segment: ch.fw - fwbias + (whole-table offset)
    long long fsoff;        // fsum of the segment: ch.fsum + fsoff + (site - site0) * (T-1)
    bool resident;          // the table written by the forward pass is still there
    bool valid;
};

AWB_HD inline AwbSeg awb_seg(const AwbChain &ch, int s)
{
    AwbSeg g;
    if (!ch.ckpt) {
        g.b0 = 0; g.b1 = ch.ntrees; g.extra = 0; g.site0 = 0; g.nsites = ch.nsites;
        g.fwbias = 0; g.fsoff = 0; g.resident = true; g.valid = true;
        return g;
    }
    g.valid = s >= 0 && s < ch.nseg;
    if (!g.valid) {
        g.b0 = g.b1 = 0; g.extra = 0; g.site0 = 0; g.nsites = 0; g.fwbias = 0;
        g.fsoff = 0; g.resident = false;
        return g;
    }
    g.b0 = ch.seg_start[s];
    g.b1 = ch.seg_start[s + 1];
    g.extra = g.b1 < ch.ntrees ? 1 : 0;
    g.site0 = ch.block_start[g.b0];
    g.nsites = ch.block_start[g.b1] - g.site0 + g.extra;
    // the last nslots segments have a table of their own (and stay resident
    // after the forward pass); all earlier ones take turns in the top nrot tables
    // (segment s in table R-1 - s mod nrot; table 0 when there are fewer than
    // three): those are the tables of the last segments, which the traceback has
    // left behind by the time it rebuilds an earlier segment -- so nrot
    // consecutive segments can be rebuilt side by side (awb_api.cu never lets
    // such a group straddle a window's resident boundary, above which its tables
    // are still in use)
    const int R = ch.nslots < ch.nseg ? ch.nslots : ch.nseg;
    g.resident = s >= ch.nseg - R;
    const int M = ch.nrot > 1 ? ch.nrot : 1;
    const int slot = g.resident ? s - (ch.nseg - R) : (R >= M ? R - 1 - (s % M) : 0);
    g.fwbias = ch.fw_off[g.b0] - (long long) slot * ch.seg_doubles;
    g.fsoff = (long long) slot * ch.seg_sites * (ch.model.ntimes - 1);
    return g;
}

// trans.h:89-128 TransMatrix::get_time.  v points at 9 vectors of stride T.
AWB_HD inline double awb_get_time(const double *v, int T, int a, int b, int c,
                                  int minage, bool same_node)
{
    if (a < minage || b < minage)
        return 0.0;
    const double *D = v + AWB_TM_D * T, *E = v + AWB_TM_E * T,
        *lnB = v + AWB_TM_LNB * T, *lnE2 = v + AWB_TM_LNE2 * T,
        *lnNegG1 = v + AWB_TM_LNNEGG1 * T, *G2 = v + AWB_TM_G2 * T,
        *G3 = v + AWB_TM_G3 * T, *lnG4 = v + AWB_TM_LNG4 * T,
        *norecombs = v + AWB_TM_NORECOMBS * T;
    const double term1 = D[a] * E[b];
    double minage_term = 0.0;
    if (minage > 0)
        minage_term = exp(lnG4[b] + lnB[minage - 1]);

    if (!same_node) {
        if (a < b)
            return term1 * (exp(lnE2[b] + lnB[a]) - exp(lnE2[b] + lnNegG1[a])
                            - minage_term);
        else if (a == b)
            return term1 * ((b > 0 ? exp(lnE2[b] + lnB[b - 1]) : 0.0) + G3[b]
                            - minage_term);
        else
            return term1 * ((b > 0 ? exp(lnE2[b] + lnB[b - 1]) : 0.0) + G2[b]
                            - minage_term);
    } else {
        const double c_term = (c > 0 ? exp(lnG4[b] + lnB[c - 1]) : 0.0);
        if (a < b)
            return term1 * (2 * (exp(lnE2[b] + lnB[a]) -
                                 exp(lnE2[b] + lnNegG1[a]))
                            - c_term - minage_term);
        else if (a == b)
            return term1 * ((2 * ((b > 0 ? exp(lnE2[b] + lnB[b - 1]) : 0.0) +
                                  G3[b])) - c_term - minage_term)
                + norecombs[a];
        else
            return term1 * ((2 * ((b > 0 ? exp(lnE2[b] + lnB[b - 1]) : 0.0) +
                                  G2[b])) - c_term - minage_term);
    }
}

// emit.cpp:86-93 (Jukes-Cantor)
AWB_HD inline double awb_prob_branch(double t, double mu, bool mut)
{
    const double f = 4. / 3.;
    if (!mut)
        return .25 * (1.0 + 3. * exp(-f * mu * t));
    else
        return .25 * (1.0 - exp(-f * mu * t));
}

AWB_HD inline int awb_imax(int a, int b) { return a > b ? a : b; }
AWB_HD inline int awb_imin(int a, int b) { return a < b ? a : b; }

AWB_HD inline void awb_store2(double *p, double a, double b)
{
#ifdef __CUDA_ARCH__
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
#else
    p[0] = a;
    p[1] = b;
#endif
}

// two doubles with one 16-byte store where the address allows it
AWB_HD inline void awb_store_pair(double *p, double a, double b)
{
    if ((((size_t) p) & 15) == 0) {
        awb_store2(p, a, b);
    } else {
        p[0] = a;
        p[1] = b;
    }
}

// two ints from an 8-byte aligned address
AWB_HD inline void awb_load2i(const int *p, int &a, int &b)
{
#ifdef __CUDA_ARCH__
    const int2 v = *reinterpret_cast<const int2 *>(p);
    a = v.x;
    b = v.y;
#else
    a = p[0];
    b = p[1];
#endif
}

// Consecutive 1-, 2- or 4-byte values of a global array, stored 8 bytes at a
// time: a thread per block writes its state tables element by
// element, every store of a warp touches 32 different sectors, and the number
// of stores is what K1's time is made of.  The first and last group of a block
// are written element-wise (they share their 8 bytes with the neighbour blocks).
template <class E>
struct AwbPacker {
    static const int N = 8 / (int) sizeof(E);     // elements per 8-byte store
    static const int BITS = 8 * (int) sizeof(E);
    E *base;
    long long pos;
    unsigned long long acc;
    int lo;             // first slot of the current group that is this block's
    AWB_HD void init(E *arr, long long start)
    {
        base = arr;
        pos = start;
        acc = 0;
        lo = (int) (start & (N - 1));
    }
    // (not through a cast pointer: the arrays are read back as E right away)
    AWB_HD void store8(E *at)
    {
#ifdef __CUDA_ARCH__
        asm volatile("st.u64 [%0], %1;" :: "l"(at), "l"(acc) : "memory");
#else
        memcpy(at, &acc, 8);
#endif
    }
    AWB_HD void part(long long g0, int from, int to)
    {
        for (int q = from; q < to; q++)
            base[g0 + q] = (E) (acc >> (BITS * q));
    }
    AWB_HD void put(unsigned v)
    {
        const int sl = (int) (pos & (N - 1));
        const unsigned long long mask = (1ull << BITS) - 1ull;
        acc |= ((unsigned long long) v & mask) << (BITS * sl);
        pos++;
        if (sl == N - 1) {
            if (lo)
                part(pos - N, lo, N);
            else
                store8(base + pos - N);
            acc = 0;
            lo = 0;
        }
    }
    // the unfinished last group
    AWB_HD void flush()
    {
        const int end = (int) (pos & (N - 1));
        if (end)
            part(pos - end, lo, end);
        acc = 0;
        lo = end;
    }
};

// Thread packing of the fast forward kernel: every branch (cnt[i] consecutive
// states of node i) gets consecutive lanes of ONE warp.  First-fit in
// decreasing branch length, so the number of warps is close to S/32.  Used by
// the host layout (to size the thread map) and by K1 (to fill it): both must
// run the same code.  A branch of 33..64 states takes the slots of two whole
// warps, starting at an even one: the forward kernel then keeps it in two
// register sets of one warp (awb_forward_fast.cuh).
//
// The placement only builds, per warp of slots ("row"), the list of its
// branches in slot order; awb_pack_emit then writes the whole map front to
// back through a packer (8 bytes a store; K1 is a thread per block and the
// number of its stores is its time).
#define AWB_PACK_ROWS (AWB_MAXS / 16 + 2)     // > 2048/17 one-per-row branches, > 2*2048/33 long ones

template <int VCAP>
struct AwbRowLists {
    unsigned short rhead[AWB_PACK_ROWS], rtail[AWB_PACK_ROWS], rnext[VCAP];
    int nw;                                   // rows in use
    AWB_HD void open_row()
    {
        rhead[nw] = 0xFFFF;
        rtail[nw] = 0xFFFF;
        nw++;
    }
    AWB_HD void append(int w, int i)
    {
        rnext[i] = 0xFFFF;
        if (rhead[w] == 0xFFFF)
            rhead[w] = (unsigned short) i;
        else
            rnext[rtail[w]] = (unsigned short) i;
        rtail[w] = (unsigned short) i;
    }
};

// First-fit decreasing.  Returns the number of thread slots (a multiple of 32;
// more than any map can hold when the rows run out).  rl may be NULL (count
// only).
template <int VCAP>
AWB_HD inline int awb_pack_place(const short *cnt, int V, AwbRowLists<VCAP> *rl)
{
    unsigned char fill[AWB_PACK_ROWS];
    int nw = 0;
    if (rl)
        rl->nw = 0;
    // the branches of each length, in node order, linked through nxt
    unsigned short head[64], nxt[VCAP];
    for (int l = 0; l < 64; l++)
        head[l] = 0xFFFF;
    for (int i = V - 1; i >= 0; i--) {
        const int c = cnt[i];
        if (c > 0) {
            const int l = c > 63 ? 63 : c - 1;
            nxt[i] = head[l];
            head[l] = (unsigned short) i;
        }
    }
    for (int len = 63; len >= 0; len--) {
        // first fit: a warp that had no room for a branch of this length has
        // none for the next one of the same length either
        int w = 0;
        for (int i = head[len]; i != 0xFFFF; i = nxt[i]) {
            const int c = len == 63 ? cnt[i] : len + 1;
            if (nw + 3 > AWB_PACK_ROWS)
                return 32 * 4 * AWB_PACK_ROWS;
            if (c > 32) {
                // 33..64 states: two whole warps' worth of slots starting at an
                // even one (the long branches come first, two each)
                if (nw & 1) {
                    fill[nw++] = 32;
                    if (rl) rl->open_row();
                }
                fill[nw++] = 32;
                fill[nw++] = 32;
                if (rl) {
                    rl->open_row();
                    rl->open_row();
                    rl->append(nw - 2, i);
                }
            } else {
                while (w < nw && fill[w] + c > 32)
                    w++;
                if (w == nw) {
                    fill[nw++] = 0;
                    if (rl) rl->open_row();
                }
                fill[w] = (unsigned char) (fill[w] + c);
                if (rl) rl->append(w, i);
            }
        }
    }
    return 32 * (nw > 0 ? nw : 1);
}

// Node order (what the host reserves when it does not run the packing): a
// branch that would straddle a warp starts the next one, a long branch takes
// two whole warps from an even one.
template <int VCAP>
AWB_HD inline int awb_pack_place_in_order(const short *cnt, int V, AwbRowLists<VCAP> *rl)
{
    int tpos = 0;
    rl->nw = 0;
    for (int i = 0; i < V; i++) {
        const int c = cnt[i];
        if (c <= 0) continue;
        if (c > 32)
            tpos = (tpos + 63) & ~63;
        else if ((tpos & 31) + c > 32)
            tpos = (tpos + 31) & ~31;
        const int w = tpos >> 5;
        if (w + 3 > AWB_PACK_ROWS)
            return 32 * 4 * AWB_PACK_ROWS;
        while (rl->nw <= w + (c > 32 ? 1 : 0))
            rl->open_row();
        rl->append(w, i);
        tpos += c;
        if (c > 32)
            tpos = (tpos + 63) & ~63;
    }
    return tpos;
}

// the map of `cap` slots from the row lists: tmap[start + slot] = state, 0xFFFF
// for a slot without one
template <int VCAP>
AWB_HD inline void awb_pack_emit(const AwbRowLists<VCAP> &rl, const short *cnt,
                                 const short *nfirst, unsigned short *tmap, long long start,
                                 int cap)
{
    AwbPacker<unsigned short> pk;
    pk.init(tmap, start);
    int done = 0;
    for (int w = 0; w < rl.nw && done < cap; w++) {
        int end = 32 * (w + 1);
        for (int i = rl.rhead[w]; i != 0xFFFF; i = rl.rnext[i]) {
            const int c = cnt[i], nf = nfirst[i];
            if (c > 32)
                end += 32;                      // this row and the next
            for (int t = 0; t < c && done < cap; t++, done++)
                pk.put((unsigned) (nf + t));
        }
        if (end > 32 * (w + 1))
            w++;
        for (; done < end && done < cap; done++)
            pk.put(0xFFFFu);
    }
    for (; done < cap; done++)
        pk.put(0xFFFFu);
    pk.flush();
}

// count only (host layout)
AWB_HD inline int awb_pack_branches(const short *cnt, int V)
{
    return awb_pack_place<AWB_MAXV>(cnt, V, (AwbRowLists<AWB_MAXV> *) 0);
}

// The F-scribes' plan of one block (awb_setup.cuh K1 fills it in,
// awb_forward_fast.cuh reads it): the AWB_NSCRIBE lanes are shared out over the
// time rows, nl = ceil(w / CH) lanes for a row of w states (at least one, a row
// never straddles a warp), with CH the smallest even slot count per lane for
// which all rows fit.  Every row is padded to nl*CH slots.  Returns the padded
// size of the column; the host layout runs the same code to size the kernel's
// buffer.  w[0..nrows) are the row widths.
AWB_HD inline int awb_scribe_plan(const int *w, int nrows, int &CHout)
{
    int CH = 2;
    for (;; CH += 2) {
        int l = 0;
        bool fits = true;
        for (int t = 0; t < nrows && fits; t++) {
            const int nl = w[t] == 0 ? 1 : (w[t] + CH - 1) / CH;
            if (nl > 32) { fits = false; break; }
            if ((l & 31) + nl > 32) l = (l + 31) & ~31;
            l += nl;
            if (l > AWB_NSCRIBE) fits = false;
        }
        if (fits) break;
    }
    int zb = 0;
    for (int t = 0; t < nrows; t++)
        zb += (w[t] == 0 ? 1 : (w[t] + CH - 1) / CH) * CH;
    CHout = CH;
    return zb;
}

#endif // AWB_COMMON_CUH
