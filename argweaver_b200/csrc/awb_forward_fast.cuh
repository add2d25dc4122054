// awb_forward_fast.cuh -- K4, the forward recursion of the threading HMM
// (latency-optimised; awb_forward.cuh keeps a generic fallback).
//
// Replaces arghmm_forward_alg / arghmm_forward_block / arghmm_forward_switch
// (reference sample_thread.cpp:394-460, :186-296, :345-389).
//
// Mathematics.  Inside a block the SMC transition is rank-(T-1) plus a
// per-branch correction:
//
//   col2[k] = emit[k] * ( R[b_k] + W_k )
//   R[b]  = sum_a tmatrix[a][b] * F[a],    F[a] = sum_{j: time_j = a} col[j]
//   W_k   = sum_{a on branch(k)} tmatrix2[a][k] * col[(node_k, a)]
//
// and tmatrix2 (reference sample_thread.cpp:218-225) separates in (a, b), so
// along the branch n of state k = (n, b), with x_a = D[a]*col[(n,a)]:
//
//   W_k = A1 * PY + A2' * col[k] + A3 * Q
//   PY = sum_{a<b} x_a (h[a] - Bc)   Q = sum_{a>b} x_a
//
// (Bc, A1, A2' = D[b]*A2 + norecombs[b], A3 per-state constants, see
// load_compute).  PY and Q are exclusive prefix / suffix sums along one branch;
// the separable form agrees with the literal band to ~4e-15 relative.
//
// Mapping to the SM (numbers measured on B200, scripts/microbench.cu:
// dependent DFMA 8.4 cycles, 64-bit SHFL+DADD 35, LDS 33, bar.sync 33,
// DDIV 130).  A site step is a chain of ~30 dependent shared-memory / shuffle /
// FP64 operations, not a bandwidth problem, so the kernel is organised around
// that chain:
//
//   compute warps   U states per thread (U = 1, 2 or 4 register sets) in
//                   NODE-MAJOR order, packed (first-fit decreasing, K1) so that
//                   the states of a branch are consecutive lanes of ONE set of
//                   one warp (a branch of 33..64 states -- possible with more
//                   than 33 time points -- takes set u and the first lanes of
//                   set u+1 of the same warp): PY and Q are segmented
//                   warp-shuffle scans in registers.  The segment bounds ride
//                   in the shuffle's own clamp operand (shfl.sync.up/down with
//                   c = first / last lane of the branch) and its predicate
//                   output gates the add: no masks, no multipliers.
//   F-scribes       two warps owning no state: between barrier 1 and barrier 2
//                   their 64 lanes turn the column (staged in shared memory in
//                   TIME-MAJOR order) into the per-time sums F[a] and then,
//                   one lane per b, into R[b]; a compute thread reads a single
//                   value.
//                   The second scribe warp (idle while the first forms R when
//                   there are at most 32 time rows) also keeps the books:
//                   column norm, logZ, the lagged rescale factor, the per-time
//                   sums of the stored column (fsum) for the traceback, and
//                   the L1 prefetches.  A CTA is then 4 + 2 warps for the bench
//                   shape: 18 warps of three CTAs spread evenly over the four
//                   register files of an SM (7-warp CTAs do not).
//
//   step(site):  STS value (time-major slot)
//                B1 (compute + F-scribes)
//                    compute:  fetch next emission ; branch scans -> W
//                    F-scribes: F[a] ; B3 (scribes) ; R[b]
//                B2 (everybody)
//                    compute:  col(site+1) = (R[b] + W) * emission -> table
//                    norm warp: norm(site) ...
//
// The forward table is written once, 8 B per site*state, in the reference's
// state order, AS CARRIED: a column is stored with the scale it has in the
// recursion (a factor published for column s is applied when column s+3 is
// formed, every AWB_FWD_RS = 4 sites, which bounds the magnitude by the product
// of at most RS+2 one-step norms), not renormalised to sum 1.  Everything
// downstream is scale-free per row: the traceback compares partial sums of
// fw[i][j]*T[j->k] with a fraction of their total, and fsum holds the per-time
// sums of the row as stored.  awb_batch_get_fw normalises the copy it returns
// (rows sum to 1, as the reference stores them).
//
// A launch works on one AwbSeg: the whole window, or one segment of a
// checkpointed table (awb_common.cuh, awb_api.cu).
#ifndef AWB_FORWARD_FAST_CUH
#define AWB_FORWARD_FAST_CUH

#include "awb_common.cuh"

#define AWB_FWD_RS 4          // rescale period (sites); power of two, >= 4
#define AWB_FWD_FSCRIBES AWB_NSCRIBE          // F-scribe lanes
#define AWB_FWD_HELPERS AWB_NSCRIBE           // the two F-scribe warps

// shared memory (doubles): Fs[2][TMAX+2] | Rs[2][TMAX+2] | scaleS[2] | invL[2] |
// dummy[4] | tmS[TMAX/2][NSCRIBE][2] | colS[2][NS] | cst[2][NS][2] | zT[zcap]
// (NS = state slots)
__host__ __device__ inline size_t awb_fwd_fast_smem_bytes(int NS, int TMAX, int zcap)
{
    return (6 * (size_t) NS + (size_t) zcap + 4 * (size_t) (TMAX + 2) + 8 +
            (size_t) (TMAX + 1) / 2 * AWB_NSCRIBE * 2) * sizeof(double);
}

__device__ __forceinline__ void awb_sts(unsigned addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" :: "r"(addr), "d"(v) : "memory");
}

__device__ __forceinline__ double awb_lds(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

// prefetch [p, p + bytes) into L1, one 128-byte line per lane and step
__device__ __forceinline__ void awb_prefetch_range(const void *p, long long bytes, int lane)
{
    // (from the line that holds p[0] to the line that holds the last byte)
    const unsigned long long a = (unsigned long long) p;
    const char *q = (const char *) (a & ~127ull);
    bytes += (long long) (a & 127ull);
    for (long long o = 128ll * lane; o < bytes; o += 128ll * 32)
        asm volatile("prefetch.global.L1 [%0];" :: "l"(q + o));
}

// First and last lane of the run of equal keys this lane is in (a branch's
// states, a row's scribe lanes: always consecutive lanes, and no key comes
// twice in a warp -- what __match_any_sync would give, in a handful of
// instructions instead of its per-value loop).
__device__ __forceinline__ void awb_lane_run(int key, int lane, int &first, int &last)
{
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned starts = __ballot_sync(0xffffffffu, lane == 0 || prev != key);
    first = 31 - __clz(starts & (0xffffffffu >> (31 - lane)));
    const unsigned above = (lane == 31) ? 0u : (starts & (0xffffffffu << (lane + 1)));
    last = above ? (__ffs(above) - 2) : 31;
}

// Segmented scan steps.  shfl.sync.up with clamp operand c = first lane of my
// segment gives p = (lane - delta >= c); shfl.sync.down with c = last lane
// gives p = (lane + delta <= c) (PTX ISA, shfl.sync: segmask bits 0).  The
// predicate gates the add, so a step is two SHFL.32 and one predicated DADD.
__device__ __forceinline__ double awb_scan_up(double v, int delta, int cfirst)
{
    double r;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi;\n\t.reg .f64 t;\n\t"
                 "mov.b64 {lo, hi}, %1;\n\t"
                 "shfl.sync.up.b32 lo|p, lo, %2, %3, 0xffffffff;\n\t"
                 "shfl.sync.up.b32 hi, hi, %2, %3, 0xffffffff;\n\t"
                 "mov.b64 t, {lo, hi};\n\t"
                 "mov.f64 %0, %1;\n\t"
                 "@p add.f64 %0, %1, t;\n\t}"
                 : "=d"(r) : "d"(v), "r"(delta), "r"(cfirst));
    return r;
}

__device__ __forceinline__ double awb_scan_down(double v, int delta, int clast)
{
    double r;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi;\n\t.reg .f64 t;\n\t"
                 "mov.b64 {lo, hi}, %1;\n\t"
                 "shfl.sync.down.b32 lo|p, lo, %2, %3, 0xffffffff;\n\t"
                 "shfl.sync.down.b32 hi, hi, %2, %3, 0xffffffff;\n\t"
                 "mov.b64 t, {lo, hi};\n\t"
                 "mov.f64 %0, %1;\n\t"
                 "@p add.f64 %0, %1, t;\n\t}"
                 : "=d"(r) : "d"(v), "r"(delta), "r"(clast));
    return r;
}

// the neighbour's value inside the segment, 0 at its edge
__device__ __forceinline__ double awb_prev_in_seg(double v, int cfirst)
{
    double r;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi;\n\t"
                 "mov.b64 {lo, hi}, %1;\n\t"
                 "shfl.sync.up.b32 lo|p, lo, 1, %2, 0xffffffff;\n\t"
                 "shfl.sync.up.b32 hi, hi, 1, %2, 0xffffffff;\n\t"
                 "mov.b64 %0, {lo, hi};\n\t"
                 "@!p mov.f64 %0, 0d0000000000000000;\n\t}"
                 : "=d"(r) : "d"(v), "r"(cfirst));
    return r;
}

__device__ __forceinline__ double awb_next_in_seg(double v, int clast)
{
    double r;
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi;\n\t"
                 "mov.b64 {lo, hi}, %1;\n\t"
                 "shfl.sync.down.b32 lo|p, lo, 1, %2, 0xffffffff;\n\t"
                 "shfl.sync.down.b32 hi, hi, 1, %2, 0xffffffff;\n\t"
                 "mov.b64 %0, {lo, hi};\n\t"
                 "@!p mov.f64 %0, 0d0000000000000000;\n\t}"
                 : "=d"(r) : "d"(v), "r"(clast));
    return r;
}

// acc += t when bit BIT of `bits` is set: one LOP3 into a predicate and a
// predicated DADD (a C++ conditional becomes a DADD and two FSEL)
template <unsigned BIT>
__device__ __forceinline__ void awb_add_if(double &acc, double t, unsigned bits)
{
    asm("{\n\t.reg .pred p;\n\t.reg .b32 m;\n\t"
        "and.b32 m, %2, %3;\n\t"
        "setp.ne.u32 p, m, 0;\n\t"
        "@p add.f64 %0, %0, %1;\n\t}"
        : "+d"(acc) : "d"(t), "r"(bits), "n"(BIT));
}

// a + b, or 0 when bit BIT of `bits` is set
template <unsigned BIT>
__device__ __forceinline__ double awb_sum_unless(double a, double b, unsigned bits)
{
    double r;
    asm("{\n\t.reg .pred p;\n\t.reg .b32 m;\n\t"
        "and.b32 m, %3, %4;\n\t"
        "setp.ne.u32 p, m, 0;\n\t"
        "mov.f64 %0, 0d0000000000000000;\n\t"
        "@!p add.f64 %0, %1, %2;\n\t}"
        : "=d"(r) : "d"(a), "d"(b), "r"(bits), "n"(BIT));
    return r;
}

__device__ __forceinline__ double2 awb_lds2(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)
                 : "memory");
    return v;
}

__device__ __forceinline__ void awb_sts2(unsigned addr, double x, double y)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(addr), "d"(x), "d"(y) : "memory");
}

// (compute-sanitizer's synccheck reports these barriers: it flags a named barrier
// that warps of one CTA enter from different places in the code -- what the warp
// roles below do by design; scripts/sync_probe.cu isolates it)
__device__ __forceinline__ void awb_bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

// U: consecutive states per compute thread; MAXREG: registers per thread (what
// lets the wanted number of CTAs share an SM -- __launch_bounds__'s own
// arithmetic rounds 224 threads up to 256 and leaves registers unused)
// BOOKW: 1 = a warp of its own keeps the books (one window per SM: the shortest
// per-site chain), 0 = the second scribe warp does (6-warp CTAs, several per SM)
template <int TMAX, int NLEV, int BOOKW, int U, int MAXREG>
__global__ void __maxnreg__(MAXREG)
awb_forward_fast_kernel(const AwbChain *chains, int seg, int pass, int zcap)
{
    static_assert(TMAX % 4 == 0, "the scribes' R loop takes four rows at a time");
    static_assert(U == 1 || U == 2 || U == 4, "register sets come singly or in pairs");
    // grid: x = chain, y = one of the segments worked on at once (the second
    // pass of a checkpointed table rebuilds independent segments: seg, seg-1, ..)
    const AwbChain &chg = chains[blockIdx.x];
    seg -= (int) blockIdx.y;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int NT = blockDim.x - AWB_FWD_HELPERS - 32 * BOOKW;     // compute threads
    const int NS = NT * U;                           // state slots
    const int NB1 = NT + AWB_FWD_FSCRIBES;           // barrier 1 participants
    const int NB2 = blockDim.x;                      // barrier 2 participants
    const int T = chg.model.ntimes;
    // the blocks / sites of this launch (checkpointed table: one segment)
    const AwbSeg g = awb_seg(chg, seg);
    if (!g.valid)
        return;
    // second pass of a checkpointed table: the last segments' tables are still
    // resident from the first pass
    if (pass == 1 && g.resident)
        return;
    const int n = g.nsites;
    const int bbeg = g.b0, bend = g.b1 + g.extra;
    const int bextra = g.extra ? g.b1 : -1;     // block of which only the first site is done
    const int *__restrict__ nstatesg = chg.nstates;
    const int *__restrict__ blocklensg = chg.blocklens;

    // the small arrays come first, at compile-time offsets from the start of
    // shared memory (the site loops address them every site)
    extern __shared__ double smem_f[];
    constexpr int TMS = (TMAX + 1) / 2 * AWB_NSCRIBE * 2;   // doubles of tmS
    double *FsS = smem_f;                      // [2][TMAX+2] per-time sums
    double *RsS = FsS + 2 * (TMAX + 2);        // [2][TMAX+2] R[b] = sum_a tm[a][b] F[a]
    double *scaleS = RsS + 2 * (TMAX + 2);     // [2] rescale factors
    double *invL = scaleS + 2;                 // [2] [0]: 1/norm of the last column
    double *dummyS = invL + 2;                 // [4] [0]: idle lanes store here; [1] = 1.0;
                                               //   [2] = 0.0
    double *tmS = dummyS + 4;                  // scribe lane sl keeps column sl of the
                                               //   block's time matrix: [(a/2)][sl][a&1]
    double *colS = tmS + TMS;                  // [2][NS] last column of a block in
                                               //   state order, by block parity
    double *cstS = colS + 2 * NS;              // [2][NS][2] per-state constants of the
                                               //   block: (A1, A2) and (A3, inv_emit),
                                               //   each thread's own entries
    double *zT = cstS + 4 * NS;                // [zcap] column, time-major rows,
                                               //   zero-padded for the scribes (K1)

    for (int x = tid; x < 6 * NS + zcap + 4 * (TMAX + 2) + 8 + TMS; x += blockDim.x) {
        const int y = x - 4 * (TMAX + 2);
        smem_f[x] = ((y >= 0 && y < 4) || y == 5) ? 1.0 : 0.0;   // scaleS, invL, one
    }
    __syncthreads();

    // 32-bit shared-memory addresses: generic pointers into dynamic shared
    // memory make the compiler rebuild the shared window (S2UR SR_CgaCtaId)
    // in front of every access, which is slow on the per-site critical path
    const unsigned zT_s = (unsigned) __cvta_generic_to_shared(zT);
    const unsigned col_s = (unsigned) __cvta_generic_to_shared(colS);
    const unsigned Fs_s = (unsigned) __cvta_generic_to_shared(FsS);
    const unsigned Rs_s = (unsigned) __cvta_generic_to_shared(RsS);
    const unsigned scale_s = (unsigned) __cvta_generic_to_shared(scaleS);
    const unsigned invl_s = (unsigned) __cvta_generic_to_shared(invL);
    const unsigned dummy_s = (unsigned) __cvta_generic_to_shared(dummyS);
    const unsigned tm_s = (unsigned) __cvta_generic_to_shared(tmS);
    const unsigned cst_s = (unsigned) __cvta_generic_to_shared(cstS);
    constexpr unsigned RSTR = (TMAX + 2) * 8;

    // ---- who keeps the books: a warp of its own, or the second scribe warp
    const bool bookkeeper = BOOKW ? (tid >= NB1) : (tid >= NT + 32);
    // ---- bookkeeping state (second warp)
    double lprod = 1.0, lacc = 0.0;
    int nprod = 0;
    double *__restrict__ fsumg = chg.fsum + g.fsoff;
    int bad_site = -1;
    // the emission rows of the variant sites a few sites ahead are
    // prefetched into L1 (K3 left them in the table; a compute thread reads
    // its entry right before it needs it).  Cursor (ab, ai) = block / offset
    // of site + LA.
    constexpr int LA = 6;
    const unsigned char *__restrict__ kindn = chg.kind + g.site0;
    const double *fwn = chg.fw - g.fwbias;
    int ab = bbeg, ai = 0, ablen = (ab == bextra) ? 1 : blocklensg[ab];
    int aS1 = 1;
    long long arow = 0;
    unsigned kahead = bookkeeper ? kindn[LA] : 0;
    if (bookkeeper) {
        for (int x = 0; x < LA && ab < bend; x++) {
            if (++ai == ablen) {
                ab++;
                ai = 0;
                ablen = (ab < bend) ? ((ab == bextra) ? 1 : blocklensg[ab]) : 0;
            }
        }
        if (ab < bend) {
            aS1 = nstatesg[ab] > 0 ? nstatesg[ab] : 1;
            arow = chg.fw_off[ab] + (long long) ai * aS1;
        }
    }

    // norm of column s (lane T-1's "R"), per-time sums, rescale factor, logZ
    auto keep_books = [&](int s) {
        const unsigned Fa_s = Fs_s + (s & 1) * RSTR + 8u * lane;
        const double nrm = awb_lds(Rs_s + (s & 1) * RSTR + 8u * (unsigned) (T - 1));
        // (T - 1 <= 63: at most two rows per lane)
        // per-time sums of the column as it is stored (the traceback forms
        // its row totals from these)
        if (lane < T - 1)
            __stcs(fsumg + (size_t) s * (T - 1) + lane, awb_lds(Fa_s));
        if (lane + 32 < T - 1)
            __stcs(fsumg + (size_t) s * (T - 1) + lane + 32, awb_lds(Fa_s + 256u));
        if (!(nrm > 0.0) && bad_site < 0)
            bad_site = s;
#ifdef AWB_K4_DEBUG_NRM
        if (lane == 0 && s < 64) chg.sink[s] = nrm;
#endif
        if ((s & (AWB_FWD_RS - 1)) == 0) {
            // this factor is applied when column s+3 is formed
            if (lane == 0)
                awb_sts(scale_s + 8u * ((s / AWB_FWD_RS) & 1), 1.0 / nrm);
            if (s + 3 <= n - 1) {
                lprod *= nrm;
                if (++nprod == 8) {
                    lacc += log(lprod);
                    lprod = 1.0;
                    nprod = 0;
                }
            }
        }
        if (s == n - 1 && lane == 0) {
            awb_sts(invl_s, 1.0 / nrm);
            // (a segment that starts from a stored, normalised column adds
            // the log-likelihood of its own sites; the recompute pass adds
            // nothing)
            const double lz = log(nrm) + log(lprod) + lacc;
            if (pass == 0) {
                if (seg <= 0 || !chg.ckpt) {
                    chg.logz[0] = lz;
                    chg.status[0] = bad_site < 0 ? -1 : g.site0 + bad_site;
                } else {
                    chg.logz[0] += lz;
                    if (bad_site >= 0 && chg.status[0] < 0)
                        chg.status[0] = g.site0 + bad_site;
                }
            }
        }
    };


    auto prefetch_next_block = [&](int b) {
        if (b + 1 < bend) {
        const int nb = b + 1;
        const long long r0n = chg.row_off[nb];
        const long long S1n = chg.row_off[nb + 1] - r0n;
        const long long tr0n = chg.trow_off[nb];
        const long long e0n = chg.ent_off[nb];
        const long long nen = chg.ent_off[nb + 1] - e0n;
        // (one compact loop over the 16 tables: this code runs once per
        // block, cold in the instruction cache, and inlined range by
        // range it was several hundred instructions)
        const void *pp[16];
        long long pl[16];
        pp[0] = chg.tmap + tr0n;      pl[0] = 2 * (chg.trow_off[nb + 1] - tr0n);
        pp[1] = chg.st_node + r0n;    pl[1] = 2 * S1n;
        pp[2] = chg.st_time + r0n;    pl[2] = S1n;
        pp[3] = chg.st_age + r0n;     pl[3] = S1n;
        pp[4] = chg.iperm + r0n;      pl[4] = 2 * S1n;
        pp[5] = chg.inv_emit + r0n;   pl[5] = 8 * S1n;
        pp[6] = chg.sw_start + r0n;   pl[6] = 2 * S1n;
        pp[7] = chg.sw_cnt + r0n;     pl[7] = 2 * S1n;
        pp[8] = chg.sw_src + e0n;     pl[8] = 2 * nen;
        pp[9] = chg.sw_prob + e0n;    pl[9] = 8 * nen;
        pp[10] = chg.lin + (size_t) nb * 7 * T;              pl[10] = 56ll * T;
        pp[11] = chg.tmatrix + (size_t) nb * T * T;          pl[11] = 8ll * T * T;
        pp[12] = chg.sc_start + (size_t) nb * AWB_NSCRIBE;   pl[12] = 2 * AWB_NSCRIBE;
        pp[13] = chg.sc_cnt + (size_t) nb * AWB_NSCRIBE;     pl[13] = 2 * AWB_NSCRIBE;
        pp[14] = chg.sc_row + (size_t) nb * AWB_NSCRIBE;     pl[14] = AWB_NSCRIBE;
        pp[15] = chg.sc_stride + (size_t) nb * AWB_NSCRIBE;  pl[15] = AWB_NSCRIBE;
#pragma unroll 1
        for (int r = 0; r < 16; r++)
            awb_prefetch_range(pp[r], pl[r], lane);
        if (lane == 0) {
            asm volatile("prefetch.global.L1 [%0];" :: "l"(nstatesg + nb));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(blocklensg + nb));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(chg.fw_off + nb));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(chg.sc_ch + nb));
        }
    }
    };

    // emission rows LA sites ahead (the kind byte was loaded one site ago:
    // nothing here waits for global memory)
    auto look_ahead = [&](int site) {
        const unsigned kcur = kahead;
        kahead = kindn[site + 1 + LA];
        if (ab < bend) {
            if (kcur == AWB_SITE_VARIANT)
                awb_prefetch_range(fwn + arow, 8ll * aS1, lane);
            arow += aS1;
            if (++ai == ablen) {
                ab++;
                ai = 0;
                if (ab < bend) {
                    ablen = (ab == bextra) ? 1 : blocklensg[ab];
                    aS1 = nstatesg[ab] > 0 ? nstatesg[ab] : 1;
                    arow = chg.fw_off[ab];
                }
            }
        }
    };

    if (BOOKW && tid >= NB1) {
        // =================================================================
        // bookkeeping warp: waits on barrier 2 only
        // =================================================================
        int site = 0;
        for (int b = bbeg; b < bend; b++) {
            const int blen = (b == bextra) ? 1 : blocklensg[b];
            prefetch_next_block(b);
            for (int i = 0; i < blen; i++, site++) {
                look_ahead(site);
                __syncwarp();
                awb_bar_sync(2, NB2);
                keep_books(site);
            }
        }
        __syncthreads();                                   // final barrier
        return;
    }

    if (tid >= NT) {
        // =================================================================
        // F-scribes: per-time sums and R between barrier 1 and barrier 2; the
        // second scribe warp also keeps the books (column norm, logZ, the
        // lagged rescale factor, fsum for the traceback) while the first one
        // forms R, and warms the cache in its idle time
        // =================================================================
        const int sl = tid - NT;                         // scribe lane 0..63
        const double *__restrict__ tmatrixg = chg.tmatrix;
        const unsigned short *__restrict__ sc_startg = chg.sc_start;
        const unsigned short *__restrict__ sc_cntg = chg.sc_cnt;
        const unsigned char *__restrict__ sc_rowg = chg.sc_row;
        const unsigned char *__restrict__ sc_strideg = chg.sc_stride;
        const unsigned tml_s = tm_s + 16u * (unsigned) sl;      // my column, pair 0

        int site = 0;
        for (int b = bbeg; b < bend; b++) {
            const int blen = (b == bextra) ? 1 : blocklensg[b];
            const int sc_start = sc_startg[(size_t) b * AWB_NSCRIBE + sl];
            const int sc_cnt = sc_cntg[(size_t) b * AWB_NSCRIBE + sl];
            const int sc_row = sc_rowg[(size_t) b * AWB_NSCRIBE + sl];
            // the lanes of a row read it interleaved: element q of this lane is
            // zstep bytes after element q-1
            const unsigned zstep = 8u * sc_strideg[(size_t) b * AWB_NSCRIBE + sl];
            const int CH = chg.sc_ch[b];            // slots every lane sums (even)
            const int key = (sc_row != 255) ? sc_row : (0x100 + lane);
            int seglane, segend;
            awb_lane_run(key, lane, seglane, segend);
            const bool sc_last = (lane == segend) && (sc_row != 255);
            const int span = __reduce_max_sync(0xffffffffu, segend - seglane);
            const unsigned z_s = zT_s + 8u * (unsigned) sc_start;
            // the slots of this lane beyond its real ones are padding: zero them
            // for this block (the compute warps only write real slots; the
            // previous block is past its last barrier 2).  A lane without a row
            // owns no slots (its sc_start is 0: zeroing "its" padding would race
            // with the compute warp that stores slot 0 of the new block).
            if (sc_row != 255)
                for (int q = sc_cnt; q < CH; q++)
                    awb_sts(z_s + zstep * (unsigned) q, 0.0);
            // column `sl` of the block's time-by-time matrix, parked in this
            // lane's own shared-memory slots (nobody else reads them): this
            // lane turns the per-time sums F into R[sl] for the compute warps
            // Lane T-1 holds a column of ones: its "R" is the column norm.
            if (sl != T - 1 || b == bbeg) {
                const bool rl = (sl < T - 1) && nstatesg[b] > 0;
                const double one = (sl == T - 1) ? 1.0 : 0.0;
                const double *tmg = tmatrixg + (size_t) b * T * T + (rl ? sl : 0);
#pragma unroll 4
                for (int a = 0; a < TMAX; a += 2) {
                    const double t0 = (a < T - 1) ? (rl ? tmg[a * T] : one) : 0.0;
                    const double t1 = (a + 1 < T - 1) ? (rl ? tmg[(a + 1) * T] : one) : 0.0;
                    awb_sts2(tml_s + (unsigned) (a / 2) * (16u * AWB_NSCRIBE), t0, t1);
                }
            }
            if (bookkeeper)
                prefetch_next_block(b);
            // (the per-block code above diverges, and bar.sync is an aligned
            // barrier: every lane of the warp has to arrive together)
            __syncwarp();

            for (int i = 0; i < blen; i++, site++) {
                const unsigned Fp_s = Fs_s + (site & 1) * RSTR;
                awb_bar_sync(1, NB1);
                // each lane sums its CH slots of one (zero-padded) row; the lanes of
                // a row combine with a segmented scan
                double v0 = 0.0, v1 = 0.0;
                {
                    unsigned a = z_s;
#pragma unroll 4
                    for (int q = 0; q < CH; q += 2) {
                        v0 += awb_lds(a);
                        v1 += awb_lds(a + zstep);
                        a += 2u * zstep;
                    }
                }
                double v = v0 + v1;
#pragma unroll
                for (int l = 0; l < 5; l++) {
                    if ((1 << l) <= span)
                        v = awb_scan_up(v, 1 << l, seglane);
                }
                if (sc_last)
                    awb_sts(Fp_s + 8u * (unsigned) sc_row, v);
                awb_bar_sync(3, AWB_FWD_FSCRIBES);
                if ((i + 1 < blen && sl < T - 1) || sl == T - 1) {
                    double ra = 0.0, rb = 0.0, rc = 0.0, rd = 0.0;
#pragma unroll
                    for (int a = 0; a + 3 < TMAX; a += 4) {
                        const double2 f0 = awb_lds2(Fp_s + 8u * a);
                        const double2 f1 = awb_lds2(Fp_s + 8u * a + 16u);
                        const double2 t0 =
                            awb_lds2(tml_s + (unsigned) (a / 2) * (16u * AWB_NSCRIBE));
                        const double2 t1 =
                            awb_lds2(tml_s + (unsigned) (a / 2 + 1) * (16u * AWB_NSCRIBE));
                        ra = fma(t0.x, f0.x, ra);
                        rb = fma(t0.y, f0.y, rb);
                        rc = fma(t1.x, f1.x, rc);
                        rd = fma(t1.y, f1.y, rd);
                    }
                    awb_sts(Rs_s + (site & 1) * RSTR + 8u * (unsigned) sl,
                            (ra + rb) + (rc + rd));
                }
                // the books of the PREVIOUS site (its norm and per-time sums stay
                // in their buffers until the site after this one)
                if (bookkeeper) {
                    if (site > 0)
                        keep_books(site - 1);
                    look_ahead(site);
                }
                awb_bar_sync(2, NB2);
            }
        }
        if (bookkeeper)
            keep_books(n - 1);
        __syncthreads();                                   // final barrier
        return;
    }

    // =====================================================================
    // compute warps
    // =====================================================================
    // Every instruction of this loop is issued by all compute warps per site on
    // 4 schedulers, so the loop is written for instruction count: 32-bit shared
    // addresses with ring offsets kept incrementally, idle lanes pointing at a
    // per-thread sink (no null checks or selects).
    const unsigned char *__restrict__ kindg = chg.kind + g.site0;
    double *__restrict__ fwg = chg.fw - g.fwbias;
    const int warp = tid >> 5;
    double *const sink = chg.sink + tid;

    // ---- my states in the current block: U consecutive slots of the packed,
    // node-major order, slot = (warp*32 + lane)*U + u.  A branch's states are
    // consecutive slots (ascending time), so a lane holds a run of a branch, and
    // the two exclusive sums along a branch are a scan inside the lane plus a
    // segmented scan of one value per lane across the warp.
    int jj[U];
    bool active[U], live[U];               // live: active and S > 0
    unsigned actm = 0;                     // bit u: active[u]
    unsigned zaddr[U], raddr[U];
    double Da[U], H[U];
    double c[U];
    constexpr bool CREG = (U <= 2);        // few states per thread: all constants in registers
    double A1r[CREG ? U : 1], A2r[CREG ? U : 1], A3r[CREG ? U : 1], ier[CREG ? U : 1];
    // the constants only needed after the scans stay in shared memory:
    // (A1, A2) at cA + 512 u, (A3, inv_emit) at cB + 512 u (16 bytes per lane)
    unsigned cA = cst_s + 16u * (unsigned) (warp * U * 32 + lane);
    asm volatile("" : "+r"(cA));           // (keep it: recomputing it costs ten instructions a site)
    const unsigned cB = cA + 16u * (unsigned) NS;
    double *nxt[U];
    int S = 0, S1 = 1;
    long long r0 = 0;
    unsigned rowstep = 0;                  // bytes between consecutive rows of the block
    // bit u of headm / lastm: slot u is the first / last state of its branch;
    // prem / postm: the slots in front of my first head / behind my last `last`
    // (they continue a branch of the lanes below / above)
    unsigned headm = 0, lastm = 0, prem = 0, postm = 0;
    // carry scans across the lanes: bit l of upok / dnok = the value 2^l lanes
    // below / above still belongs to my branch
    unsigned upok = 0, dnok = 0;

    auto load_compute = [&](int bb) {
        S = nstatesg[bb];
        S1 = S > 0 ? S : 1;
        r0 = chg.row_off[bb];
        rowstep = 8u * (unsigned) S1;
        const long long tr0 = chg.trow_off[bb];
        const int NSb = (int) (chg.trow_off[bb + 1] - tr0);
        int node[U];
        actm = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int slot = (warp * 32 + lane) * U + u;
            unsigned short tj = 0xFFFF;
            if (slot < NSb)
                tj = chg.tmap[tr0 + slot];
            active[u] = (tj != 0xFFFF);
            if (active[u]) actm |= 1u << u;
            live[u] = active[u] && S > 0;
            jj[u] = active[u] ? (int) tj : 0;
            int atime = 0, cage = 0, tpos = 0;
            node[u] = -1 - u;
            double inv_e = active[u] ? 1.0 : 0.0;
            double A1, A2, A3;
            if (live[u]) {
                atime = chg.st_time[r0 + jj[u]];
                cage = chg.st_age[r0 + jj[u]];
                node[u] = chg.st_node[r0 + jj[u]];
                tpos = chg.iperm[r0 + jj[u]];
                inv_e = chg.inv_emit[r0 + jj[u]];
            }
            zaddr[u] = active[u] ? zT_s + 8u * (unsigned) tpos : dummy_s;
            raddr[u] = Rs_s + 8u * (unsigned) atime;
            if (live[u]) {
                const double *lin = chg.lin + (size_t) bb * 7 * T;
                const double Bc = cage > 0 ? lin[2 * T + cage - 1] : 0.0;
                const double a1 = lin[3 * T + atime];
                Da[u] = lin[0 * T + atime];
                H[u] = Da[u] * (lin[1 * T + atime] - Bc);
                A1 = a1;
                A2 = fma(Da[u], lin[4 * T + atime] - a1 * Bc, lin[6 * T + atime]);
                A3 = lin[5 * T + atime] - a1 * Bc;
            } else {
                // idle lane (everything 0), or the size-1 state space (identity)
                Da[u] = 0.0; H[u] = 0.0; A1 = 0.0; A3 = 0.0;
                A2 = active[u] ? 1.0 : 0.0;
            }
            if constexpr (CREG) {
                A1r[u] = A1; A2r[u] = A2; A3r[u] = A3; ier[u] = inv_e;
            } else {
                awb_sts2(cA + 512u * u, A1, A2);
                awb_sts2(cB + 512u * u, A3, inv_e);
            }
        }
        // branch structure of my slots (a slot that holds no state is a branch
        // of its own: node < 0, different for neighbours)
        const int below = __shfl_up_sync(0xffffffffu, node[U - 1], 1);
        const int above = __shfl_down_sync(0xffffffffu, node[0], 1);
        headm = lastm = 0;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int pn = (u == 0) ? (lane == 0 ? -100 : below) : node[u - 1];
            const int nn = (u == U - 1) ? (lane == 31 ? -100 : above) : node[u + 1];
            if (node[u] < 0 || pn != node[u]) headm |= 1u << u;
            if (node[u] < 0 || nn != node[u]) lastm |= 1u << u;
        }
        constexpr unsigned FULL = (1u << U) - 1u;
        prem = headm ? (((headm & (0u - headm)) - 1u) & FULL) : FULL;
        postm = lastm ? (~((2u << (31 - __clz(lastm))) - 1u) & FULL) : FULL;
        const unsigned hb = __ballot_sync(0xffffffffu, headm != 0);
        const unsigned lb = __ballot_sync(0xffffffffu, lastm != 0);
        const unsigned hbelow = hb & ((1u << lane) - 1u);
        const unsigned labove = (lane == 31) ? 0u : (lb & (0xffffffffu << (lane + 1)));
        // my first branch starts in lane L0 (its last head), my last one ends in
        // lane L1 (its first `last`).  The carries are scans of one value per
        // lane, shifted by one lane: the value of lane x sits in lane x + 1 (up)
        // or x - 1 (down), so the valid range is [L0 + 1, me] / [me, L1 - 1]
        const int L0 = hbelow ? (31 - __clz(hbelow)) : 0;
        const int L1 = labove ? (__ffs(labove) - 1) : 31;
        upok = dnok = 0;
#pragma unroll
        for (int l = 0; l < NLEV; l++) {
            if (lane - (1 << l) >= L0 + 1) upok |= 1u << l;
            if (lane + (1 << l) <= L1 - 1) dnok |= 1u << l;
        }
    };

    load_compute(bbeg);
#pragma unroll
    for (int u = 0; u < U; u++) {
        c[u] = 0.0;
        nxt[u] = sink;
        if (active[u]) {
            // first column: the prior (K1 or caller), kept as given.  With a
            // checkpointed table the prior is saved aside on the first pass (the
            // table memory is reused by the other segments) and put back for the
            // second; a later segment starts from its stored first column, which
            // goes into the table like any other column.
            double *row0 = fwg + chg.fw_off[bbeg] + jj[u];
            if (!chg.ckpt) {
                c[u] = *row0;
            } else if (seg == 0) {
                // (K1 / the caller put the prior at the start of table 0; the
                // segment's own table may be another one)
                if (pass == 0) {
                    c[u] = chg.fw[chg.fw_off[bbeg] + jj[u]];
                    chg.ckptcol[jj[u]] = c[u];
                } else {
                    c[u] = chg.ckptcol[jj[u]];
                }
                *row0 = c[u];
            } else {
                c[u] = chg.ckptcol[(size_t) seg * chg.maxS + jj[u]];
                *row0 = c[u];
            }
            nxt[u] = row0 + S1;
        }
    }
    const unsigned char *kp = kindg + 2;
    unsigned kind_next = (n > 1) ? kindg[1] : 0;
    // ring offsets (bytes): rofs = (site&1)*RSTR into Rs, sofs = (((site-2)/RS)&1)*8
    // into scaleS; rcnt = (site - 2) & (RS - 1)
    unsigned rofs = 0, sofs = 0;
    int rcnt = AWB_FWD_RS - 2;

    // one site that is followed by a site of the same block
    auto site_step = [&]() {
#pragma unroll
        for (int u = 0; u < U; u++)
            awb_sts(zaddr[u], c[u]);
        awb_bar_sync(1, NB1);
        // (nothing below may move in front of the barrier the scribes wait at --
        // ptxas would hoist the register-only scans: they hang on a value, 0.0,
        // loaded behind it)
        const double zero = awb_lds(dummy_s + 16u);

        // branch sums in registers while the F-scribes sum the rows.
        // Exclusive sums PY = sum_{a<b} x_a (h[a] - Bc), Q = sum_{a>b} x_a along
        // the branch (exclusive by construction: an inclusive scan minus the own
        // term cancels -- h grows like exp(cumulative coalescent rate), so a
        // term can exceed the sum below it by many orders): first inside the
        // lane, then one carry per lane and direction across the warp.
        double py[U], q[U], x0[U], y0[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            x0[u] = fma(Da[u], c[u], zero);
            y0[u] = fma(H[u], c[u], zero);
        }
        py[0] = 0.0;
        q[U - 1] = 0.0;
        if constexpr (U == 2) {
            py[1] = awb_sum_unless<2u>(py[0], y0[0], headm);
            q[0] = awb_sum_unless<1u>(q[1], x0[1], lastm);
        }
        if constexpr (U == 4) {
            py[1] = awb_sum_unless<2u>(py[0], y0[0], headm);
            py[2] = awb_sum_unless<4u>(py[1], y0[1], headm);
            py[3] = awb_sum_unless<8u>(py[2], y0[2], headm);
            q[2] = awb_sum_unless<4u>(q[3], x0[3], lastm);
            q[1] = awb_sum_unless<2u>(q[2], x0[2], lastm);
            q[0] = awb_sum_unless<1u>(q[1], x0[1], lastm);
        }
        // vU: what my last branch has so far (for the lanes above), vD: what my
        // first branch has from here up (for the lanes below); shifted by one
        // lane (the end lanes receive their own value and never use it)
        double vU = __shfl_up_sync(0xffffffffu, py[U - 1] + y0[U - 1], 1);
        double vD = __shfl_down_sync(0xffffffffu, q[0] + x0[0], 1);
        {
            double t;
#define AWB_CARRY_LEVEL(LV)                                                   \
            if (NLEV > LV) {                                                  \
                t = __shfl_up_sync(0xffffffffu, vU, 1 << LV);                 \
                awb_add_if<(1u << LV)>(vU, t, upok);                          \
                t = __shfl_down_sync(0xffffffffu, vD, 1 << LV);               \
                awb_add_if<(1u << LV)>(vD, t, dnok);                          \
            }
            AWB_CARRY_LEVEL(0)
            AWB_CARRY_LEVEL(1)
            AWB_CARRY_LEVEL(2)
            AWB_CARRY_LEVEL(3)
            AWB_CARRY_LEVEL(4)
#undef AWB_CARRY_LEVEL
        }
        awb_add_if<1u>(py[0], vU, prem);
        awb_add_if<(1u << (U - 1))>(q[U - 1], vD, postm);
        if constexpr (U == 2) {
            awb_add_if<2u>(py[1], vU, prem);
            awb_add_if<1u>(q[0], vD, postm);
        }
        if constexpr (U == 4) {
            awb_add_if<2u>(py[1], vU, prem);
            awb_add_if<4u>(py[2], vU, prem);
            awb_add_if<8u>(py[3], vU, prem);
            awb_add_if<4u>(q[2], vD, postm);
            awb_add_if<2u>(q[1], vD, postm);
            awb_add_if<1u>(q[0], vD, postm);
        }
        const unsigned kd = kind_next;
        kind_next = *kp++;
        // the lagged rescale factor every fourth site ((site & 3) == 2), the
        // constant 1.0 otherwise: branch-free.  (The factor was published after
        // barrier 2 of an earlier site: it can be read in front of this one.)
        const bool resc = rcnt == 0;
        const double sc = awb_lds(resc ? scale_s + sofs : dummy_s + 8u);
        double W[U], e[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            if constexpr (CREG) {
                W[u] = fma(A1r[u], py[u], fma(c[u], A2r[u], A3r[u] * q[u]));
                e[u] = ier[u];
            } else {
                const double2 a12 = awb_lds2(cA + 512u * u);
                const double2 a3e = awb_lds2(cB + 512u * u);
                W[u] = fma(a12.x, py[u], fma(c[u], a12.y, a3e.x * q[u]));
                e[u] = a3e.y;
            }
        }
        if (kd != AWB_SITE_INVARIANT) {             // uniform, ~3 % of the sites
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (kd == AWB_SITE_VARIANT)
                    e[u] = live[u] ? *nxt[u] : e[u];
                else
                    e[u] = active[u] ? 1.0 : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
            e[u] *= sc;
        awb_bar_sync(2, NB2);

        // R[atime] = sum_a tm[a][atime] * F[a], formed by the F-scribes
#pragma unroll
        for (int u = 0; u < U; u++) {
            const double cn = (awb_lds(raddr[u] + rofs) + W[u]) * e[u];
            if (actm & (1u << u))
                __stcs(nxt[u], cn);             // streaming: L1 is for the block tables
            c[u] = cn;
            nxt[u] = (double *) ((char *) nxt[u] + rowstep);
        }
        sofs ^= resc ? 8u : 0u;
        rcnt = (rcnt + 1) & (AWB_FWD_RS - 1);
        rofs ^= RSTR;
    };

    for (int b = bbeg; b < bend; b++) {
        const int blen = (b == bextra) ? 1 : blocklensg[b];

        // ---------------- sites that are followed by a site of the same block
#ifdef AWB_K4_UNROLL1
#pragma unroll 1
#endif
        for (int i = blen - 1; i > 0; i--)
            site_step();

        // ---------------- last site of the block
        {
#pragma unroll
            for (int u = 0; u < U; u++) {
                awb_sts(zaddr[u], c[u]);
                awb_sts(active[u] ? col_s + 8u * (unsigned) ((b & 1) * NS + jj[u]) : dummy_s,
                        c[u]);
            }
            awb_bar_sync(1, NB1);
            const unsigned kd = kind_next;
            kind_next = *kp++;
            awb_bar_sync(2, NB2);
            if (b == bend - 1)
                break;

            // breakpoint: gather through the switch CSR (sample_thread.cpp:345-389)
            double scale = 1.0;
            if (rcnt == 0) {
                scale = awb_lds(scale_s + sofs);
                sofs ^= 8u;
            }
            rcnt = (rcnt + 1) & (AWB_FWD_RS - 1);
            rofs ^= RSTR;
            load_compute(b + 1);
            // the old block's last column; the buffer alternates with the
            // block so a one-site block cannot overwrite it early
            const double *cold = colS + (b & 1) * NS;
            const long long e0 = chg.ent_off[b + 1];
            double *rowb = fwg + chg.fw_off[b + 1];
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (active[u]) {
                    const int st = chg.sw_start[r0 + jj[u]];
                    const int cnt = chg.sw_cnt[r0 + jj[u]];
                    const unsigned short *es = chg.sw_src + e0 + st;
                    const double *ep = chg.sw_prob + e0 + st;
                    double sum = 0.0;
                    for (int x = 0; x < cnt; x++)
                        sum += cold[es[x]] * ep[x];
                    double *w0 = rowb + jj[u];
                    double e = 1.0;
                    if (S > 0) {
                        if constexpr (CREG)
                            e = ier[u];
                        else
                            e = awb_lds2(cB + 512u * u).y;
                        if (kd == AWB_SITE_VARIANT)
                            e = *w0;
                        else if (kd == AWB_SITE_MASKED)
                            e = 1.0;
                    }
                    c[u] = sum * e * scale;
                    __stcs(w0, c[u]);
                    nxt[u] = w0 + S1;
                } else {
                    c[u] = 0.0;
                    nxt[u] = sink;
                }
            }
            __syncwarp();       // (the gather above diverges)
        }
    }

    // ---- the column of the extra site is the first column of the next segment
    // (normalised: its 1/norm is complete after the final barrier)
    __syncthreads();
    if (g.extra && pass == 0) {
        const double il = invL[0];
#pragma unroll
        for (int u = 0; u < U; u++)
            if (active[u])
                chg.ckptcol[(size_t) (seg + 1) * chg.maxS + jj[u]] = c[u] * il;
    }
}

#endif // AWB_FORWARD_FAST_CUH
