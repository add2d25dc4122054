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
// with x_a = D[a]*col[(n,a)] along the branch n of state k = (n, b):
//
//   W_k = E e2[b] * PY + x_b * A2 + A3 * Q + norecombs[b]*col[k]
//   PY = sum_{a<b} x_a (h[a] - Bc)   Q = sum_{a>b} x_a
//
// (Bc, A2, A3 per-state constants, see load_compute).  PY and Q are exclusive
// prefix / suffix sums along one branch; the separable form agrees with the
// literal band to ~4e-15 relative (checked on random columns).
//
// Mapping to the SM (numbers measured on B200, scripts/microbench.cu:
// dependent DFMA 8.4 cycles, 64-bit SHFL+DADD 35, LDS 33, bar.sync 33,
// DDIV 130).  A site step is a chain of ~30 dependent shared-memory / shuffle /
// FP64 operations, not a bandwidth problem, so the kernel is organised around
// that chain:
//
//   compute warps   one thread per state in NODE-MAJOR order, packed
//                   (first-fit decreasing, K1) so that the states of a branch
//                   never straddle a warp: PY and Q are segmented warp-shuffle
//                   scans in registers, branch-free (0/1 multipliers, one DFMA
//                   per level and quantity); a warp runs only the levels its
//                   longest branch needs.
//   F-scribes       two warps owning no state: between barrier 1 and barrier 2
//                   their 64 lanes turn the column (staged in shared memory in
//                   TIME-MAJOR order) into the per-time sums F[a] and then,
//                   one lane per b with the matrix column in registers, into
//                   R[b]; a compute thread reads a single value.
//   norm warp       one warp that only waits on barrier 2: column norm, 1/norm,
//                   logZ, rescale factor, and the per-time sums of the stored
//                   column (fsum) for the traceback.  Its division and log()
//                   are off the critical cycle; consumers pick the results up
//                   two steps later.
//
//   step(site):  STS value (time-major slot)
//                B1 (compute + F-scribes)
//                    compute:  store column site-2 to HBM scaled by its 1/norm ;
//                              fetch next emission ; branch scans -> W
//                    F-scribes: F[a] ; B3 (scribes) ; R[b]
//                B2 (everybody)
//                    compute:  col(site+1) = (R[b] + W) * emission
//                    norm warp: norm(site) ...
//
// The forward table is written once, 8 B per site*state, in the reference's
// state order.  Columns are carried unnormalised; a factor published for
// column s is applied when column s+3 is formed (every AWB_FWD_RS = 4 sites),
// which bounds the magnitude by the product of at most RS+2 one-step norms.
//
// A launch works on one AwbSeg: the whole window, or one segment of a
// checkpointed table (awb_common.cuh, awb_api.cu).
#ifndef AWB_FORWARD_FAST_CUH
#define AWB_FORWARD_FAST_CUH

#include "awb_common.cuh"

#define AWB_FWD_RS 4          // rescale period (sites); power of two, >= 4
#define AWB_FWD_FSCRIBES AWB_NSCRIBE          // F-scribe lanes
#define AWB_FWD_HELPERS (AWB_NSCRIBE + 32)    // F-scribes + norm warp

// shared memory (doubles): Fs[2][TMAX+2] | Rs[2][TMAX+2] | scaleS[2] | invS[4] |
// dummy[2] | colS[2][NS] | zT[zcap]
__host__ __device__ inline size_t awb_fwd_fast_smem_bytes(int NS, int TMAX, int zcap)
{
    return (2 * (size_t) NS + (size_t) zcap + 4 * (size_t) (TMAX + 2) + 6 + 2) * sizeof(double);
}

template <int N> struct AwbInt { static constexpr int value = N; };

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

__device__ __forceinline__ double2 awb_lds2(unsigned addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)
                 : "memory");
    return v;
}

__device__ __forceinline__ void awb_bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory");
}

template <int TMAX, int NLEV, int MAXTHREADS>
__global__ void __launch_bounds__(MAXTHREADS, 1)
awb_forward_fast_kernel(const AwbChain *chains, int seg, int pass, int zcap)
{
    const AwbChain &chg = chains[blockIdx.x];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int NS = blockDim.x - AWB_FWD_HELPERS;     // compute threads
    const int NB1 = NS + AWB_FWD_FSCRIBES;           // barrier 1 participants
    const int NB2 = NS + AWB_FWD_HELPERS;            // barrier 2 participants
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
    // shared memory (the site loops address them every site; offsets that
    // depend on the CTA size were being recomputed from blockDim there)
    extern __shared__ double smem_f[];
    double *FsS = smem_f;                      // [2][TMAX+2] per-time sums
    double *RsS = FsS + 2 * (TMAX + 2);        // [2][TMAX+2] R[b] = sum_a tm[a][b] F[a]
    double *scaleS = RsS + 2 * (TMAX + 2);     // [2] rescale factors
    double *invS = scaleS + 2;                 // [4] 1/norm of recent columns
    double *dummyS = invS + 4;                 // [2] [0]: idle lanes store here; [1] = 1.0
    double *colS = dummyS + 2;                 // [2][NS] last column of a block in
                                               //   state order, by block parity
    double *zT = colS + 2 * NS;                // [zcap] column, time-major rows,
                                               //   zero-padded for the scribes (K1)

    for (int x = tid; x < 2 * NS + zcap + 4 * (TMAX + 2) + 8; x += blockDim.x) {
        const int y = x - 4 * (TMAX + 2);
        smem_f[x] = ((y >= 0 && y < 6) || y == 7) ? 1.0 : 0.0;   // scaleS, invS, one
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
    const unsigned inv_s = (unsigned) __cvta_generic_to_shared(invS);
    const unsigned dummy_s = (unsigned) __cvta_generic_to_shared(dummyS);
    constexpr unsigned RSTR = (TMAX + 2) * 8;

    if (tid >= NB1) {
        // =================================================================
        // norm warp: waits on barrier 2 only
        // =================================================================
        double lprod = 1.0, lacc = 0.0;
        int nprod = 0;
        double *__restrict__ fsumg = chg.fsum + g.fsoff;
        int bad_site = -1;
        // This warp has slack, so it also warms the cache for the others: when
        // a block starts, the per-block tables of the NEXT block (what
        // load_compute, the switch gather and the scribes read in dependent
        // chains at the block boundary) are prefetched into L1.
        int pb = bbeg, pnext = 0;
        // ... and the emission rows of the variant sites a few sites ahead
        // (K3 left them in the table; a compute thread reads its entry right
        // before it needs it).  Cursor (ab, ai) = block / offset of site + LA.
        constexpr int LA = 6;
        const unsigned char *__restrict__ kindn = chg.kind + g.site0;
        const double *fwn = chg.fw - g.fwbias;
        int ab = bbeg, ai = 0, ablen = (ab == bextra) ? 1 : blocklensg[ab];
        for (int x = 0; x < LA && ab < bend; x++) {
            if (++ai == ablen) {
                ab++;
                ai = 0;
                ablen = (ab < bend) ? ((ab == bextra) ? 1 : blocklensg[ab]) : 0;
            }
        }
        int aS1 = 1;
        long long arow = 0;
        if (ab < bend) {
            aS1 = nstatesg[ab] > 0 ? nstatesg[ab] : 1;
            arow = chg.fw_off[ab] + (long long) ai * aS1;
        }
        for (int site = 0; site < n; site++) {
            if (ab < bend) {
                if (kindn[site + LA] == AWB_SITE_VARIANT)
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
            if (site == pnext) {
                pnext += (pb == bextra) ? 1 : blocklensg[pb];
                const int nb = ++pb;
                if (nb < bend) {
                    const long long r0n = chg.row_off[nb];
                    const long long S1n = chg.row_off[nb + 1] - r0n;
                    const long long tr0n = chg.trow_off[nb];
                    const long long e0n = chg.ent_off[nb];
                    const long long nen = chg.ent_off[nb + 1] - e0n;
                    // (one compact loop over the 16 tables: this code runs once
                    // per block, cold in the instruction cache, and inlined
                    // range by range it was several hundred instructions)
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
            }
            const unsigned Fa_s = Fs_s + (site & 1) * RSTR + 8u * lane;
            awb_bar_sync(2, NB2);
            // (T - 1 <= 63: at most two rows per lane)
            const double f0 = (lane < T - 1) ? awb_lds(Fa_s) : 0.0;
            const double f1 = (lane + 32 < T - 1) ? awb_lds(Fa_s + 256u) : 0.0;
            double x = f0 + f1;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1)
                x += __shfl_xor_sync(0xffffffffu, x, d);
            const double nrm = x;
            const double inv = 1.0 / nrm;
            if (lane == 0)
                awb_sts(inv_s + 8u * (site & 3), inv);
            // per-time sums of the column as it is stored (the traceback forms
            // its row totals from these); the prior column is stored unscaled
            {
                const double sc = (site == 0 && !(chg.ckpt && seg > 0)) ? 1.0 : inv;
                if (lane < T - 1)
                    __stcs(fsumg + (size_t) site * (T - 1) + lane, f0 * sc);
                if (lane + 32 < T - 1)
                    __stcs(fsumg + (size_t) site * (T - 1) + lane + 32, f1 * sc);
            }
            if (!(nrm > 0.0) && bad_site < 0)
                bad_site = site;
            if ((site & (AWB_FWD_RS - 1)) == 0) {
                // this factor is applied when column site+3 is formed
                if (lane == 0)
                    awb_sts(scale_s + 8u * ((site / AWB_FWD_RS) & 1), inv);
                if (site + 3 <= n - 1) {
                    lprod *= nrm;
                    if (++nprod == 8) {
                        lacc += log(lprod);
                        lprod = 1.0;
                        nprod = 0;
                    }
                }
            }
            if (site == n - 1 && lane == 0) {
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
        }
        __syncthreads();                                   // final barrier
        return;
    }

    if (tid >= NS) {
        // =================================================================
        // F-scribes: per-time sums between barrier 1 and barrier 2
        // =================================================================
        const int sl = tid - NS;                         // scribe lane 0..63
        const double *__restrict__ tmatrixg = chg.tmatrix;
        const unsigned short *__restrict__ sc_startg = chg.sc_start;
        const unsigned short *__restrict__ sc_cntg = chg.sc_cnt;
        const unsigned char *__restrict__ sc_rowg = chg.sc_row;
        const unsigned char *__restrict__ sc_strideg = chg.sc_stride;
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
            double um[5];
#pragma unroll
            for (int l = 0; l < 5; l++)
                um[l] = (lane - (1 << l) >= seglane) ? 1.0 : 0.0;
            const unsigned z_s = zT_s + 8u * (unsigned) sc_start;
            // the slots of this lane beyond its real ones are padding: zero them
            // for this block (the compute warps only write real slots; the
            // previous block is past its last barrier 2)
            for (int q = sc_cnt; q < CH; q++)
                awb_sts(z_s + zstep * (unsigned) q, 0.0);
            // column `sl` of the block's time-by-time matrix: this lane turns
            // the per-time sums F into R[sl] for the compute warps
            double tmc[TMAX];
            {
                const bool rl = (sl < T - 1) && nstatesg[b] > 0;
                const double *tmg = tmatrixg + (size_t) b * T * T + (rl ? sl : 0);
#pragma unroll
                for (int a = 0; a < TMAX; a++)
                    tmc[a] = (rl && a < T - 1) ? tmg[a * T] : 0.0;
            }

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
                    if ((1 << l) <= span) {
                        const double t = __shfl_up_sync(0xffffffffu, v, 1 << l);
                        v = fma(t, um[l], v);
                    }
                }
                if (sc_last)
                    awb_sts(Fp_s + 8u * (unsigned) sc_row, v);
                awb_bar_sync(3, AWB_FWD_FSCRIBES);
                if (i + 1 < blen && sl < T - 1) {
                    double2 f[(TMAX + 1) / 2];
#pragma unroll
                    for (int a = 0; a < (TMAX + 1) / 2; a++)
                        f[a] = awb_lds2(Fp_s + 16u * a);
                    double ra = 0.0, rb = 0.0, rc = 0.0, rd = 0.0;
#pragma unroll
                    for (int a = 0; a + 3 < TMAX; a += 4) {
                        ra = fma(tmc[a], f[a / 2].x, ra);
                        rb = fma(tmc[a + 1], f[a / 2].y, rb);
                        rc = fma(tmc[a + 2], f[a / 2 + 1].x, rc);
                        rd = fma(tmc[a + 3], f[a / 2 + 1].y, rd);
                    }
#pragma unroll
                    for (int a = TMAX - (TMAX % 4); a < TMAX; a++)
                        ra = fma(tmc[a], (a & 1) ? f[a / 2].y : f[a / 2].x, ra);
                    awb_sts(Rs_s + (site & 1) * RSTR + 8u * (unsigned) sl,
                            (ra + rb) + (rc + rd));
                }
                awb_bar_sync(2, NB2);
            }
        }
        __syncthreads();                                   // final barrier
        return;
    }

    // =====================================================================
    // compute warps
    // =====================================================================
    // Every instruction of this loop is issued by ~15 warps per site on 4
    // schedulers, so the loop is written for instruction count: 32-bit shared
    // addresses with ring offsets kept incrementally, a pointer ring for the
    // lagged table stores (idle lanes and not-to-be-stored columns point at a
    // per-thread sink).
    const unsigned char *__restrict__ kindg = chg.kind + g.site0;
    double *__restrict__ fwg = chg.fw - g.fwbias;
    const long long *__restrict__ row_offg = chg.row_off;
    const long long *__restrict__ fw_offg = chg.fw_off;
    const long long *__restrict__ trow_offg = chg.trow_off;
    const long long *__restrict__ ent_offg = chg.ent_off;
    const unsigned short *__restrict__ tmapg = chg.tmap;
    const unsigned short *__restrict__ ipermg = chg.iperm;
    const short *__restrict__ st_nodeg = chg.st_node;
    const signed char *__restrict__ st_timeg = chg.st_time;
    const signed char *__restrict__ st_ageg = chg.st_age;
    const double *__restrict__ inv_emitg = chg.inv_emit;
    const double *__restrict__ ling = chg.lin;
    const unsigned short *__restrict__ sw_startg = chg.sw_start;
    const unsigned short *__restrict__ sw_cntg = chg.sw_cnt;
    const unsigned short *__restrict__ sw_srcg = chg.sw_src;
    const double *__restrict__ sw_probg = chg.sw_prob;
    double *const sink = chg.sink + tid;

    // ---- my state in the current block
    int jj = 0, S = 0, S1 = 1;
    long long r0 = 0;
    bool active = false, live = false;     // live: active and S > 0
    unsigned zaddr = dummy_s, raddr = Rs_s;
    long long step = 0;                    // bytes from my entry of one row to the next
    double inv_e = 0.0, em = 0.0, Da = 0.0, hb = 0.0, A1 = 0.0, A2 = 0.0,
        A3 = 0.0, nrb = 1.0;
    double upm[NLEV], dnm[NLEV];

    auto load_compute = [&](int bb) {
        S = nstatesg[bb];
        S1 = S > 0 ? S : 1;
        r0 = row_offg[bb];
        const long long tr0 = trow_offg[bb];
        const int NSb = (int) (trow_offg[bb + 1] - tr0);
        unsigned short tj = 0xFFFF;
        if (tid < NSb)
            tj = tmapg[tr0 + tid];
        active = (tj != 0xFFFF);
        live = active && S > 0;
        jj = active ? (int) tj : 0;
        int atime = 0, cage = 0, node = -1, tpos = 0;
        inv_e = active ? 1.0 : 0.0;
        em = inv_e;
        if (live) {
            atime = st_timeg[r0 + jj];
            cage = st_ageg[r0 + jj];
            node = st_nodeg[r0 + jj];
            tpos = ipermg[r0 + jj];
            inv_e = inv_emitg[r0 + jj];
        }
        zaddr = active ? zT_s + 8u * (unsigned) tpos : dummy_s;
        raddr = Rs_s + 8u * (unsigned) atime;
        step = active ? 8ll * S1 : 0ll;
        const int key = live ? node : (0x10000 + lane);
        int seglane, segend;
        awb_lane_run(key, lane, seglane, segend);
#pragma unroll
        for (int l = 0; l < NLEV; l++) {
            upm[l] = (lane - (1 << l) >= seglane) ? 1.0 : 0.0;
            dnm[l] = (lane + (1 << l) <= segend) ? 1.0 : 0.0;
        }
        if (live) {
            const double *lin = ling + (size_t) bb * 7 * T;
            const double Bc = cage > 0 ? lin[2 * T + cage - 1] : 0.0;
            Da = lin[0 * T + atime];
            hb = lin[1 * T + atime] - Bc;
            A1 = lin[3 * T + atime];
            A2 = lin[4 * T + atime] - A1 * Bc;
            A3 = lin[5 * T + atime] - A1 * Bc;
            nrb = lin[6 * T + atime];
        } else {
            // idle lane (everything 0), or the size-1 state space (identity)
            Da = 0.0; hb = 0.0; A1 = 0.0; A2 = 0.0; A3 = 0.0;
            nrb = active ? 1.0 : 0.0;
        }
    };

    load_compute(bbeg);
    // columns site, site-1, site-2 of my state; w0/w1/w2 = where they go in the
    // table (the sink for the prior column, which is kept as the caller gave it);
    // nxt = my entry of the row of the next site (emission in, column out)
    double c = 0.0, c1 = 0.0, c2 = 0.0;
    double *w0 = sink, *w1 = sink, *w2 = sink;
    double *nxt = sink;
    if (active) {
        // first column: the prior (K1 or caller), kept as given.  With a
        // checkpointed table the prior is saved aside on the first pass (the
        // table memory is reused by the other segments) and put back for the
        // second; a later segment starts from its stored first column, which
        // goes into the table like any other column.
        double *row0 = fwg + fw_offg[bbeg] + jj;
        if (!chg.ckpt) {
            c = *row0;
        } else if (seg == 0) {
            if (pass == 0) {
                c = *row0;
                chg.ckptcol[jj] = c;
            } else {
                c = chg.ckptcol[jj];
                *row0 = c;
            }
        } else {
            c = chg.ckptcol[(size_t) seg * chg.maxS + jj];
            w0 = row0;
        }
        nxt = fwg + fw_offg[bbeg] + S1 + jj;
    }
    const unsigned char *kp = kindg + 2;
    unsigned kind_next = (n > 1) ? kindg[1] : 0;
    // ring offsets (bytes): iofs = ((site-2)&3)*8 into invS, rofs = (site&1)*RSTR
    // into Rs, sofs = (((site-2)/RS)&1)*8 into scaleS
    unsigned iofs = 16, rofs = 0, sofs = 0;

    // one site that is followed by a site of the same block
    auto site_step = [&](auto nlc) {
        constexpr int NL = decltype(nlc)::value;
        awb_sts(zaddr, c);
        awb_bar_sync(1, NB1);

        // branch scans in registers while the F-scribes sum the rows
        const double x0 = Da * c;
        const double y0 = x0 * hb;
        // exclusive sums PY = sum_{a<b} x_a (h[a] - Bc), Q = sum_{a>b} x_a: the
        // neighbour's term first, then an inclusive scan of those (an inclusive
        // scan minus the own term cancels: h grows like exp(cumulative
        // coalescent rate), so y0 can exceed the sum below it by many orders)
        double py = __shfl_up_sync(0xffffffffu, y0, 1) * upm[0];
        double q = __shfl_down_sync(0xffffffffu, x0, 1) * dnm[0];
#pragma unroll
        for (int l = 0; l < NL; l++) {
            const double ty = __shfl_up_sync(0xffffffffu, py, 1 << l);
            const double tq = __shfl_down_sync(0xffffffffu, q, 1 << l);
            py = fma(ty, upm[l], py);
            q = fma(tq, dnm[l], q);
        }
        const double PY = py, Q = q;
        const double W = fma(A1, PY, fma(x0, A2, fma(A3, Q, nrb * c)));
        // store column site-2 scaled by its 1/norm (norm warp, 2 steps ago)
        __stcs(w2, c2 * awb_lds(inv_s + iofs));      // streaming: L1 is for the block tables
        const unsigned kd = kind_next;
        kind_next = *kp++;
        double e = inv_e;
        if (kd != AWB_SITE_INVARIANT)               // uniform, ~3 % of the sites
            e = (kd == AWB_SITE_VARIANT) ? (live ? *nxt : inv_e) : em;
        awb_bar_sync(2, NB2);

        // R[atime] = sum_a tm[a][atime] * F[a], formed by the F-scribes
        // the lagged rescale factor every fourth site ((site & 3) == 2), the
        // constant 1.0 otherwise: branch-free (a branch around a volatile load
        // costs a branch resolution per warp and site)
        const bool resc = iofs == 0;
        const double sc = awb_lds(resc ? scale_s + sofs : dummy_s + 8u);
        double cn = (awb_lds(raddr + rofs) + W) * (e * sc);
        sofs ^= resc ? 8u : 0u;
        iofs = (iofs + 8u) & 24u;
        rofs ^= RSTR;
        c2 = c1; c1 = c; c = cn;
        w2 = w1; w1 = w0; w0 = nxt;
        nxt = (double *) ((char *) nxt + step);
    };

    for (int b = bbeg; b < bend; b++) {
        const int blen = (b == bextra) ? 1 : blocklensg[b];

        // ---------------- sites that are followed by a site of the same block
        // (ONE scan variant for all warps.  A variant per warp with just the
        // levels its longest branch needs was no faster per site and cost 0.6 us
        // per block: a warp changing variant at a block boundary runs cold
        // instructions, and the variants compete for the instruction cache.)
        for (int i = blen - 1; i > 0; i--)
            site_step(AwbInt<NLEV>());

        // ---------------- last site of the block
        {
            awb_sts(zaddr, c);
            awb_sts(active ? col_s + 8u * (unsigned) ((b & 1) * NS + jj) : dummy_s, c);
            awb_bar_sync(1, NB1);
            __stcs(w2, c2 * awb_lds(inv_s + iofs));      // streaming: L1 is for the block tables
            const unsigned kd = kind_next;
            kind_next = *kp++;
            awb_bar_sync(2, NB2);
            if (b == bend - 1)
                break;

            // breakpoint: gather through the switch CSR (sample_thread.cpp:345-389)
            double scale = 1.0;
            if (iofs == 0) {
                scale = awb_lds(scale_s + sofs);
                sofs ^= 8u;
            }
            iofs = (iofs + 8u) & 24u;
            rofs ^= RSTR;
            c2 = c1; c1 = c;
            w2 = w1; w1 = w0;
            load_compute(b + 1);
            double sum = 0.0;
            double e = em;
            if (active) {
                const int st = sw_startg[r0 + jj];
                const int cnt = sw_cntg[r0 + jj];
                const unsigned short *es = sw_srcg + ent_offg[b + 1] + st;
                const double *ep = sw_probg + ent_offg[b + 1] + st;
                // the old block's last column; the buffer alternates with the
                // block so a one-site block cannot overwrite it early
                const double *cold = colS + (b & 1) * NS;
                for (int x = 0; x < cnt; x++)
                    sum += cold[es[x]] * ep[x];
                w0 = fwg + fw_offg[b + 1] + jj;
                if (S > 0) {
                    e = inv_e;
                    if (kd == AWB_SITE_VARIANT)
                        e = *w0;
                    else if (kd == AWB_SITE_MASKED)
                        e = 1.0;
                }
                nxt = w0 + S1;
            } else {
                w0 = sink;
                nxt = sink;
            }
            c = sum * e * scale;
        }
    }

    // ---- the last two columns: their 1/norm is complete after the final barrier
    __syncthreads();
    *w1 = c1 * invS[(n - 2) & 3];
    *w0 = c * invS[(n - 1) & 3];
    // the column of the extra site is the first column of the next segment
    if (g.extra && pass == 0 && active)
        chg.ckptcol[(size_t) (seg + 1) * chg.maxS + jj] = c * invS[(n - 1) & 3];
}

#endif // AWB_FORWARD_FAST_CUH
