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
//   W_k = E e2[b] * ( P1 - Bc*P0 ) + x_b * A2 + A3 * Q + norecombs[b]*col[k]
//   P0 = sum_{a<b} x_a   P1 = sum_{a<b} x_a h[a]   Q = sum_{a>b} x_a
//
// (Bc, A2, A3 per-state constants, see load_thread).  P0, P1, Q are exclusive
// prefix / suffix sums along one branch; the two forms agree to ~4e-15
// relative (checked against the literal band on random columns).
//
// Mapping to the SM (numbers measured on B200, scripts/microbench.cu:
// dependent DFMA 8.4 cycles, 64-bit SHFL+DADD 35, LDS 33, __syncthreads 33,
// DDIV 130):
//
//   compute warps  one thread per state in NODE-MAJOR order, packed so that the
//                  states of a branch never straddle a warp: P0/P1/Q are
//                  segmented warp-shuffle scans in registers (no shared
//                  memory).  The column of the (T x T) time matrix a state
//                  needs sits in registers (tmc[]), loaded once per block.
//   scribe warps   two warps that own no state.  Between the two barriers of a
//                  site step their 64 lanes turn the column (staged in shared
//                  memory in TIME-MAJOR order) into the per-time sums F[a];
//                  after the second barrier they form the column norm, 1/norm,
//                  logZ and the rescale factor -- the division and log() never
//                  sit on the critical path.
//
//   step(site):  STS value (time-major slot) ; branch scans -> W
//                barrier 1   scribes: F[a]   | compute: store column site-1 to
//                                              HBM scaled by its 1/norm,
//                                              fetch the next emission
//                barrier 2   R = sum_a tmc[a]*F[a] (10 LDS.128 + 19 DFMA)
//                            col(site+1) = (R + W) * emission
//
// The forward table is written once, 8 B per site*state, in the reference's
// state order.  Columns are carried unnormalised; a factor published for
// column s is applied when column s+2 is formed (every AWB_FWD_RS sites), which
// bounds the magnitude by the product of at most RS+1 one-step norms.
#ifndef AWB_FORWARD_FAST_CUH
#define AWB_FORWARD_FAST_CUH

#include "awb_common.cuh"

#define AWB_FWD_RS 4          // rescale period (sites)
#define AWB_FWD_SCRIBES 64    // scribe lanes (2 warps)

// shared memory (doubles): zT[NS] | colS[2][NS] | Fs[2][TMAX+2] | scaleS[2] | invS[2]
__host__ __device__ inline size_t awb_fwd_fast_smem_bytes(int NS, int TMAX)
{
    return (3 * (size_t) NS + 2 * (size_t) (TMAX + 2) + 4) * sizeof(double);
}

template <int TMAX, int MAXTHREADS>
__global__ void __launch_bounds__(MAXTHREADS, 1)
awb_forward_fast_kernel(const AwbChain *chains, int maxd)
{
    const AwbChain &chg = chains[blockIdx.x];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int NS = blockDim.x - AWB_FWD_SCRIBES;     // compute threads
    const bool scribe = tid >= NS;
    const int sl = tid - NS;                         // scribe lane 0..63

    // loop-invariant fields, hoisted out of the (aliasing) struct
    const int T = chg.model.ntimes;
    const int n = chg.nsites;
    const unsigned char *__restrict__ kindg = chg.kind;
    double *__restrict__ fwg = chg.fw;
    const int *__restrict__ nstatesg = chg.nstates;
    const int *__restrict__ blocklensg = chg.blocklens;
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
    const double *__restrict__ tmatrixg = chg.tmatrix;
    const unsigned short *__restrict__ sw_startg = chg.sw_start;
    const unsigned short *__restrict__ sw_cntg = chg.sw_cnt;
    const unsigned short *__restrict__ sw_srcg = chg.sw_src;
    const double *__restrict__ sw_probg = chg.sw_prob;
    const unsigned short *__restrict__ sc_startg = chg.sc_start;
    const unsigned short *__restrict__ sc_cntg = chg.sc_cnt;
    const unsigned char *__restrict__ sc_rowg = chg.sc_row;

    extern __shared__ double smem_f[];
    double *zT = smem_f;                       // [NS] column, time-major slots
    double *colS = zT + NS;                    // [2][NS] last column of a block in
                                               //   state order, by block parity
    double *FsS = colS + 2 * NS;               // [2][TMAX+2] per-time sums
    double *scaleS = FsS + 2 * (TMAX + 2);     // [2] rescale factors
    double *invS = scaleS + 2;                 // [2] 1/norm of the last columns

    // ---- block bookkeeping (uniform across the CTA)
    int b = 0, ib = 0, blen = 0, S = 0, S1 = 1;
    long long r0 = 0, fwbase = 0;

    // ---- per-thread description (compute: my state; scribe: my chunk)
    int jj = 0, tpos = 0, seglane = 0, segend = 0;
    bool active = false;
    double inv_e = 1.0, Da = 0.0, ha = 0.0, Bc = 0.0, A1 = 0.0, A2 = 0.0,
        A3 = 0.0, nrb = 1.0;
    double tmc[TMAX];
    int sc_start = 0, sc_cnt = 0, sc_row = 255;
    bool sc_last = false;

    auto load_block_scalars = [&](int bb) {
        S = nstatesg[bb];
        S1 = S > 0 ? S : 1;
        r0 = row_offg[bb];
        fwbase = fw_offg[bb];
        blen = blocklensg[bb];
    };

    auto load_thread = [&](int bb) {
        if (!scribe) {
            const long long tr0 = trow_offg[bb];
            const int NSb = (int) (trow_offg[bb + 1] - tr0);
            unsigned short tj = 0xFFFF;
            if (tid < NSb)
                tj = tmapg[tr0 + tid];
            active = (tj != 0xFFFF);
            jj = active ? (int) tj : 0;
            int atime = 0, cage = 0, node = -1;
            tpos = 0;
            inv_e = 1.0;
            if (active && S > 0) {
                atime = st_timeg[r0 + jj];
                cage = st_ageg[r0 + jj];
                node = st_nodeg[r0 + jj];
                tpos = ipermg[r0 + jj];
                inv_e = inv_emitg[r0 + jj];
            }
            const int key = (active && S > 0) ? node : (0x10000 + lane);
            const unsigned m = __match_any_sync(0xffffffffu, key);
            seglane = __ffs(m) - 1;
            segend = 31 - __clz(m);
            if (active && S > 0) {
                const double *lin = ling + (size_t) bb * 7 * T;
                Da = lin[0 * T + atime];
                ha = lin[1 * T + atime];
                Bc = cage > 0 ? lin[2 * T + cage - 1] : 0.0;
                A1 = lin[3 * T + atime];
                A2 = lin[4 * T + atime] - A1 * Bc;
                A3 = lin[5 * T + atime] - A1 * Bc;
                nrb = lin[6 * T + atime];
            } else {
                // idle lane, or the size-1 state space (identity transition)
                Da = 0.0; ha = 0.0; Bc = 0.0; A1 = 0.0; A2 = 0.0; A3 = 0.0;
                nrb = 1.0;
            }
            const double *tmg = tmatrixg + (size_t) bb * T * T + atime;
#pragma unroll
            for (int a = 0; a < TMAX; a++)
                tmc[a] = (active && S > 0 && a < T - 1) ? tmg[a * T] : 0.0;
        } else {
            sc_start = sc_startg[(size_t) bb * 64 + sl];
            sc_cnt = sc_cntg[(size_t) bb * 64 + sl];
            sc_row = sc_rowg[(size_t) bb * 64 + sl];
            const int key = (sc_row != 255) ? sc_row : (0x100 + lane);
            const unsigned m = __match_any_sync(0xffffffffu, key);
            seglane = __ffs(m) - 1;
            segend = 31 - __clz(m);
            sc_last = (lane == segend) && (sc_row != 255);
        }
    };

    for (int x = tid; x < 3 * NS + 2 * (TMAX + 2) + 4; x += blockDim.x)
        smem_f[x] = (x < 3 * NS + 2 * (TMAX + 2)) ? 0.0 : 1.0;
    __syncthreads();

    load_block_scalars(0);
    load_thread(0);
    double c = 0.0, c_prev = 0.0;
    double *fw_cur = nullptr, *fw_prev = nullptr;
    if (!scribe && active)
        c = fwg[fwbase + jj];                      // prior column (K1 or caller)
    double lacc = 0.0;          // scribe: sum of log(norm) of the applied rescales
    int bad_site = -1;
    int buf = 0;
    unsigned char kind_next = (n > 1) ? kindg[1] : 0;

    for (int site = 0; site < n; site++) {
        double *Fs = FsS + buf * (TMAX + 2);
        const bool last = (site == n - 1);
        const bool sw = !last && (ib + 1 == blen);
        double W = 0.0;

        // ---- publish column `site`; branch scans (registers only)
        if (!scribe) {
            if (active) {
                zT[tpos] = c;
                if (sw)
                    colS[(b & 1) * NS + jj] = c;
            }
            const double x0 = Da * c;
            const double x1 = x0 * ha;
            double p0 = x0, p1 = x1, q = x0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                if (d < maxd) {
                    const double t0 = __shfl_up_sync(0xffffffffu, p0, d);
                    const double t1 = __shfl_up_sync(0xffffffffu, p1, d);
                    const double tq = __shfl_down_sync(0xffffffffu, q, d);
                    if (lane - d >= seglane) {
                        p0 += t0;
                        p1 += t1;
                    }
                    if (lane + d <= segend)
                        q += tq;
                }
            }
            const double P0 = p0 - x0, P1 = p1 - x1, Q = q - x0;
            W = fma(A1, fma(-Bc, P0, P1), fma(x0, A2, fma(A3, Q, nrb * c)));
        }
        __syncthreads();                                   // barrier 1

        const unsigned char kd = kind_next;
        if (site + 2 < n)
            kind_next = kindg[site + 2];
        double e = 1.0;
        if (scribe) {
            // ---- per-time sums F[a]: each lane sums its chunk of one row, the
            //      lanes of a row combine with a segmented shuffle scan
            double v0 = 0.0, v1 = 0.0;
            const double *z = zT + sc_start;
            int i = 0;
            for (; i + 2 <= sc_cnt; i += 2) {
                v0 += z[i];
                v1 += z[i + 1];
            }
            if (i < sc_cnt)
                v0 += z[i];
            double v = v0 + v1;
            const int span = __reduce_max_sync(0xffffffffu, segend - seglane);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                if (d <= span) {
                    const double t = __shfl_up_sync(0xffffffffu, v, d);
                    if (lane - d >= seglane)
                        v += t;
                }
            }
            if (sc_last)
                Fs[sc_row] = v;
        } else {
            // ---- store column site-1 (scaled by its 1/norm, published by the
            //      scribes one step ago); fetch the next site's emission
            if (fw_prev)
                *fw_prev = c_prev * invS[(site - 1) & 1];
            if (!last && !sw && active && S > 0) {
                if (kd == AWB_SITE_VARIANT)
                    e = fwg[fwbase + (long long) (ib + 1) * S + jj];
                else if (kd == AWB_SITE_INVARIANT)
                    e = inv_e;
            }
        }
        __syncthreads();                                   // barrier 2

        if (scribe && sl < 32) {
            // ---- norm of column `site`, 1/norm, rescale factor, logZ
            double x = 0.0;
            for (int a = lane; a < T - 1; a += 32)
                x += Fs[a];
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1)
                x += __shfl_xor_sync(0xffffffffu, x, d);
            const double nrm = x;
            const double inv = 1.0 / nrm;
            if (lane == 0)
                invS[site & 1] = inv;
            if (!(nrm > 0.0) && bad_site < 0)
                bad_site = site;
            if (site % AWB_FWD_RS == 0) {
                // this factor is applied when column site+2 is formed
                if (lane == 0)
                    scaleS[(site / AWB_FWD_RS) & 1] = inv;
                if (site + 2 <= n - 1)
                    lacc += log(nrm);
            }
            if (last && lane == 0) {
                chg.logz[0] = log(nrm) + lacc;
                chg.status[0] = bad_site;
            }
            if (sw) {
                // rows that are empty in the next block must read as zero
                for (int a = lane; a < 2 * (TMAX + 2); a += 32)
                    FsS[a] = 0.0;
            }
        }
        if (last)
            break;

        // ---- form column site+1
        c_prev = c;
        fw_prev = fw_cur;
        if (site == 0)
            fw_prev = nullptr;             // the prior column is stored as is
        ib++;
        if (sw) {
            b++;
            ib = 0;
            load_block_scalars(b);
            load_thread(b);
        }
        if (!scribe) {
            double scale = 1.0;
            if (site >= 1 && (site - 1) % AWB_FWD_RS == 0)
                scale = scaleS[((site - 1) / AWB_FWD_RS) & 1];
            if (sw) {
                // breakpoint: gather through the switch CSR (sample_thread.cpp:345-389)
                double sum = 0.0;
                if (active) {
                    const int st = sw_startg[r0 + jj];
                    const int cn = sw_cntg[r0 + jj];
                    const unsigned short *es = sw_srcg + ent_offg[b] + st;
                    const double *ep = sw_probg + ent_offg[b] + st;
                    // the old block's last column; the buffer alternates with the
                    // block so a one-site block cannot overwrite it early
                    const double *cold = colS + ((b - 1) & 1) * NS;
                    for (int q = 0; q < cn; q++)
                        sum += cold[es[q]] * ep[q];
#ifdef AWB_DEBUG_SWITCH
                    if (site == AWB_DEBUG_SWITCH && (jj == 91 || jj == 33)) {
                        printf("dbg site %d tid %d jj %d cn %d sum %.6e e %.6e scale %.6e\n",
                               site, tid, jj, cn, sum, e, scale);
                        for (int q = 0; q < cn; q++)
                            printf("   jj %d src %d col %.6e prob %.6e\n", jj,
                                   (int) es[q], cold[es[q]], ep[q]);
                    }
#endif
                }
                e = 1.0;
                if (active && S > 0) {
                    if (kd == AWB_SITE_VARIANT)
                        e = fwg[fwbase + jj];
                    else if (kd == AWB_SITE_INVARIANT)
                        e = inv_e;
                }
                c = sum * e * scale;
            } else {
                // R_k = sum_a tmc[a] * F[a]
                const double2 *F2 = reinterpret_cast<const double2 *>(Fs);
                double2 f[(TMAX + 1) / 2];
#pragma unroll
                for (int a = 0; a < (TMAX + 1) / 2; a++)
                    f[a] = F2[a];
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
                c = active ? ((ra + rb) + (rc + rd) + W) * e * scale : 0.0;
            }
            fw_cur = active ? fwg + fwbase + (long long) ib * S1 + jj : nullptr;
        }
        buf ^= 1;
    }

    // ---- the last column: its 1/norm is published after the final barrier 2
    __syncthreads();
    if (!scribe && fw_cur && n > 1)
        *fw_cur = c * invS[(n - 1) & 1];
}

#endif // AWB_FORWARD_FAST_CUH
