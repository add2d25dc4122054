// awb_forward.cuh -- K4: the forward recursion of the threading HMM.
//
// Replaces arghmm_forward_alg / arghmm_forward_block / arghmm_forward_switch
// (reference sample_thread.cpp:394-460, :186-296, :345-389).
//
// One persistent CTA per chain walks the sites sequentially (column i needs
// column i-1); the state axis is the parallel axis: ONE THREAD PER STATE, in
// TIME-MAJOR order.  The SMC transition inside a block is
//
//   col2[k] = emit[k] * ( R[b_k] + W_k ),
//   R[b]  = sum_a tmatrix[a][b] * F[a],      F[a] = sum_{j: time_j = a} col[j]
//   W_k   = sum_{j on branch(k)} band_k[j] * col[j]        (same-branch band)
//
// i.e. rank-(T-1) through the per-time group sums F plus a banded per-branch
// correction.  Per site:
//   1. publish: each thread stores its value into the node-major column copy
//      in shared memory and the lanes of one time row reduce their values with
//      a segmented warp-shuffle scan (time-major order makes a row a contiguous
//      lane segment); the last lane of each segment stores a partial sum.
//   2. __syncthreads.  Warp 0 combines the partials into F[a], forms the column
//      norm and the T-vector R (tmatrix lives in shared memory); all other warps
//      meanwhile gather their band term W_k from the shared column copy.
//   3. __syncthreads.  Every thread forms its new value; the PREVIOUS column is
//      streamed to HBM normalised (coalesced, in the reference's state order).
// At a breakpoint the new block's threads gather from the old column through
// the CSR switch lists built by awb_switch_setup.
//
// Columns are carried unnormalised by one step and rescaled by 1/norm when
// used/stored, which is algebraically the reference's per-column normalisation
// (sample_thread.cpp:292-294); logZ = sum_i log(norm_i) is accumulated on the
// side (the reference drops it).
#ifndef AWB_FORWARD_CUH
#define AWB_FORWARD_CUH

#include "awb_common.cuh"

struct AwbFwdSmem {
    double *colS;      // [2][NPT*NS]
    double *part;      // [NP]
    double *Fs;        // [AWB_MAXT]
    double *Rs;        // [AWB_MAXT]
    double *tmS;       // [T*T]
    double *bandS;     // [bandcap] (0: the band is read from global memory)
    double *scal;      // [2] inv, norm
    unsigned short *pstartS; // [AWB_MAXT+1]
};

// ncol = NPT * NS state slots; bandcap = 0 leaves the band in global memory
__host__ __device__ inline size_t awb_fwd_smem_bytes(int ncol, int T, int bandcap)
{
    size_t nd = 2 * (size_t) ncol + (AWB_MAXT + ncol / 32 + 2) + AWB_MAXT + AWB_MAXT +
        (size_t) T * T + (size_t) bandcap + 2;
    return nd * sizeof(double) + (AWB_MAXT + 1 + 3) * sizeof(unsigned short);
}

// NPT = states per thread: thread t owns the time-major positions t, t+NS, ...
template <int NPT>
__global__ void __launch_bounds__(1024)
awb_forward_kernel(const AwbChain *chains, int bandcap)
{
    const AwbChain &ch = chains[blockIdx.x];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int NS = blockDim.x;
    const int NC = NPT * NS;
    const int T = ch.model.ntimes;
    const int n = ch.nsites;
    const int B = ch.ntrees;

    extern __shared__ double smem_d[];
    AwbFwdSmem sm;
    sm.colS = smem_d;
    sm.part = sm.colS + 2 * NC;
    sm.Fs = sm.part + (AWB_MAXT + NC / 32 + 2);
    sm.Rs = sm.Fs + AWB_MAXT;
    sm.tmS = sm.Rs + AWB_MAXT;
    sm.bandS = sm.tmS + T * T;
    sm.scal = sm.bandS + bandcap;
    sm.pstartS = (unsigned short *) (sm.scal + 2);

    // ---- per-thread description of "my" states in the current block
    int b = 0, S = 0, S1 = 1, blen = 0, ib = 0;
    long long r0 = 0, fwbase = 0;
    int j[NPT], atime[NPT], myslot[NPT], seglane[NPT], j1[NPT], len[NPT], boff[NPT];
    bool active[NPT], seg_last[NPT];
    double inv_e[NPT];
    const double *bandp = sm.bandS;

    auto load_block = [&](int bb) {
        S = ch.nstates[bb];
        S1 = S > 0 ? S : 1;
        r0 = ch.row_off[bb];
        fwbase = ch.fw_off[bb];
        blen = ch.blocklens[bb];
#pragma unroll
        for (int q = 0; q < NPT; q++) {
            const int p = tid + q * NS;
            active[q] = p < S1;
            j[q] = 0; atime[q] = 0; myslot[q] = 0; j1[q] = 0; len[q] = 0; boff[q] = 0;
            inv_e[q] = 1.0;
            if (active[q] && S > 0) {
                j[q] = ch.perm[r0 + p];
                atime[q] = ch.st_time[r0 + j[q]];
                myslot[q] = ch.pslot[r0 + p];
                j1[q] = ch.band_j1[r0 + j[q]];
                len[q] = ch.band_len[r0 + j[q]];
                boff[q] = ch.band_boff[r0 + j[q]];
                inv_e[q] = ch.inv_emit[r0 + j[q]];
            }
            const int key = active[q] ? myslot[q] : (0x10000 + lane);
            const unsigned m = __match_any_sync(0xffffffffu, key);
            seglane[q] = __ffs(m) - 1;
            seg_last[q] = (lane == 31 - __clz(m));
        }
        if (S > 0) {
            const double *tmg = ch.tmatrix + (size_t) bb * T * T;
            for (int x = tid; x < T * T; x += NS)
                sm.tmS[x] = tmg[x];
            const double *bg = ch.band + ch.band_off[bb];
            if (bandcap > 0) {
                const int bl = (int) (ch.band_off[bb + 1] - ch.band_off[bb]);
                for (int x = tid; x < bl; x += NS)
                    sm.bandS[x] = bg[x];
            } else {
                bandp = bg;
            }
            const unsigned short *pg = ch.pstart + (size_t) bb * (T + 1);
            for (int x = tid; x <= T; x += NS)
                sm.pstartS[x] = pg[x];
        }
    };

    load_block(0);
    double c[NPT];
#pragma unroll
    for (int q = 0; q < NPT; q++)
        c[q] = active[q] ? ch.fw[fwbase + j[q]] : 0.0;   // prior column (K1 or caller)
    double logz = 0.0;
    int bad_site = -1;
    int buf = 0;
    unsigned char kind_next = (n > 1) ? ch.kind[1] : 0;

    for (int site = 0; site < n; site++) {
        double *col = sm.colS + buf * NC;

        // ---- 1. publish
#pragma unroll
        for (int q = 0; q < NPT; q++) {
            double v = active[q] ? c[q] : 0.0;
            if (active[q])
                col[j[q]] = c[q];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, v, d);
                if (lane - d >= seglane[q])
                    v += t;
            }
            if (active[q] && seg_last[q])
                sm.part[myslot[q]] = v;
        }
        __syncthreads();

        // ---- 2. band term (all warps) and the T-vector R (warp 0)
        double W[NPT];
#pragma unroll
        for (int q = 0; q < NPT; q++) {
            W[q] = 0.0;
            if (active[q] && S > 0) {
                const double *cf = bandp + boff[q];
                const double *cj = col + j1[q];
                for (int x = 0; x < len[q]; x++)
                    W[q] += cf[x] * cj[x];
            }
        }
        if (warp == 0) {
            if (S > 0) {
                for (int a = lane; a < T - 1; a += 32) {
                    double Fa = 0.0;
                    const int p1 = sm.pstartS[a + 1];
                    for (int p = sm.pstartS[a]; p < p1; p++)
                        Fa += sm.part[p];
                    sm.Fs[a] = Fa;
                }
                __syncwarp();
                double nrm = 0.0;
                for (int a = 0; a < T - 1; a++)
                    nrm += sm.Fs[a];
                for (int bb = lane; bb < T - 1; bb += 32) {
                    double r = 0.0;
                    for (int a = 0; a < T - 1; a++)
                        r += sm.tmS[a * T + bb] * sm.Fs[a];
                    sm.Rs[bb] = r;
                }
                if (lane == 0) {
                    sm.scal[0] = 1.0 / nrm;
                    sm.scal[1] = nrm;
                }
                // per-time sums of the column as stored (used by the traceback)
                for (int a = lane; a < T - 1; a += 32)
                    ch.fsum[(size_t) site * (T - 1) + a] =
                        sm.Fs[a] * (site == 0 ? 1.0 : 1.0 / nrm);
            } else if (lane == 0) {
                const double nrm = sm.part[0];
                sm.scal[0] = 1.0 / nrm;
                sm.scal[1] = nrm;
            }
        }
        __syncthreads();

        // ---- 3. stream the finished column, form the next one
        const double inv = sm.scal[0];
        if (site > 0)
            for (int x = tid; x < S1; x += NS)
                ch.fw[fwbase + (long long) ib * S1 + x] = col[x] * inv;
        if (tid == NS - 32) {
            const double nrm = sm.scal[1];
            logz += log(nrm);
            if (!(nrm > 0.0) && bad_site < 0)
                bad_site = site;
        }
        if (site == n - 1)
            break;

        const unsigned char kd = kind_next;
        if (site + 2 < n)
            kind_next = ch.kind[site + 2];
        ib++;
        if (ib == blen) {
            // breakpoint: gather through the switch CSR (sample_thread.cpp:345-389)
            b++;
            ib = 0;
            load_block(b);
#pragma unroll
            for (int q = 0; q < NPT; q++) {
                double sum = 0.0;
                if (active[q]) {
                    const int st = ch.sw_start[r0 + j[q]];
                    const int cn = ch.sw_cnt[r0 + j[q]];
                    const unsigned short *es = ch.sw_src + ch.ent_off[b] + st;
                    const double *ep = ch.sw_prob + ch.ent_off[b] + st;
                    for (int x = 0; x < cn; x++)
                        sum += col[es[x]] * ep[x];
                }
                double e = 1.0;
                if (active[q] && S > 0) {
                    if (kd == AWB_SITE_VARIANT)
                        e = ch.fw[fwbase + j[q]];
                    else if (kd == AWB_SITE_INVARIANT)
                        e = inv_e[q];
                }
                c[q] = sum * e * inv;
            }
        } else {
#pragma unroll
            for (int q = 0; q < NPT; q++) {
                double e = 1.0;
                if (active[q] && S > 0) {
                    if (kd == AWB_SITE_VARIANT)
                        e = ch.fw[fwbase + (long long) ib * S + j[q]];
                    else if (kd == AWB_SITE_INVARIANT)
                        e = inv_e[q];
                    c[q] = (sm.Rs[atime[q]] + W[q]) * e * inv;
                } else {
                    c[q] = c[q] * inv;
                }
            }
        }
        buf ^= 1;
    }

    if (tid == NS - 32) {
        ch.logz[0] = logz;
        ch.status[0] = bad_site;
    }
    (void) B;
}

#endif // AWB_FORWARD_CUH
