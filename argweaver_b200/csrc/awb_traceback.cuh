// awb_traceback.cuh -- K5: stochastic traceback of the threading HMM.
//
// Replaces stochastic_traceback / sample_hmm_posterior /
// sample_hmm_posterior_step (reference sample_thread.cpp:522-569, :470-503,
// :506-519) and sample() (common.h:272-290).
//
// Semantics (per site i, walking from the last site to the first):
//   A[j] = fw[i][j] * T[j -> k]        k = state already sampled at site i+1
//   total = sum_j A[j];  pick = rand()/RAND_MAX * total
//   path[i] = first j whose running sum reaches pick
// with T[j->k] from the reference's closed form (trans.h:89-128) and the draws
// taken from the caller's libc rand() integers in the reference's consumption
// order (one per site, last site first), so the sampled path equals the
// reference's for the same draws (up to last-bit ties of a running sum,
// probability ~1e-13 per site).
//
// Parallelisation: "run-length speculation".  The draw of a site does not
// depend on the path, and the path is piecewise constant (the no-recombination
// diagonal carries ~98 % of the mass), so for the current state k a whole WAVE
// of sites is tested in parallel -- one warp per site computes
//     pre = sum_{j<k} A[j],  total = sum_j A[j],  stays <=> pre < pick <= pre + A[k]
// (exactly sample()'s "first j with running sum >= pick" being k).  All sites
// of the wave before the first failure keep k; the failing site is sampled in
// full with a block-wide scan, k changes, the transition column is rebuilt and
// the walk continues from there.  One CTA per chain.
#ifndef AWB_TRACEBACK_CUH
#define AWB_TRACEBACK_CUH

#include "awb_common.cuh"

#define AWB_TB_SPW 2          // sites per warp in one speculative wave

struct AwbTbSmem {
    double wsum[32];
    int kmin;
    int kcur;
    int fail;
};

// block-wide sample() over one value per thread: returns the chosen index
__device__ inline int awb_block_sample(double A, bool valid, int S1, int r,
                                       int rand_max, AwbTbSmem *sm)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    double x = valid ? A : 0.0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d)
            x += t;
    }
    if (lane == 31)
        sm->wsum[warp] = x;
    if (tid == 0)
        sm->kmin = S1 - 1;          // common.h:289 fallback
    __syncthreads();
    double off = 0.0, total = 0.0;
    for (int w = 0; w < nwarps; w++) {
        const double s = sm->wsum[w];
        if (w < warp)
            off += s;
        total += s;
    }
    x += off;
    const double pick = (double) r / (double) rand_max * total;
    const bool hit = valid && (x >= pick);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (m != 0 && lane == 0)
        atomicMin(&sm->kmin, warp * 32 + __ffs(m) - 1);
    __syncthreads();
    const int k = sm->kmin;
    __syncthreads();                // kmin / wsum are reused by the next call
    return k;
}

__global__ void __launch_bounds__(1024)
awb_traceback_kernel(const AwbChain *chains, int rand_max)
{
    const AwbChain &ch = chains[blockIdx.x];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int NW = blockDim.x >> 5;
    const int T = ch.model.ntimes;
    const int V = ch.nnodes;
    const int n = ch.nsites;
    const int B = ch.ntrees;
    const bool internal = ch.internal != 0;
    const double *__restrict__ fwg = ch.fw;
    const int *__restrict__ randg = ch.rand_ints;
    int *__restrict__ pathg = ch.path;
    const double inv_rand_max = 1.0;   // (division kept literal below)
    (void) inv_rand_max;

    __shared__ AwbTbSmem sm;
    __shared__ double tvS[AWB_TM_NVEC * AWB_MAXT];
    __shared__ double transS[AWB_MAXS];
    __shared__ double swA[AWB_MAXS];          // switch step scratch (thread 0)
    __shared__ unsigned short swJ[AWB_MAXS];
    __shared__ int flagS[32 * AWB_TB_SPW];

    // draw used for site s: the reference consumes one rand() per sampled site,
    // last site first
    const int roff = (ch.last_state < 0) ? (n - 1) : (n - 2);
    int k;

    // ---- last column (sample_thread.cpp:534-539)
    {
        const int S1 = ch.nstates[B - 1] > 0 ? ch.nstates[B - 1] : 1;
        if (ch.last_state < 0) {
            const bool valid = tid < S1;
            const double A = valid ? fwg[ch.fw_off[B] - S1 + tid] : 0.0;
            k = awb_block_sample(A, valid, S1, randg[0], rand_max, &sm);
        } else {
            k = ch.last_state;
        }
        if (tid == 0)
            pathg[n - 1] = k;
    }

    for (int b = B - 1; b >= 0; b--) {
        const int S = ch.nstates[b];
        const int S1 = S > 0 ? S : 1;
        const long long r0 = ch.row_off[b];
        const int pos = ch.block_start[b];
        const int blen = ch.blocklens[b];
        const double *fw = fwg + ch.fw_off[b];
        const int *age = ch.ages + (size_t) b * V;
        const short *st_node = ch.st_node + r0;
        const signed char *st_time = ch.st_time + r0;

        // TransMatrix::get uses minage = age[subtree_root] (trans.h:67-74)
        int minage = 0;
        if (internal && S > 0)
            minage = age[ch.child0[(size_t) b * V + ch.root[b]]];
        __syncthreads();
        for (int x = tid; x < AWB_TM_NVEC * T; x += blockDim.x)
            tvS[x] = ch.tmvec[(size_t) b * AWB_TM_NVEC * T + x];
        __syncthreads();

        // ---- sample_hmm_posterior (sample_thread.cpp:470-503), speculative
        int i_hi = blen - 2;
        int trans_k = -1;
        while (i_hi >= 0) {
            if (trans_k != k) {
                // transition column into k (recomputed only when k changes)
                if (S > 0) {
                    const int node_k = st_node[k];
                    const int b_k = st_time[k];
                    const int c_k = age[node_k];
                    for (int j = tid; j < S; j += blockDim.x)
                        transS[j] = awb_get_time(tvS, T, st_time[j], b_k, c_k,
                                                 minage, st_node[j] == node_k);
                } else if (tid == 0) {
                    transS[0] = 1.0;
                }
                trans_k = k;
                __syncthreads();
            }

            // ---- one wave: NW * SPW sites tested for "stays in k"
#pragma unroll
            for (int u = 0; u < AWB_TB_SPW; u++) {
                const int w = warp * AWB_TB_SPW + u;       // w-th site of the wave
                const int i = i_hi - w;
                if (i >= 0) {
                    const double *row = fw + (long long) i * S1;
                    double tot0 = 0.0, tot1 = 0.0, pre0 = 0.0, pre1 = 0.0;
                    int j = lane;
                    for (; j + 32 < S1; j += 64) {
                        const double v0 = row[j] * transS[j];
                        const double v1 = row[j + 32] * transS[j + 32];
                        tot0 += v0;
                        tot1 += v1;
                        if (j < k) pre0 += v0;
                        if (j + 32 < k) pre1 += v1;
                    }
                    if (j < S1) {
                        const double v0 = row[j] * transS[j];
                        tot0 += v0;
                        if (j < k) pre0 += v0;
                    }
                    double tot = tot0 + tot1, pre = pre0 + pre1;
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
                        tot += __shfl_xor_sync(0xffffffffu, tot, d);
                        pre += __shfl_xor_sync(0xffffffffu, pre, d);
                    }
                    if (lane == 0) {
                        const double Ak = row[k] * transS[k];
                        const double pick = (double) randg[roff - (pos + i)] /
                            (double) rand_max * tot;
                        flagS[w] = (pre < pick) && (pre + Ak >= pick);
                    }
                } else if (lane == 0) {
                    flagS[w] = 1;
                }
            }
            __syncthreads();

            // first site of the wave (highest i) that leaves k
            int fail = -1;
            const int wave = NW * AWB_TB_SPW;
            for (int w = 0; w < wave; w++) {
                if (!flagS[w]) { fail = w; break; }
            }
            const int nkeep = (fail < 0) ? wave : fail;
            for (int w = tid; w < nkeep; w += blockDim.x) {
                const int i = i_hi - w;
                if (i >= 0)
                    pathg[pos + i] = k;
            }
            if (fail >= 0) {
                // sample the failing site in full
                const int i = i_hi - fail;
                const bool valid = tid < S1;
                const double A = valid ? fw[(long long) i * S1 + tid] * transS[tid] : 0.0;
                const int knew = awb_block_sample(A, valid, S1,
                                                  randg[roff - (pos + i)],
                                                  rand_max, &sm);
                if (tid == 0)
                    pathg[pos + i] = knew;
                k = knew;
                i_hi = i - 1;
            } else {
                i_hi -= wave;
            }
            __syncthreads();
        }

        // ---- sample_hmm_posterior_step through the switch matrix (:506-519)
        if (b > 0) {
            if (tid == 0) {
                const int n1 = ch.nstates[b - 1] > 0 ? ch.nstates[b - 1] : 1;
                const double *col1 = fwg + ch.fw_off[b] - n1;
                const int st = ch.sw_start[r0 + k];
                const int cn = ch.sw_cnt[r0 + k];
                const unsigned short *es = ch.sw_src + ch.ent_off[b] + st;
                const double *ep = ch.sw_prob + ch.ent_off[b] + st;
                // entries sorted by source index = the order sample() walks A[]
                for (int q = 0; q < cn; q++) {
                    const unsigned short jj = es[q];
                    const double val = col1[jj] * ep[q];
                    int w = q;
                    while (w > 0 && swJ[w - 1] > jj) {
                        swJ[w] = swJ[w - 1];
                        swA[w] = swA[w - 1];
                        w--;
                    }
                    swJ[w] = jj;
                    swA[w] = val;
                }
                double total = 0.0;
                for (int q = 0; q < cn; q++)
                    total += swA[q];
                const double pick = (double) randg[roff - (pos - 1)] /
                    (double) rand_max * total;
                // zero-weight states before the first entry win when pick == 0
                int kk = n1 - 1;
                double x = 0.0;
                if (0.0 >= pick && (cn == 0 || swJ[0] > 0)) {
                    kk = 0;
                } else {
                    for (int q = 0; q < cn; q++) {
                        x += swA[q];
                        if (x >= pick) { kk = swJ[q]; break; }
                    }
                }
                sm.kcur = kk;
                pathg[pos - 1] = kk;
            }
            __syncthreads();
            k = sm.kcur;
            __syncthreads();
        }
    }
}

#endif // AWB_TRACEBACK_CUH
