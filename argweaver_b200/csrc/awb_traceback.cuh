// awb_traceback.cuh -- K5: stochastic traceback of the threading HMM.
//
// Replaces stochastic_traceback / sample_hmm_posterior /
// sample_hmm_posterior_step (reference sample_thread.cpp:522-569, :470-503,
// :506-519) and sample() (common.h:272-290).
//
// One CTA per chain, one thread per state in the reference's (node-major) state
// order, walking the sites from last to first.  Per site:
//   A[j] = fw[i][j] * T[j -> k]        k = state already sampled at site i+1
//   total = sum_j A[j];  pick = rand()/RAND_MAX * total
//   path[i] = first j whose running sum reaches pick
// T[j->k] is evaluated with the reference's own closed form (trans.h:89-128)
// and only recomputed when k changes.  The running sums are a block-wide
// inclusive scan (warp shuffles + one shared-memory hop); the draws are the
// caller's libc rand() integers in the reference's consumption order, so the
// sampled path equals the reference's for the same draws (up to last-bit ties
// of the running sum, probability ~1e-13 per site).
#ifndef AWB_TRACEBACK_CUH
#define AWB_TRACEBACK_CUH

#include "awb_common.cuh"

struct AwbTbSmem {
    double wsum[32];
    double total;
    int kmin;
    int kcur;
};

// block-wide sample(): returns the chosen index to every thread
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
    const int T = ch.model.ntimes;
    const int V = ch.nnodes;
    const int n = ch.nsites;
    const int B = ch.ntrees;
    const bool internal = ch.internal != 0;

    __shared__ AwbTbSmem sm;
    __shared__ double tvS[AWB_TM_NVEC * AWB_MAXT];
    __shared__ double swA[AWB_MAXS];          // switch step scratch (thread 0)
    __shared__ unsigned short swJ[AWB_MAXS];

    int r = 0;                                 // index of the next rand() draw
    int k;

    // ---- last column (sample_thread.cpp:534-539)
    {
        const int S1 = ch.nstates[B - 1] > 0 ? ch.nstates[B - 1] : 1;
        if (ch.last_state < 0) {
            const bool valid = tid < S1;
            const double A = valid ? ch.fw[ch.fw_off[B] - S1 + tid] : 0.0;
            k = awb_block_sample(A, valid, S1, ch.rand_ints[r++], rand_max, &sm);
        } else {
            k = ch.last_state;
        }
        if (tid == 0)
            ch.path[n - 1] = k;
    }

    for (int b = B - 1; b >= 0; b--) {
        const int S = ch.nstates[b];
        const int S1 = S > 0 ? S : 1;
        const long long r0 = ch.row_off[b];
        const int pos = ch.block_start[b];
        const int blen = ch.blocklens[b];
        const double *fw = ch.fw + ch.fw_off[b];
        const int *age = ch.ages + (size_t) b * V;
        const bool valid = tid < S1;

        // TransMatrix::get uses minage = age[subtree_root] (trans.h:67-74)
        int minage = 0;
        if (internal && S > 0)
            minage = age[ch.child0[(size_t) b * V + ch.root[b]]];
        __syncthreads();
        for (int x = tid; x < AWB_TM_NVEC * T; x += blockDim.x)
            tvS[x] = ch.tmvec[(size_t) b * AWB_TM_NVEC * T + x];
        __syncthreads();

        int node_j = -1, a_j = 0;
        if (valid && S > 0) {
            node_j = ch.st_node[r0 + tid];
            a_j = ch.st_time[r0 + tid];
        }

        // ---- sample_hmm_posterior (sample_thread.cpp:470-503)
        int last_k = -1;
        double trans = 1.0;
        double fnext = (valid && blen >= 2) ?
            fw[(long long) (blen - 2) * S1 + tid] : 0.0;
        for (int i = blen - 2; i >= 0; i--) {
            const double f = fnext;
            if (i > 0 && valid)
                fnext = fw[(long long) (i - 1) * S1 + tid];
            if (k != last_k) {
                if (S > 0 && valid) {
                    const int node_k = ch.st_node[r0 + k];
                    const int b_k = ch.st_time[r0 + k];
                    trans = awb_get_time(tvS, T, a_j, b_k, age[node_k], minage,
                                         node_j == node_k);
                } else {
                    trans = 1.0;
                }
                last_k = k;
            }
            k = awb_block_sample(f * trans, valid, S1, ch.rand_ints[r++],
                                 rand_max, &sm);
            if (tid == 0)
                ch.path[pos + i] = k;
        }

        // ---- sample_hmm_posterior_step through the switch matrix (:506-519)
        if (b > 0) {
            if (tid == 0) {
                const int n1 = ch.nstates[b - 1] > 0 ? ch.nstates[b - 1] : 1;
                const double *col1 = ch.fw + ch.fw_off[b] - n1;
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
                const double pick = (double) ch.rand_ints[r] / (double) rand_max * total;
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
                ch.path[pos - 1] = kk;
            }
            __syncthreads();
            k = sm.kcur;
            r++;
            __syncthreads();
        }
    }
}

#endif // AWB_TRACEBACK_CUH
