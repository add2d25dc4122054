// awb_totalprob.cu -- log-likelihood and log-prior of a complete ARG, the two
// numbers arg-sample writes to its .stats file every iteration.
//
// Replaces
//   calc_arg_likelihood   total_prob.cpp:19-42   (likelihood_tree emit.cpp:399-449)
//   calc_arg_prior        total_prob.cpp:262-299 (calc_spr_prob :210-258,
//                         calc_coal_rates_full_tree :196-207)
// and the C exports arghmm_likelihood / arghmm_prior_prob / arghmm_joint_prob
// (total_prob.cpp:316-377; bound in awb_compat.cu).
//
// Kernels (one launch each over all blocks of the ARG):
//   awb_tp_kind_kernel   thread per site   invariant / variant over ALL sequences
//   awb_tp_block_kernel  4 lanes per block Felsenstein pruning of the block's
//                        variant sites and of its first invariant site, one lane
//                        per base, children before parents in an explicit
//                        post-order made by lane 0; the lineage counts and the
//                        SPR probability of the block's right-hand breakpoint
//   awb_tp_reduce_kernel one CTA           fixed-order sum of the per-block terms
// The per-site likelihoods follow the reference's arithmetic term by term
// (same post-order independence: a node's row is a product of two 4-term sums
// in base order), including its quirk that every invariant site of a block
// takes the likelihood of the block's FIRST invariant site, whatever its base
// (emit.cpp:431-442): an all-'N' first column makes the others count as 1.
//
// There is no CPU fallback.

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "argweaver_b200.h"
#include "awb_common.cuh"

extern int awb_fail_msg(const std::string &msg);      // awb_api.cu

#define TP_OK(call)                                                           \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess)                                                \
            return awb_fail_msg(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct AwbTp {
    int T, V, B, n, nseqs, seqlen, start_coord, nleaves;
    double rho, mu, mintime;
    double times[AWB_MAXT], time_steps[AWB_MAXT], popsizes[AWB_MAXT];
    double coal_time_steps[2 * AWB_MAXT];
    const int *ptrees, *ages, *sprs, *blocklens, *block_start, *rowidx;
    const unsigned char *seqs;
    unsigned char *kind;
    double *blk_lik, *blk_prior, *out;      // [B], [B], [2]
    short *scratch_i;                       // [groups][3V] child0, child1, order
    double *scratch_d;                      // [groups][6V] inner[4V], mut, nomut
};

__device__ __forceinline__ int tp_base(unsigned char c)
{
    switch (c) {                            // seq.cpp:15-43 (dna2int)
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    }
    return -1;
}

// emit.cpp:16-27 over the rows the trees' leaves map to
__global__ void awb_tp_kind_kernel(AwbTp P)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n;
         i += gridDim.x * blockDim.x) {
        const size_t col = (size_t) P.start_coord + i;
        const unsigned char c = P.seqs[(size_t) P.rowidx[0] * P.seqlen + col];
        bool mut = false;
        for (int r = 1; r < P.nseqs; r++)
            if (P.seqs[(size_t) P.rowidx[r] * P.seqlen + col] != c) {
                mut = true;
                break;
            }
        P.kind[i] = mut ? AWB_SITE_VARIANT : AWB_SITE_INVARIANT;
    }
}

// likelihood_site_inner (emit.cpp:248-266) of one site by a group of 4 lanes
// (lane `a` = base a); inner[4V], mut/nomut[V] in the group's scratch
__device__ double tp_site_lk(const AwbTp &P, int site, int V, int root,
                             const short *c0, const short *c1, const short *order,
                             double *inner, const double *mut, const double *nomut,
                             int a, unsigned gmask)
{
    const size_t col = (size_t) P.start_coord + site;
    for (int q = 0; q < V; q++) {
        const int j = order[q];
        double v;
        if (c0[j] < 0) {
            const int x = tp_base(P.seqs[(size_t) P.rowidx[j] * P.seqlen + col]);
            v = (x < 0 || x == a) ? 1.0 : 0.0;
        } else {
            const int k1 = c0[j], k2 = c1[j];
            double p1 = 0.0, p2 = 0.0;
#pragma unroll
            for (int x = 0; x < 4; x++) {
                if (a == x) {
                    p1 += inner[4 * k1 + x] * nomut[k1];
                    p2 += inner[4 * k2 + x] * nomut[k2];
                } else {
                    p1 += inner[4 * k1 + x] * mut[k1];
                    p2 += inner[4 * k2 + x] * mut[k2];
                }
            }
            v = p1 * p2;
        }
        inner[4 * j + a] = v;
        __syncwarp(gmask);
    }
    double p = 0.0;
    for (int x = 0; x < 4; x++)
        p += inner[4 * root + x] * .25;
    return p;
}

__global__ void awb_tp_block_kernel(AwbTp P, int want_lik, int want_prior)
{
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;   // group = block
    const int a = threadIdx.x & 3;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = 0xFu << (lane & ~3);
    const int b = gid;
    if (b >= P.B)
        return;
    const int V = P.V, T = P.T;
    const int *parent = P.ptrees + (size_t) b * V;
    const int *age = P.ages + (size_t) b * V;
    short *c0 = P.scratch_i + (size_t) gid * 3 * V;
    short *c1 = c0 + V;
    short *order = c1 + V;
    double *inner = P.scratch_d + (size_t) gid * 6 * V;
    double *mut = inner + 4 * (size_t) V;
    double *nomut = mut + V;

    // ---- children, root, post-order (local_tree.h:274-304), branch probabilities
    int root = -1;
    for (int j = a; j < V; j += 4) {
        c0[j] = -1;
        c1[j] = -1;
    }
    __syncwarp(gmask);
    if (a == 0) {
        for (int j = 0; j < V; j++) {
            const int p = parent[j];
            if (p < 0) { root = j; continue; }
            if (c0[p] < 0) c0[p] = (short) j; else c1[p] = (short) j;
        }
        // children before parents: explicit stack, output reversed pre-order
        int top = 0, nout = V;
        short *stack = (short *) inner;             // free until the first site
        stack[top++] = (short) root;
        while (top > 0) {
            const int j = stack[--top];
            order[--nout] = (short) j;
            if (c0[j] >= 0) {
                stack[top++] = c0[j];
                stack[top++] = c1[j];
            }
        }
    }
    root = __shfl_sync(gmask, root, lane & ~3);
    __syncwarp(gmask);
    for (int j = a; j < V; j += 4) {
        double m = 0.0, nm = 0.0;
        if (j != root) {
            // emit.cpp:417-425
            const double t = fmax(P.times[age[parent[j]]] - P.times[age[j]], P.mintime);
            m = awb_prob_branch(t, P.mu, true);
            nm = awb_prob_branch(t, P.mu, false);
        }
        mut[j] = m;
        nomut[j] = nm;
    }
    __syncwarp(gmask);

    // ---- likelihood of the block (likelihood_tree, emit.cpp:399-449)
    if (want_lik) {
        const int s0 = P.block_start[b], s1 = P.block_start[b + 1];
        double lnl = 0.0, inv_lk = -1.0;
        if (V < 3) {
            lnl = log(.25) * (s1 - s0);             // total_prob.cpp:26-27
        } else {
            for (int i = s0; i < s1; i++) {
                double lk;
                const bool inv = P.kind[i] == AWB_SITE_INVARIANT;
                if (inv && inv_lk > 0) {
                    lk = inv_lk;
                } else {
                    lk = tp_site_lk(P, i, V, root, c0, c1, order, inner, mut, nomut, a,
                                    gmask);
                    if (inv)
                        inv_lk = lk;
                }
                lnl += log(lk);
            }
        }
        if (a == 0)
            P.blk_lik[b] = lnl;
    }

    // ---- prior term of the block (calc_arg_prior, total_prob.cpp:262-299)
    if (want_prior && a == 0) {
        double treelen = 0.0;                       // get_treelen(tree, false)
        for (int j = 0; j < V; j++)
            if (parent[j] >= 0)
                treelen += P.times[age[parent[j]]] - P.times[age[j]];
        const double recomb_rate = fmax(P.rho * treelen, P.rho);
        const int blocklen = P.block_start[b + 1] - P.block_start[b];
        double lnl;
        if (b + 1 < P.B) {
            lnl = log(recomb_rate) - recomb_rate * blocklen;
            // calc_spr_prob (total_prob.cpp:210-258) of the next block's SPR
            const int *spr = P.sprs + (size_t) (b + 1) * 4;
            const int rnode = spr[0], k = spr[1], j2 = spr[3];
            const int root_age = age[root];
            const double treelen_b = treelen + P.time_steps[root_age];
            // lineage counts at the times the formula reads (local_tree.cpp:34-69)
            int nbr_k = 0, nrec_k = 0, ncoal_j = 0;
            for (int x = 0; x < V; x++) {
                const int pa = parent[x] < 0 ? T - 2 : age[parent[x]];
                const int ag = age[x];
                if (ag <= k && k < pa) { nbr_k++; nrec_k++; }
                if (k == pa) { nrec_k++; if (parent[x] < 0) nbr_k++; }
                if (ag <= j2 && j2 < pa) ncoal_j++;
                if (j2 == pa) ncoal_j++;
            }
            if (k == T - 1) nbr_k = 1;
            if (k == root_age) nrec_k--;            // lineages.nrecombs[root_age]--
            lnl += log(nbr_k * P.time_steps[k] / (nrec_k * treelen_b));
            const int broken_age = age[parent[rnode]];
            const int ncoals_j = ncoal_j - (j2 <= broken_age ? 1 : 0) -
                (j2 == broken_age ? 1 : 0);
            lnl -= log((double) ncoals_j);
            // coal_rates[m] = coal_time_steps[m] * (nbranches[m/2] - [m/2 < broken_age])
            //                 / (2 popsizes[m/2])      (:196-207)
            auto coal_rate = [&](int m) {
                const int t = m / 2;
                int nb = 0;
                for (int x = 0; x < V; x++) {
                    const int pa = parent[x] < 0 ? T - 2 : age[parent[x]];
                    if (age[x] <= t && t < pa) nb++;
                    if (parent[x] < 0 && t == pa) nb++;
                }
                if (t == T - 1) nb = 1;
                nb -= (t < broken_age) ? 1 : 0;
                return P.coal_time_steps[m] * nb / (2.0 * P.popsizes[t]);
            };
            if (j2 < T - 2)
                lnl += log(1.0 - exp(-coal_rate(2 * j2) -
                                     (j2 > k ? coal_rate(2 * j2 - 1) : 0.0)));
            for (int m = 2 * k; m < 2 * j2 - 1; m++)
                lnl -= coal_rate(m);
        } else {
            lnl = -recomb_rate * blocklen;
        }
        P.blk_prior[b] = lnl;
    }
}

// fixed-order sum: every thread a contiguous range, then a tree over the threads
__global__ void awb_tp_reduce_kernel(AwbTp P)
{
    __shared__ double part[2][256];
    const int tid = threadIdx.x;
    const int per = (P.B + 255) / 256;
    double s0 = 0.0, s1 = 0.0;
    for (int b = tid * per; b < (tid + 1) * per && b < P.B; b++) {
        s0 += P.blk_lik[b];
        s1 += P.blk_prior[b];
    }
    part[0][tid] = s0;
    part[1][tid] = s1;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if (tid < d) {
            part[0][tid] += part[0][tid + d];
            part[1][tid] += part[1][tid + d];
        }
        __syncthreads();
    }
    if (tid == 0) {
        P.out[0] = part[0][0];
        P.out[1] = part[1][0];
    }
}

// ---------------------------------------------------------------- host side

static int tp_run(const awb_problem *p, double *lik, double *prior)
{
    if (!p || !p->ptrees || !p->ages || !p->sprs || !p->blocklens || !p->times)
        return awb_fail_msg("awb_arg_*: missing arrays");
    if (lik && ((!p->seqs && !p->var_cols) || !p->seqids))
        return awb_fail_msg("awb_arg_likelihood: sequences are required");
    // the alignment as variant columns: dense rows are made here (this is not
    // the throughput path)
    std::vector<unsigned char> dense;
    awb_problem pd;
    if (lik && !p->seqs) {
        dense.assign((size_t) p->nseqs * p->seqlen, p->default_char ? p->default_char : 'A');
        for (int v = 0; v < p->nvar; v++) {
            if (p->var_pos[v] < 0 || p->var_pos[v] >= p->seqlen)
                return awb_fail_msg("awb_arg_likelihood: variant position out of range");
            for (int r = 0; r < p->nseqs; r++)
                dense[(size_t) r * p->seqlen + p->var_pos[v]] =
                    p->var_cols[(size_t) v * p->nseqs + r];
        }
        pd = *p;
        pd.seqs = dense.data();
        p = &pd;
    }
    if (prior && !p->popsizes)
        return awb_fail_msg("awb_arg_prior: population sizes are required");
    if (p->ntimes < 2 || p->ntimes > AWB_MAXT)
        return awb_fail_msg("awb_arg_*: ntimes out of range");
    if (p->nnodes < 1 || p->nnodes > AWB_MAXV || p->ntrees < 1)
        return awb_fail_msg("awb_arg_*: bad tree dimensions");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return awb_fail_msg("no CUDA device available (this library has no CPU fallback)");
    const char *env = getenv("AWB_DEVICE");
    TP_OK(cudaSetDevice(env ? atoi(env) : 0));

    AwbTp P;
    memset(&P, 0, sizeof(P));
    const int T = P.T = p->ntimes, V = P.V = p->nnodes, B = P.B = p->ntrees;
    P.nleaves = (V + 1) / 2;
    P.nseqs = P.nleaves;
    P.seqlen = p->seqlen;
    P.start_coord = p->start_coord;
    P.rho = p->rho;
    P.mu = p->mu;
    // model.h:21-32, :322-333, model.cpp:9-23 (as awb_layout.h awb_model_fill)
    for (int i = 0; i < T; i++) {
        P.times[i] = p->times[i];
        P.popsizes[i] = p->popsizes ? p->popsizes[i] : 1.0;
    }
    for (int i = 0; i < T - 1; i++)
        P.time_steps[i] = P.times[i + 1] - P.times[i];
    P.time_steps[T - 1] = INFINITY;
    {
        std::vector<double> mid(T);
        for (int i = 0; i < T - 1; i++)
            mid[i] = sqrt((P.times[i + 1] + 1.0) * (P.times[i] + 1.0));
        for (int i = 0; i < T - 1; i++) {
            P.coal_time_steps[2 * i] = mid[i] - P.times[i];
            P.coal_time_steps[2 * i + 1] = P.times[i + 1] - mid[i];
        }
        P.coal_time_steps[2 * T - 2] = INFINITY;
    }
    P.mintime = 0.1 * P.times[1];

    std::vector<int> bstart(B + 1, 0);
    for (int b = 0; b < B; b++) {
        if (p->blocklens[b] < 1)
            return awb_fail_msg("awb_arg_*: blocklen must be >= 1");
        bstart[b + 1] = bstart[b] + p->blocklens[b];
    }
    const int n = P.n = bstart[B];
    for (size_t x = 0; x < (size_t) B * V; x++) {
        const int pa = p->ptrees[x], ag = p->ages[x];
        if (pa < -1 || pa >= V || ag < 0 || ag >= T - 1)
            return awb_fail_msg("awb_arg_*: tree arrays out of range");
    }
    if (lik) {
        if (p->nseqs < P.nleaves || p->start_coord < 0 ||
            p->start_coord + n > p->seqlen)
            return awb_fail_msg("awb_arg_likelihood: sequences do not cover the trees");
        for (int j = 0; j < P.nleaves; j++)
            if (p->seqids[j] < 0 || p->seqids[j] >= p->nseqs)
                return awb_fail_msg("awb_arg_likelihood: bad seqids");
    }

    // one device allocation
    const int ngroups = B;
    size_t off = 0;
    auto place = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t) 255; return o; };
    const size_t o_pt = place((size_t) B * V * 4), o_ag = place((size_t) B * V * 4),
        o_sp = place((size_t) B * 16), o_bs = place((size_t) (B + 1) * 4),
        o_ri = place((size_t) P.nleaves * 4),
        o_sq = place(lik ? (size_t) p->nseqs * p->seqlen : 0),
        o_kd = place(n), o_bl = place((size_t) B * 8), o_bp = place((size_t) B * 8),
        o_out = place(16), o_si = place((size_t) ngroups * 3 * V * 2),
        o_sd = place((size_t) ngroups * 6 * V * 8);
    char *d = NULL;
    TP_OK(cudaMalloc((void **) &d, off));
    int rc = 0;
    do {
#define TP_TRY(call) if ((call) != cudaSuccess) { rc = awb_fail_msg(std::string(#call) + ": " + cudaGetErrorString(cudaGetLastError())); break; }
        TP_TRY(cudaMemcpy(d + o_pt, p->ptrees, (size_t) B * V * 4, cudaMemcpyHostToDevice));
        TP_TRY(cudaMemcpy(d + o_ag, p->ages, (size_t) B * V * 4, cudaMemcpyHostToDevice));
        TP_TRY(cudaMemcpy(d + o_sp, p->sprs, (size_t) B * 16, cudaMemcpyHostToDevice));
        TP_TRY(cudaMemcpy(d + o_bs, bstart.data(), (size_t) (B + 1) * 4, cudaMemcpyHostToDevice));
        if (lik) {
            TP_TRY(cudaMemcpy(d + o_ri, p->seqids, (size_t) P.nleaves * 4, cudaMemcpyHostToDevice));
            TP_TRY(cudaMemcpy(d + o_sq, p->seqs, (size_t) p->nseqs * p->seqlen, cudaMemcpyHostToDevice));
        }
        TP_TRY(cudaMemset(d + o_bl, 0, (size_t) B * 8));
        TP_TRY(cudaMemset(d + o_bp, 0, (size_t) B * 8));
        P.ptrees = (const int *) (d + o_pt);
        P.ages = (const int *) (d + o_ag);
        P.sprs = (const int *) (d + o_sp);
        P.block_start = (const int *) (d + o_bs);
        P.rowidx = (const int *) (d + o_ri);
        P.seqs = (const unsigned char *) (d + o_sq);
        P.kind = (unsigned char *) (d + o_kd);
        P.blk_lik = (double *) (d + o_bl);
        P.blk_prior = (double *) (d + o_bp);
        P.out = (double *) (d + o_out);
        P.scratch_i = (short *) (d + o_si);
        P.scratch_d = (double *) (d + o_sd);
        if (lik) {
            int gx = (n + 255) / 256;
            if (gx > 4096) gx = 4096;
            awb_tp_kind_kernel<<<gx, 256>>>(P);
        }
        awb_tp_block_kernel<<<(B * 4 + 127) / 128, 128>>>(P, lik ? 1 : 0, prior ? 1 : 0);
        awb_tp_reduce_kernel<<<1, 256>>>(P);
        TP_TRY(cudaGetLastError());
        double out[2];
        TP_TRY(cudaMemcpy(out, d + o_out, 16, cudaMemcpyDeviceToHost));
        if (lik) *lik = out[0];
        if (prior) *prior = out[1];
#undef TP_TRY
    } while (0);
    cudaFree(d);
    return rc;
}

extern "C" int awb_arg_likelihood(const awb_problem *arg, double *lnl)
{
    return tp_run(arg, lnl, NULL);
}

extern "C" int awb_arg_prior(const awb_problem *arg, double *lnl)
{
    return tp_run(arg, NULL, lnl);
}

extern "C" int awb_arg_joint(const awb_problem *arg, double *lik, double *prior)
{
    return tp_run(arg, lik, prior);
}
