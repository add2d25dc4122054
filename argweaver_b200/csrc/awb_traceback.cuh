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
//
// Latency.  A block (one local tree, ~30 sites) is a chain of dependent phases,
// so every global load that sits between two phases costs a full DRAM round
// trip.  Nothing but the forward rows themselves is read from global memory on
// that chain: the per-block tables (transition vectors, the time-by-time
// matrix, per-state node/time/age, the switch CSR of the block and the last
// forward row of the block before it) are copied into shared memory with
// cp.async.bulk (one instruction per table, completion counted in bytes by an
// mbarrier) ONE BLOCK AHEAD, double buffered; the block scalars are fetched two
// blocks ahead; the forward rows of the next block's first wave are prefetched
// into L2; and a warp issues all loads of its rows before it consumes any.
#ifndef AWB_TRACEBACK_CUH
#define AWB_TRACEBACK_CUH

#include "awb_common.cuh"

#define AWB_TB_THREADS 512

struct AwbTbSmem {
    double wsum[32];
    int kmin;
    int kcur;
    int br_first, br_cnt;           // states of the branch of k: [first, end)
    unsigned failmask[3];
};

// scalars of one block.  awb_tb_blk only LOADS (two blocks ahead of use, so no
// instruction waits on the loads); the derived values come from the accessors.
struct AwbTbBlk {
    int S, blen, pos, Sprev, minage;
    long long r0, fwoff, entoff, entend;
    __device__ int S1() const { return S > 0 ? S : 1; }
    __device__ int n1() const { return Sprev > 0 ? Sprev : 1; }
    __device__ int nent() const { return (int) (entend - entoff); }
};

// pointers of the chain the kernel walks (read once from the chain record)
struct AwbTbPtrs {
    const int *nstates, *blocklens, *block_start;
    const long long *row_off, *fw_off, *ent_off;
    const int *tm_minage;       // only read for the closed-form transitions
};

// (bb below the first block of the launch, or the block of which a segment
// only holds the first site, are handled through bmin / bextra)
__device__ inline AwbTbBlk awb_tb_blk(const AwbTbPtrs &p, int bb, int bmin, int bextra)
{
    AwbTbBlk m;
    if (bb < bmin) {
        m.S = 0; m.blen = 0; m.pos = 0; m.Sprev = 0; m.minage = 0;
        m.r0 = 0; m.fwoff = 0; m.entoff = 0; m.entend = 0;
        return m;
    }
    m.S = p.nstates[bb];
    m.blen = (bb == bextra) ? 1 : p.blocklens[bb];
    m.pos = p.block_start[bb];
    m.r0 = p.row_off[bb];
    m.fwoff = p.fw_off[bb];
    m.entoff = p.ent_off[bb];
    m.entend = p.ent_off[bb + 1];
    m.Sprev = bb > bmin ? p.nstates[bb - 1] : 0;
    m.minage = p.tm_minage ? p.tm_minage[bb] : 0;
    return m;
}

// per-block tables in shared memory (one of two buffers); byte offsets
struct AwbTbBuf {
    unsigned tv, tm, last, ep;          // doubles
    unsigned es, sws, swc, stn;         // 16-bit
    unsigned stt, sta;                  // 8-bit
    unsigned bytes;
};

__host__ __device__ inline AwbTbBuf awb_tb_buf_layout(int maxS1, int maxT, int maxent)
{
    // every array starts on a 16-byte boundary and keeps the source's offset
    // within a 16-byte line (the bulk copies move whole aligned lines): slack
    // for that offset and for the rounded-up tail
    AwbTbBuf L;
    unsigned o = 0;
#define AWB_TB_SLOT(name, bytes) do { L.name = o; o += (((unsigned) (bytes)) + 32u + 15u) & ~15u; } while (0)
    AWB_TB_SLOT(tv, (AWB_TM_NVEC * maxT) * 8);
    AWB_TB_SLOT(tm, (maxT * maxT) * 8);
    AWB_TB_SLOT(last, maxS1 * 8);
    AWB_TB_SLOT(ep, (maxent + 1) * 8);
    AWB_TB_SLOT(es, maxent * 2);
    AWB_TB_SLOT(sws, maxS1 * 2);
    AWB_TB_SLOT(swc, maxS1 * 2);
    AWB_TB_SLOT(stn, maxS1 * 2);
    AWB_TB_SLOT(stt, maxS1);
    AWB_TB_SLOT(sta, maxS1);
#undef AWB_TB_SLOT
    L.bytes = (o + 15u) & ~15u;
    return L;
}

// dynamic shared memory of the kernel
__host__ __device__ inline size_t awb_tb_smem_bytes(int maxS1, int maxT, int maxent)
{
    const AwbTbBuf L = awb_tb_buf_layout(maxS1, maxT, maxent);
    size_t n = 2 * (size_t) L.bytes;
    n += 2 * (size_t) maxS1 * 8;                   // transS, corrS
    n += (size_t) AWB_MAXT * 8;                    // tmcolS
    n += (size_t) (maxent + 1) * 8;                // swA
    n += (((size_t) (maxent + 1) * 2) + 15) & ~(size_t) 15;   // swJ
    return n + 16;
}

// ---- bulk asynchronous copies (cp.async.bulk, the non-tensor form of TMA) with
// an mbarrier that counts the bytes: one thread issues one instruction per
// table instead of 512 threads issuing 4- and 8-byte cp.async each.
__device__ __forceinline__ void awb_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void awb_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void awb_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "AWB_MBAR_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra AWB_MBAR_DONE;\n\t"
                 "bra AWB_MBAR_WAIT;\n\t"
                 "AWB_MBAR_DONE:\n\t}"
                 :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ void awb_bulk_g2s(unsigned dst, const void *src, unsigned bytes,
                                             unsigned bar)
{
    // (whole aligned 16-byte lines: the caller passes the line that holds the
    // first byte and a size rounded up; element 0 lands at dst + (src & 15))
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
                 "[%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Block-wide sample() over the S1 weights held VPT per thread (thread t holds
// indices t*VPT .. t*VPT+VPT-1): returns the chosen index.
template <int VPT>
__device__ inline int awb_block_sample(const double (&A)[VPT], int S1, int r,
                                       int rand_max, AwbTbSmem *sm)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    double loc[VPT];
    double x = 0.0;
#pragma unroll
    for (int v = 0; v < VPT; v++) {
        x += A[v];
        loc[v] = x;                     // inclusive within the thread
    }
    const double mine = x;
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
    const double before = off + (x - mine);     // sum of everything before my chunk
    const double pick = (double) r / (double) rand_max * total;
    int hitidx = 0x7fffffff;
#pragma unroll
    for (int v = VPT - 1; v >= 0; v--) {
        const int j = tid * VPT + v;
        if (j < S1 && before + loc[v] >= pick)
            hitidx = j;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hitidx != 0x7fffffff);
    if (m != 0) {
        const int first = __shfl_sync(0xffffffffu, hitidx, __ffs(m) - 1);
        if (lane == 0)
            atomicMin(&sm->kmin, first);
    }
    __syncthreads();
    const int k = sm->kmin;
    __syncthreads();                // kmin / wsum are reused by the next call
    return k;
}

// NV = ceil(max states / 32) values per lane and row, SPW = rows per warp and
// wave, VPT = ceil(max states / threads)
template <int NV, int SPW, int VPT>
__global__ void __launch_bounds__(AWB_TB_THREADS, 1)
awb_traceback_kernel(const AwbChain *chains, int rand_max, int maxS1, int maxT,
                     int maxent, int seg, int closed_form)
{
    extern __shared__ __align__(16) unsigned char tb_smem[];
    const AwbChain &ch = chains[blockIdx.x];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int NW = AWB_TB_THREADS >> 5;
    constexpr int NVH = (NV + 1) / 2;      // the shorter side of a row, per lane
    constexpr int NVC = NVH < 16 ? NVH : 16;   // ... read NVC values per lane and pass
    const int T = ch.model.ntimes;
    const int n = ch.nsites;
    const int B = ch.ntrees;
    // the blocks of this launch (checkpointed table: one segment, see AwbSeg)
    const AwbSeg g = awb_seg(ch, seg);
    if (!g.valid)
        return;
    const int bmin = g.b0, btop = g.b1 - 1 + g.extra;
    const int bextra = g.extra ? g.b1 : -1;
    const double *__restrict__ fwg = ch.fw - g.fwbias;
    const double *__restrict__ fsumg = ch.fsum + g.fsoff - (long long) g.site0 * (T - 1);
    const int *__restrict__ randg = ch.rand_ints;
    int *__restrict__ pathg = ch.path;
    // closed_form: the batch's linear-domain vectors may overflow (awb_layout.h,
    // lin_unsafe); same-branch transitions then come from the reference's closed
    // form (awb_get_time over the 9 log-domain vectors) instead
    const AwbTbPtrs P = { ch.nstates, ch.blocklens, ch.block_start,
                          ch.row_off, ch.fw_off, ch.ent_off,
                          closed_form ? ch.tm_minage : (const int *) 0 };
    const double *__restrict__ tmvecg = ch.tmvec;
    const double *__restrict__ ling = ch.lin;
    const double *__restrict__ tmatrixg = ch.tmatrix;
    const short *__restrict__ st_nodeg = ch.st_node;
    const signed char *__restrict__ st_timeg = ch.st_time;
    const signed char *__restrict__ st_ageg = ch.st_age;
    const double *__restrict__ sw_probg = ch.sw_prob;
    const unsigned short *__restrict__ sw_srcg = ch.sw_src;
    const unsigned short *__restrict__ sw_startg = ch.sw_start;
    const unsigned short *__restrict__ sw_cntg = ch.sw_cnt;

    __shared__ AwbTbSmem sm;
    const AwbTbBuf BL = awb_tb_buf_layout(maxS1, maxT, maxent);
    unsigned char *bufp[2] = { tb_smem, tb_smem + BL.bytes };
    double *transS = (double *) (tb_smem + 2 * (size_t) BL.bytes);
    double *corrS = transS + maxS1;         // transS[j] - tm[a_j][b_k] (branch of k)
    double *tmcolS = corrS + maxS1;         // tm[.][b_k]
    double *swA = tmcolS + AWB_MAXT;
    unsigned short *swJ = (unsigned short *) (swA + (maxent + 1));
    const unsigned smem_s = (unsigned) __cvta_generic_to_shared(tb_smem);

    // offset of element 0 of every array of each buffer within its 16-byte line
    // (source alignment)
    unsigned sk_es[2], sk_sws[2], sk_swc[2], sk_stn[2], sk_stt[2], sk_sta[2];
    unsigned sk_tv[2], sk_tm[2], sk_last[2], sk_ep[2];
    // one mbarrier per buffer: thread 0 announces the bytes of a block's tables
    // and sends the copies, everybody waits on the phase
    __shared__ __align__(8) unsigned long long tb_mbar[2];
    const unsigned mbar_s = (unsigned) __cvta_generic_to_shared(tb_mbar);
    unsigned phase[2] = { 0u, 0u };
    if (tid == 0) {
        awb_mbar_init(mbar_s, 1);
        awb_mbar_init(mbar_s + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // issue the asynchronous copy of block bb's tables into buffer q.  Every
    // thread needs the skews; thread 0 alone adds up the bytes, announces them to
    // the barrier and sends the copies (its plan is kept in registers meanwhile).
    auto preload = [&](const AwbTbBlk &m, int bb, int q) {
        const unsigned base = smem_s + (unsigned) q * BL.bytes;
        const unsigned bar = mbar_s + 8u * (unsigned) q;
        sk_es[q] = sk_sws[q] = sk_swc[q] = sk_stn[q] = sk_stt[q] = sk_sta[q] = 0;
        sk_tv[q] = sk_tm[q] = sk_last[q] = sk_ep[q] = 0;
        const bool have = bb >= bmin && m.S > 0, sw = bb > bmin;
        // table x of the block: source, bytes, destination (nothing kept in
        // registers between the uses: thread 0 evaluates it again)
        auto table = [&](int x, const void *&src, int &nby, unsigned &dst) {
            src = 0; nby = 0; dst = base;
            switch (x) {
            case 0:
                src = closed_form ? (const void *) (tmvecg + (size_t) bb * AWB_TM_NVEC * T)
                                  : (const void *) (ling + (size_t) bb * 7 * T);
                nby = have ? (closed_form ? 8 * AWB_TM_NVEC * T : 8 * 7 * T) : 0;
                dst = base + BL.tv; break;
            case 1: src = tmatrixg + (size_t) bb * T * T; nby = have ? 8 * T * T : 0;
                dst = base + BL.tm; break;
            case 2: src = st_nodeg + m.r0; nby = have ? 2 * m.S : 0; dst = base + BL.stn; break;
            case 3: src = st_timeg + m.r0; nby = have ? m.S : 0; dst = base + BL.stt; break;
            case 4: src = st_ageg + m.r0; nby = have ? m.S : 0; dst = base + BL.sta; break;
            case 5: src = fwg + m.fwoff - m.n1(); nby = sw ? 8 * m.n1() : 0;
                dst = base + BL.last; break;
            case 6: src = sw_probg + m.entoff; nby = sw ? 8 * m.nent() : 0;
                dst = base + BL.ep; break;
            case 7: src = sw_srcg + m.entoff; nby = sw ? 2 * m.nent() : 0;
                dst = base + BL.es; break;
            case 8: src = sw_startg + m.r0; nby = sw ? 2 * m.S1() : 0; dst = base + BL.sws; break;
            default: src = sw_cntg + m.r0; nby = sw ? 2 * m.S1() : 0; dst = base + BL.swc; break;
            }
        };
        unsigned sk[10];
#pragma unroll
        for (int x = 0; x < 10; x++) {
            const void *src; int nby; unsigned dst;
            table(x, src, nby, dst);
            sk[x] = nby > 0 ? (unsigned) ((unsigned long long) src & 15ull) : 0u;
        }
        sk_tv[q] = sk[0]; sk_tm[q] = sk[1]; sk_stn[q] = sk[2]; sk_stt[q] = sk[3];
        sk_sta[q] = sk[4]; sk_last[q] = sk[5]; sk_ep[q] = sk[6]; sk_es[q] = sk[7];
        sk_sws[q] = sk[8]; sk_swc[q] = sk[9];
        if (tid == 0) {
            unsigned total = 0;
#pragma unroll
            for (int x = 0; x < 10; x++) {
                const void *src; int nby; unsigned dst;
                table(x, src, nby, dst);
                if (nby > 0)
                    total += (sk[x] + (unsigned) nby + 15u) & ~15u;
            }
            awb_mbar_expect_tx(bar, total);
#pragma unroll
            for (int x = 0; x < 10; x++) {
                const void *src; int nby; unsigned dst;
                table(x, src, nby, dst);
                if (nby > 0)
                    awb_bulk_g2s(dst, (const char *) src - sk[x],
                                 (sk[x] + (unsigned) nby + 15u) & ~15u, bar);
            }
        }
    };
    // wait until buffer q's tables have landed
    auto landed = [&](int q) {
        awb_mbar_wait(mbar_s + 8u * (unsigned) q, phase[q]);
        phase[q] ^= 1u;
    };
    // forward rows of the first wave of block m into L2
    auto prefetch_rows = [&](const AwbTbBlk &m) {
        const int nrows = m.blen - 1 < NW * SPW ? m.blen - 1 : NW * SPW;
        if (nrows <= 0)
            return;
        const char *p0 = (const char *) (fwg + m.fwoff +
                                         (long long) (m.blen - 1 - nrows) * m.S1());
        const long long nbytes = (long long) nrows * m.S1() * 8;
        for (long long o = (long long) tid * 128; o < nbytes; o += 128ll * AWB_TB_THREADS)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(p0 + o));
        // ... and their per-time sums (the wave's row totals come from them)
        const char *f0 = (const char *) (fsumg + (size_t) (m.pos + m.blen - 1 - nrows) * (T - 1));
        const long long fbytes = (long long) nrows * (T - 1) * 8;
        for (long long o = (long long) tid * 128; o < fbytes + 128; o += 128ll * AWB_TB_THREADS)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(f0 + o));
    };

    // draw used for site s: the reference consumes one rand() per sampled site,
    // last site first
    const int roff = (ch.last_state < 0) ? (n - 1) : (n - 2);
    int k;

    AwbTbBlk mC = awb_tb_blk(P, btop, bmin, bextra);
    AwbTbBlk mN = awb_tb_blk(P, btop - 1, bmin, bextra);
    if (tid == 0) {
        sm.failmask[0] = 0;
        sm.failmask[1] = 0;
        sm.failmask[2] = 0;
    }
    preload(mC, btop, btop & 1);

    // ---- last column (sample_thread.cpp:534-539)
    {
        const int S1 = mC.S1();
        if (g.extra) {
            // the state at the extra site was sampled with the segment after this one
            k = pathg[ch.block_start[g.b1]];
        } else if (ch.last_state < 0) {
            double A[VPT];
#pragma unroll
            for (int v = 0; v < VPT; v++) {
                const int j = tid * VPT + v;
                A[v] = j < S1 ? fwg[ch.fw_off[B] - S1 + j] : 0.0;
            }
            k = awb_block_sample<VPT>(A, S1, randg[0], rand_max, &sm);
        } else {
            k = ch.last_state;
        }
        if (tid == 0 && !g.extra)
            pathg[n - 1] = k;
    }
    landed(btop & 1);
    __syncthreads();

    int par = 0;
    for (int b = btop; b >= bmin; b--) {
        const int q = b & 1;
        // ---- one block ahead: tables of block b-1; two ahead: scalars of b-2
        preload(mN, b - 1, q ^ 1);
        const AwbTbBlk mNN = awb_tb_blk(P, b - 2, bmin, bextra);
        if (b > bmin)
            prefetch_rows(mN);

        const int S = mC.S, S1 = mC.S1(), blen = mC.blen, pos = mC.pos;
        const double *fw = fwg + mC.fwoff;
        // draw of the switch step, fetched now so that it is there when needed
        const int r_sw = (b > bmin) ? randg[roff - (pos - 1)] : 0;
        const unsigned char *bp = bufp[q];
        const double *linS = (const double *) (bp + BL.tv + sk_tv[q]);   // lin[7][T] (K1)
        const double *tmS = (const double *) (bp + BL.tm + sk_tm[q]);
        const short *stN = (const short *) (bp + BL.stn + sk_stn[q]);
        const signed char *stT = (const signed char *) (bp + BL.stt + sk_stt[q]);
        const signed char *stA = (const signed char *) (bp + BL.sta + sk_sta[q]);

        // ---- sample_hmm_posterior (sample_thread.cpp:470-503), speculative
        int i_hi = blen - 2;
        int trans_k = -1;
        if (S == 0) {
            // one-state space: every site keeps state 0 (its draw is skipped,
            // the draws are indexed by position)
            for (int i = tid; i <= i_hi; i += AWB_TB_THREADS)
                pathg[pos + i] = 0;
            i_hi = -1;
        }
        while (i_hi >= 0) {
            if (trans_k != k) {
                // transition column into k (recomputed only when k changes):
                // other branches read the time-by-time matrix (which already
                // holds the minage rule), the branch of k adds its band term
                if (S > 0) {
                    const int node_k = stN[k];
                    const int b_k = stT[k];
                    const int c_k = stA[k];
                    // same-branch transitions in the separable, linear-domain
                    // form the forward kernel uses (awb_forward_fast.cuh):
                    // T[(n,a)->(n,b)] = tm[a][b] + D[a] {A1_b (h[a]-Bc) | A2_b | A3_b}
                    //                   (+ norecombs[b] when a == b)
                    const double Bc = c_k > 0 ? linS[2 * T + c_k - 1] : 0.0;
                    const double A1 = linS[3 * T + b_k];
                    const double A2 = linS[4 * T + b_k] - A1 * Bc;
                    const double A3 = linS[5 * T + b_k] - A1 * Bc;
                    const double nrb = linS[6 * T + b_k];
                    for (int j = tid; j < S; j += AWB_TB_THREADS) {
                        const int a_j = stT[j];
                        const double other = tmS[a_j * T + b_k];
                        const bool same = stN[j] == node_k;
                        double tr = other;
                        if (same && closed_form) {
                            tr = awb_get_time(linS, T, a_j, b_k, c_k, mC.minage, true);
                        } else if (same) {
                            const double Da = linS[a_j];
                            const double co = (a_j < b_k) ? A1 * (linS[T + a_j] - Bc) :
                                (a_j == b_k ? A2 : A3);
                            tr = other + (Da * co + (a_j == b_k ? nrb : 0.0));
                        }
                        transS[j] = tr;
                        corrS[j] = tr - other;
                        // the states of a node are contiguous in state order
                        if (same && (j == 0 || stN[j - 1] != node_k))
                            sm.br_first = j;
                        if (same && (j == S - 1 || stN[j + 1] != node_k))
                            sm.br_cnt = j + 1;          // end of the branch
                    }
                    if (tid < T - 1)
                        tmcolS[tid] = tmS[tid * T + b_k];
                } else if (tid == 0) {
                    transS[0] = 1.0;
                }
                trans_k = k;
                __syncthreads();
            }

            // ---- one wave: NW * SPW sites tested for "stays in k".  The row
            // total comes from the per-time sums the forward kernel stored
            // (tot = sum_a Fn[a] tm[a][b_k] + correction on the branch of k), so
            // only the shorter side of the row is read: the states before k
            // (pre directly) or the states after k (pre = tot - A_k - suffix).
            {
                const int fk = sm.br_first;
                const int ck = sm.br_cnt - fk;
                const bool lower = (k <= S1 - 1 - k);
                const int scnt = lower ? k : S1 - 1 - k;
                const int sbase = lower ? 0 : k + 1;
                int r[SPW];
                double v[SPW][NVC], f[SPW][2], br[SPW][2], rk[SPW];
                double tot[SPW], side[SPW];
#pragma unroll
                for (int u = 0; u < SPW; u++) {
                    const int i = i_hi - (warp * SPW + u);
                    const bool ok = i >= 0;
                    r[u] = (ok && lane == 0) ? randg[roff - (pos + i)] : 0;
                    const double *row = fw + (long long) (ok ? i : 0) * S1;
                    const double *Fn = fsumg + (size_t) (pos + (ok ? i : 0)) * (T - 1);
                    rk[u] = (ok && lane == 0) ? row[k] : 0.0;
#pragma unroll
                    for (int x = 0; x < 2; x++) {
                        const int a = lane + 32 * x;
                        f[u][x] = (ok && a < T - 1) ? Fn[a] : 0.0;
                        br[u][x] = (ok && a < ck) ? row[fk + a] : 0.0;
                    }
#pragma unroll
                    for (int x = 0; x < NVC; x++) {
                        const int j = lane + 32 * x;
                        v[u][x] = (ok && j < scnt) ? row[sbase + j] : 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < SPW; u++) {
                    tot[u] = 0.0;
                    side[u] = 0.0;
#pragma unroll
                    for (int x = 0; x < 2; x++) {
                        const int a = lane + 32 * x;
                        if (a < T - 1) tot[u] = fma(f[u][x], tmcolS[a], tot[u]);
                        if (a < ck) tot[u] = fma(br[u][x], corrS[fk + a], tot[u]);
                    }
#pragma unroll
                    for (int x = 0; x < NVC; x++) {
                        const int j = lane + 32 * x;
                        if (j < scnt) side[u] = fma(v[u][x], transS[sbase + j], side[u]);
                    }
                }
                if (NVH > NVC) {
                    // state spaces beyond 1024: the rest of the side in more passes
                    for (int base = 32 * NVC; base < scnt; base += 32 * NVC) {
#pragma unroll
                        for (int u = 0; u < SPW; u++) {
                            const int i = i_hi - (warp * SPW + u);
                            const double *row = fw + (long long) (i >= 0 ? i : 0) * S1;
#pragma unroll
                            for (int x = 0; x < NVC; x++) {
                                const int j = base + lane + 32 * x;
                                v[u][x] = (i >= 0 && j < scnt) ? row[sbase + j] : 0.0;
                            }
                        }
#pragma unroll
                        for (int u = 0; u < SPW; u++)
#pragma unroll
                            for (int x = 0; x < NVC; x++) {
                                const int j = base + lane + 32 * x;
                                if (j < scnt)
                                    side[u] = fma(v[u][x], transS[sbase + j], side[u]);
                            }
                    }
                }
#pragma unroll
                for (int u = 0; u < SPW; u++) {
                    const int w = warp * SPW + u;          // w-th site of the wave
                    const int i = i_hi - w;
                    double tt = tot[u], ss = side[u];
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) {
                        tt += __shfl_xor_sync(0xffffffffu, tt, d);
                        ss += __shfl_xor_sync(0xffffffffu, ss, d);
                    }
                    if (lane == 0 && i >= 0) {
                        const double ak = rk[u] * transS[k];
                        const double pick = (double) r[u] / (double) rand_max * tt;
                        const double hi = lower ? ss + ak : tt - ss;
                        const double lo = lower ? ss : hi - ak;
                        if (!((lo < pick) && (hi >= pick)))
                            atomicOr(&sm.failmask[par], 1u << w);
                    }
                }
            }
            __syncthreads();

            // first site of the wave (highest i) that leaves k
            // three masks in rotation: the one reset here is used two waves
            // from now, i.e. after the next barrier
            const unsigned fm = sm.failmask[par];
            if (tid == 0)
                sm.failmask[par == 0 ? 2 : par - 1] = 0;
            par = (par == 2) ? 0 : par + 1;
            const int wave = NW * SPW;
            const int fail = fm ? __ffs(fm) - 1 : -1;
            const int nkeep = (fail < 0) ? wave : fail;
            if (tid < nkeep && i_hi - tid >= 0)
                pathg[pos + i_hi - tid] = k;
            if (fail >= 0) {
                // sample the failing site in full
                const int i = i_hi - fail;
                double A[VPT];
#pragma unroll
                for (int x = 0; x < VPT; x++) {
                    const int j = tid * VPT + x;
                    A[x] = j < S1 ? fw[(long long) i * S1 + j] * transS[j] : 0.0;
                }
                const int knew = awb_block_sample<VPT>(A, S1, randg[roff - (pos + i)],
                                                       rand_max, &sm);
                if (tid == 0)
                    pathg[pos + i] = knew;
                k = knew;
                i_hi = i - 1;
            } else {
                i_hi -= wave;
            }
        }

        // ---- sample_hmm_posterior_step through the switch matrix (:506-519)
        if (b > bmin) {
            if (tid == 0) {
                const int n1 = mC.n1();
                const double *col1 = (const double *) (bp + BL.last + sk_last[q]);
                const unsigned short *sws = (const unsigned short *) (bp + BL.sws + sk_sws[q]);
                const unsigned short *swc = (const unsigned short *) (bp + BL.swc + sk_swc[q]);
                const int st = sws[k];
                const int cn = swc[k];
                const unsigned short *es =
                    (const unsigned short *) (bp + BL.es + sk_es[q]) + st;
                const double *ep = (const double *) (bp + BL.ep + sk_ep[q]) + st;
                // entries sorted by source index = the order sample() walks A[]
                for (int x = 0; x < cn; x++) {
                    const unsigned short jj = es[x];
                    const double val = col1[jj] * ep[x];
                    int w = x;
                    while (w > 0 && swJ[w - 1] > jj) {
                        swJ[w] = swJ[w - 1];
                        swA[w] = swA[w - 1];
                        w--;
                    }
                    swJ[w] = jj;
                    swA[w] = val;
                }
                double total = 0.0;
                for (int x = 0; x < cn; x++)
                    total += swA[x];
                const double pick = (double) r_sw / (double) rand_max * total;
                // zero-weight states before the first entry win when pick == 0
                int kk = n1 - 1;
                double x = 0.0;
                if (0.0 >= pick && (cn == 0 || swJ[0] > 0)) {
                    kk = 0;
                } else {
                    for (int y = 0; y < cn; y++) {
                        x += swA[y];
                        if (x >= pick) { kk = swJ[y]; break; }
                    }
                }
                sm.kcur = kk;
                pathg[pos - 1] = kk;
            }
        }

        // ---- the tables of block b-1 have landed; everyone is done with buffer q
        // (one barrier for both: it also publishes the state thread 0 sampled)
        landed(q ^ 1);
        __syncthreads();
        if (b > bmin)
            k = sm.kcur;
        mC = mN;
        mN = mNN;
    }
}

#endif // AWB_TRACEBACK_CUH
