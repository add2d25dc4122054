// awb_api.cu -- the flat C ABI (include/argweaver_b200.h) over the CUDA kernels.
//
// Kernels (every launch covers all chains of the batch, or one upload group
// of them during setup):
//   awb_kind_kernel          thread per site     site classification
//   awb_block_setup_kernel   thread per block    K1 (awb_setup.cuh)
//   awb_tmatrix_kernel       thread per entry    time-by-time matrices of K1
//   awb_switch_setup_kernel  warp per breakpoint K2 (awb_setup.cuh)
//   awb_emit_kernel          warp per site       K3 (awb_emit.cuh), variant sites
//   awb_forward_fast_kernel  CTA per chain       K4 (awb_forward_fast.cuh);
//   awb_forward_kernel                           generic K4 (awb_forward.cuh)
//   awb_traceback_kernel     CTA per chain       K5 (awb_traceback.cuh)
// With AWB_CHECKPOINT the last three run once per segment of the window, and
// the forward table holds one segment at a time.
//
// There is no CPU fallback: without a usable CUDA device every entry point
// returns an error.

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "argweaver_b200.h"
#include "awb_common.cuh"
#include "awb_emit.cuh"
#include "awb_forward.cuh"
#include "awb_forward_fast.cuh"
#include "awb_layout.h"
#include "awb_recomb.cuh"
#include "awb_setup.cuh"
#include "awb_traceback.cuh"

// ---------------------------------------------------------------- errors

static thread_local std::string g_err;

static int fail(const std::string &msg)
{
    g_err = msg;
    return 1;
}

// (for the other translation units of the library)
int awb_fail_msg(const std::string &msg) { return fail(msg); }

#define CUDA_OK(call)                                                        \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess)                                               \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char *awb_last_error(void) { return g_err.c_str(); }

extern "C" int awb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}

// Host threads for the layout / staging work of one batch: the machine's cores
// shared out over the ranks of this node (torchrun's LOCAL_WORLD_SIZE: eight
// ranks that each spawn a thread per core oversubscribe the host eightfold);
// AWB_HOST_THREADS overrides.
static int host_threads(int nwork)
{
    unsigned hw = std::thread::hardware_concurrency();
    int n = hw ? (int) hw : 1;
    const char *lws = getenv("LOCAL_WORLD_SIZE");
    if (lws && atoi(lws) > 1)
        n = (n + atoi(lws) - 1) / atoi(lws);
    const char *env = getenv("AWB_HOST_THREADS");
    if (env && atoi(env) > 0)
        n = atoi(env);
    if (n > nwork) n = nwork;
    return n < 1 ? 1 : n;
}

// ---------------------------------------------------------------- kernels

__global__ void awb_kind_kernel(const AwbChain *chains)
{
    const AwbChain &ch = chains[blockIdx.y];
    if (ch.seqs) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ch.nsites;
             i += gridDim.x * blockDim.x)
            awb_site_kind(ch, i);
        return;
    }
    // the alignment as variant columns: every other site is the default
    // character in every row; then one thread per variant column
    const unsigned char dflt = ch.default_char == 'N' ? AWB_SITE_MASKED : AWB_SITE_INVARIANT;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ch.nsites;
         i += gridDim.x * blockDim.x)
        ch.kind[i] = dflt;
}

__global__ void awb_kind_packed_kernel(const AwbChain *chains)
{
    const AwbChain &ch = chains[blockIdx.y];
    if (ch.seqs)
        return;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < ch.nvar;
         v += gridDim.x * blockDim.x) {
        const long long i = (long long) ch.var_pos[v] - ch.start_coord;
        if (i >= 0 && i < ch.nsites)
            awb_site_kind(ch, (int) i, v);
    }
}

// K1: a thread per block.  Thread-per-block work is bound by the number of
// memory instructions -- every access of a warp goes to 32 different lines, ~32
// cycles of the L1 pipe each -- which is why the worker packs its state tables
// into 8-byte stores and keeps counters in difference arrays (awb_setup.cuh).
// (tried: __launch_bounds__(64, 16) -- 64 registers instead of 86, 16 instead of
// 10 CTAs per SM -- 58 % slower: the spilled arrays cost more than the extra
// warps hide.  Tried: the node arrays of a warp's 32 blocks staged in shared
// memory as [node][lane], copied in and out coalesced -- 47 KB a warp at 99
// nodes leaves 4 warps per SM, and the latency of one warp per scheduler costs
// more (11.5 ms) than the bank-conflict-free accesses save (10.8 ms in place).)
template <int VCAP>
__global__ void __launch_bounds__(32)
awb_block_setup_kernel(const AwbChain *chains, int *err)
{
    const AwbChain &ch = chains[blockIdx.y];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= ch.ntrees)
        return;
    const int rc = awb_block_setup_t<VCAP>(ch, b);
    if (rc)
        atomicMax(err, 100 + rc);
}

// time-by-time matrix of every block, one thread per entry
__global__ void awb_tmatrix_kernel(const AwbChain *chains)
{
    const AwbChain &ch = chains[blockIdx.y];
    const int b = blockIdx.x;
    if (b >= ch.ntrees)
        return;
    awb_tmatrix_fill(ch, b, threadIdx.x, blockDim.x);
}

// one warp per breakpoint
__global__ void awb_switch_setup_kernel(const AwbChain *chains, int *err,
                                        int scratch_bytes)
{
    extern __shared__ unsigned char sw_smem[];
    const AwbChain &ch = chains[blockIdx.y];
    const int wpc = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b = 1 + blockIdx.x * wpc + warp;
    if (b >= ch.ntrees)
        return;
    const int rc = awb_switch_setup_warp(ch, b, lane,
                                         sw_smem + (size_t) warp * scratch_bytes);
    if (rc && lane == 0)
        atomicMax(err, 200 + rc);
}

// one warp per site; invariant / masked sites exit at once.  The warp stages
// the block's tree arrays in shared memory before the pruning passes.
__global__ void awb_emit_kernel(const AwbChain *chains, int scratch_bytes, int seg,
                                int pass)
{
    extern __shared__ unsigned char emit_smem[];
    const AwbChain &ch = chains[blockIdx.y];
    const int wpc = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int V = ch.nnodes;
    unsigned char *scratch = emit_smem + (size_t) warp * scratch_bytes;
    // staging area after the pruning scratch
    int *sparent = (int *) (scratch + ((awb_emit_scratch_bytes(V) + 15) & ~(size_t) 15));
    int *sage = sparent + V;
    short *sc0 = (short *) (sage + V);
    short *sc1 = sc0 + V;
    short *sorder = sc1 + V;
    short *slstart = sorder + V;        // [V + 2]
    int staged = -1;
    const AwbSeg g = awb_seg(ch, seg);
    if (!g.valid)
        return;
    // second pass of a checkpointed table: the last segments' tables are still
    // resident from the first pass
    if (pass == 1 && g.resident)
        return;
    // (the first column of a table is the prior / the stored first column of
    // the segment: no emission applied)
    // every warp takes a contiguous range of sites (consecutive variant sites
    // mostly share their block: the tree stays staged, and the block index only
    // moves forward), scans it 32 sites at a time for variant ones (one
    // coalesced load and a ballot; ~97 % of the sites are skipped) and works
    // through those
    const int first = g.site0 + 1, last = g.site0 + g.nsites;
    const int nw = gridDim.x * wpc;
    auto stage = [&](int b) {
        const size_t o = (size_t) b * V;
        __syncwarp();
        for (int x = lane; x < V; x += 32) {
            sparent[x] = ch.ptrees[o + x];
            sage[x] = ch.ages[o + x];
            sc0[x] = ch.child0[o + x];
            sc1[x] = ch.child1[o + x];
            sorder[x] = ch.order[o + x];
        }
        for (int x = lane; x < V + 2; x += 32)
            slstart[x] = ch.lstart[(size_t) b * (V + 2) + x];
        staged = b;
        __syncwarp();
    };
    if (!ch.seqs) {
        // the alignment as variant columns: the sites with work are listed
        // (var_pos, ascending).  The warps take the columns of the segment in
        // turn (every column costs about the same: an even share without any
        // scanning), and the column index spares the per-site search.
        int v0, v1;
        {
            const long long c0s = (long long) ch.start_coord + first;
            const long long c1s = (long long) ch.start_coord + last;
            int lo = 0, hi = ch.nvar;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (ch.var_pos[mid] < c0s) lo = mid + 1; else hi = mid; }
            v0 = lo;
            hi = ch.nvar;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (ch.var_pos[mid] < c1s) lo = mid + 1; else hi = mid; }
            v1 = lo;
        }
        int b = g.b0;
        for (int v = v0 + blockIdx.x * wpc + warp; v < v1; v += nw) {
            const int i = (int) ((long long) ch.var_pos[v] - ch.start_coord);
            if (ch.kind[i] != AWB_SITE_VARIANT)
                continue;
            // the block of site i, from the block of the previous column on
            if (ch.block_start[b + 1] <= i) {
                int lo = b + 1, hi = ch.ntrees - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (ch.block_start[mid] <= i) lo = mid; else hi = mid - 1;
                }
                b = lo;
            }
            if (b != staged)
                stage(b);
            awb_emit_site(ch, i, b, lane, 32, scratch, sparent, sage, sc0, sc1, sorder,
                          slstart, g.fwbias, v);
            __syncwarp();
        }
        return;
    }
    // dense rows: every warp takes a contiguous range of sites (consecutive
    // variant sites mostly share their block: the tree stays staged, and the
    // block index only moves forward), scans it 32 sites at a time for variant
    // ones (one coalesced load and a ballot; ~97 % of the sites are skipped) and
    // works through those
    const int per = (((last - first) + nw - 1) / nw + 31) & ~31;
    const int w0 = first + (blockIdx.x * wpc + warp) * per;
    const int w1 = w0 + per < last ? w0 + per : last;
    int b = -1;
    for (int base = w0; base < w1; base += 32) {
      const int mine = base + lane;
      unsigned vm = __ballot_sync(0xffffffffu, mine < w1 &&
                                  ch.kind[mine] == AWB_SITE_VARIANT);
      while (vm) {
        const int i = base + __ffs(vm) - 1;
        vm &= vm - 1;
        if (b < 0)
            b = awb_find_block(ch, i);
        else
            while (ch.block_start[b + 1] <= i)
                b++;
        if (b != staged)
            stage(b);
        awb_emit_site(ch, i, b, lane, 32, scratch, sparent, sage, sc0, sc1, sorder,
                      slstart, g.fwbias);
        __syncwarp();
      }
    }
}

// Unphased data, after the traceback: one warp per site range, like the emission
// kernel; for every heterozygous variant site the two phasings are evaluated at
// the sampled state alone.  out[site] = e1 / (e1 + e2), -1 elsewhere.
__global__ void awb_phase_kernel(const AwbChain *chains, int scratch_bytes, double *out,
                                 const long long *out_off)
{
    extern __shared__ unsigned char emit_smem[];
    const AwbChain &ch = chains[blockIdx.y];
    const int wpc = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int V = ch.nnodes;
    unsigned char *scratch = emit_smem + (size_t) warp * scratch_bytes;
    int *sparent = (int *) (scratch + ((awb_emit_scratch_bytes(V) + 15) & ~(size_t) 15));
    int *sage = sparent + V;
    short *sc0 = (short *) (sage + V);
    short *sc1 = sc0 + V;
    short *sorder = sc1 + V;
    short *slstart = sorder + V;        // [V + 2]
    int staged = -1;
    double *o = out + out_off[blockIdx.y];
    const int first = 0, last = ch.nsites;
    const int nw = gridDim.x * wpc;
    const int per = (((last - first) + nw - 1) / nw + 31) & ~31;
    const int w0 = first + (blockIdx.x * wpc + warp) * per;
    const int w1 = w0 + per < last ? w0 + per : last;
    int b = -1;
    for (int base = w0; base < w1; base += 32) {
      const int mine = base + lane;
      if (mine < w1)
          o[mine] = -1.0;
      unsigned vm = __ballot_sync(0xffffffffu, mine < w1 &&
                                  ch.kind[mine] == AWB_SITE_VARIANT);
      while (vm) {
        const int i = base + __ffs(vm) - 1;
        vm &= vm - 1;
        if (!awb_site_het(ch, i))
            continue;
        if (b < 0)
            b = awb_find_block(ch, i);
        else
            while (ch.block_start[b + 1] <= i)
                b++;
        if (ch.nstates[b] == 0)
            continue;
        if (b != staged) {
            const size_t ob = (size_t) b * V;
            __syncwarp();
            for (int x = lane; x < V; x += 32) {
                sparent[x] = ch.ptrees[ob + x];
                sage[x] = ch.ages[ob + x];
                sc0[x] = ch.child0[ob + x];
                sc1[x] = ch.child1[ob + x];
                sorder[x] = ch.order[ob + x];
            }
            for (int x = lane; x < V + 2; x += 32)
                slstart[x] = ch.lstart[(size_t) b * (V + 2) + x];
            staged = b;
            __syncwarp();
        }
        const double pr = awb_phase_prob(ch, i, b, ch.path[i], lane, 32, scratch, sparent,
                                         sage, sc0, sc1, sorder, slstart);
        __syncwarp();
        if (lane == 0)
            o[i] = pr;
      }
    }
}

// one warp per window (awb_recomb.cuh)
__global__ void awb_recomb_kernel(const AwbChain *chains, const int *rng_states,
                                  int rand_max, int *out, const long long *out_off,
                                  int *info)
{
    const AwbChain &ch = chains[blockIdx.x];
    const int *st = rng_states + (size_t) blockIdx.x * AWB_RNG_WORDS;
    AwbRng rng;
    for (int i = 0; i < 31; i++)
        rng.r[i] = st[i];
    rng.f = st[31];
    rng.b = st[32];
    const int cap = ch.nsites;
    int *o = out + out_off[blockIdx.x];
    awb_sample_recombs(ch, rng, rand_max, o, o + cap, o + 2 * (size_t) cap, cap,
                       info + 2 * blockIdx.x, (int) threadIdx.x);
}

// ---------------------------------------------------------------- host side

#define AWB_UPLOAD_GROUPS 4

struct awb_ctx {
    int device;
    cudaStream_t stream;
    // inputs go up on their own stream, in groups of problems, so that the
    // setup kernels of one group overlap with the copies of the next
    cudaStream_t copy_stream;
    cudaEvent_t up_ev[AWB_UPLOAD_GROUPS], seq_ev, rand_ev, order_ev;
    cudaEvent_t ev[6];
    cudaEvent_t user_ev[8];
    int sm_count;
    // device arena kept across batches (cudaMalloc / cudaFree of a table-sized
    // arena cost tens to hundreds of milliseconds); lent to one batch at a time
    char *arena_cache;
    size_t arena_cap;
    bool arena_busy;
    // pinned staging buffer for the arrays the host layout makes (a copy from
    // pageable memory blocks the calling thread until the stream has drained);
    // lent together with the arena
    char *stage;
    size_t stage_cap;
};

struct awb_batch {
    awb_ctx *ctx;
    int C;
    std::vector<AwbLayout> L;
    std::vector<awb_problem> P;
    std::vector<size_t> arena_off;
    std::vector<AwbChain> h_chains;
    AwbChain *stage_chains;      // pinned copy of h_chains (or NULL)
    size_t windows_bytes;        // arena bytes of the windows (before the chain records)
    bool with_band;              // the generic kernel's tables are part of the arena
    int nslots;                  // checkpointed table: segment tables per window
    int nrot;                    //   ... and segments the second pass rebuilds at once
    bool bound;                  // batch_bind has run
    char *arena;
    size_t arena_bytes;
    AwbChain *d_chains;
    int *d_err;
    int maxB, maxn, maxS, maxV, maxT, maxband, maxNS, maxcnt, zcap;
    float ms[3];
    int launches;
    int64_t h2d_bytes;
    bool uploaded, setup_done, forward_done, rand_uploaded, rand_in_use;
    bool ckpt;                 // checkpointed forward table (AWB_CHECKPOINT)
    bool any_packed;           // some problem gives its alignment as variant columns
    int maxnvar;
    bool lin_unsafe;           // some problem's linear-domain vectors may overflow
                               //   (awb_layout.h): generic forward kernel, closed-form
                               //   transitions in the traceback
    bool tables_stale;         // checkpointed table: a traceback has rebuilt segments
                               //   over the tables the forward pass left resident
    int maxseg;                // most segments of any problem
    int maxsegsites;           // most sites of any segment
    // recombination points (awb_batch_sample_recombs): [pos | node | time] per
    // window, the (count, draws) pairs and the generator states
    int *d_rec;
    long long *d_rec_off;
    int *d_rec_info, *d_rng;
    double *d_phase;             // awb_batch_phase_probs: [sum of nsites]
    long long *d_phase_off;
    std::vector<long long> phase_off;
    std::vector<long long> rec_off;
    bool recombs_done;
    // per-kernel device times (awb_batch_kernel_times): event pairs around the
    // launches of each kernel class, summed when they are read
    bool ktimes;
    std::vector<cudaEvent_t> kev;      // start, stop, start, stop, ...
    std::vector<int> kclass;
    size_t kev_used;
    double k4_bytes;                   // algorithmic bytes of the forward launches
    int k4_launches;
};

enum { AWB_K_KIND = 0, AWB_K_BLOCK, AWB_K_TMATRIX, AWB_K_SWITCH, AWB_K_EMIT,
       AWB_K_FORWARD, AWB_K_TRACEBACK, AWB_K_RECOMB, AWB_K_NCLASS };

// brackets one launch (or a run of launches of one class) with events
struct KTimer {
    awb_batch *b;
    bool on;
    cudaStream_t st;
    size_t at;
    KTimer(awb_batch *b_, int cls, cudaStream_t st_ = 0) : b(b_), on(b_->ktimes) {
        st = st_ ? st_ : b->ctx->stream;
        if (!on) return;
        if (b->kev_used + 2 > b->kev.size()) {
            cudaEvent_t e0, e1;
            if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) {
                on = false;
                return;
            }
            b->kev.push_back(e0);
            b->kev.push_back(e1);
        }
        b->kclass.push_back(cls);
        at = b->kev_used;
        b->kev_used += 2;
        cudaEventRecord(b->kev[at], st);
    }
    ~KTimer() {
        if (!on) return;
        cudaEventRecord(b->kev[at + 1], st);
    }
};

extern "C" int awb_ctx_create(int device, awb_ctx **out)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev)
        return fail("bad device ordinal");
    CUDA_OK(cudaSetDevice(device));
    awb_ctx *ctx = new awb_ctx;
    ctx->device = device;
    CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));

    for (int i = 0; i < AWB_UPLOAD_GROUPS; i++)
        CUDA_OK(cudaEventCreateWithFlags(&ctx->up_ev[i], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ctx->order_ev, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ctx->seq_ev, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ctx->rand_ev, cudaEventDisableTiming));
    for (int i = 0; i < 6; i++)
        CUDA_OK(cudaEventCreate(&ctx->ev[i]));
    for (int i = 0; i < 8; i++)
        CUDA_OK(cudaEventCreate(&ctx->user_ev[i]));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    ctx->arena_cache = NULL;
    ctx->arena_cap = 0;
    ctx->arena_busy = false;
    ctx->stage = NULL;
    ctx->stage_cap = 0;
    CUDA_OK(cudaFuncSetAttribute(awb_forward_kernel<1>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
    CUDA_OK(cudaFuncSetAttribute(awb_forward_kernel<2>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
    CUDA_OK(cudaFuncSetAttribute(awb_switch_setup_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
    CUDA_OK(cudaFuncSetAttribute(awb_emit_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
    *out = ctx;
    return 0;
}

extern "C" void awb_ctx_destroy(awb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < 6; i++)
        cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 8; i++)
        cudaEventDestroy(ctx->user_ev[i]);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->copy_stream);

    for (int i = 0; i < AWB_UPLOAD_GROUPS; i++)
        cudaEventDestroy(ctx->up_ev[i]);
    cudaEventDestroy(ctx->order_ev);
    cudaEventDestroy(ctx->seq_ev);
    cudaEventDestroy(ctx->rand_ev);
    if (ctx->arena_cache) cudaFree(ctx->arena_cache);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    delete ctx;
}

extern "C" int awb_ctx_record(awb_ctx *ctx, int slot)
{
    if (slot < 0 || slot >= 8) return fail("bad event slot");
    CUDA_OK(cudaSetDevice(ctx->device));
    CUDA_OK(cudaEventRecord(ctx->user_ev[slot], ctx->stream));
    return 0;
}

extern "C" int awb_ctx_elapsed_ms(awb_ctx *ctx, int slot0, int slot1, float *ms)
{
    if (slot0 < 0 || slot0 >= 8 || slot1 < 0 || slot1 >= 8)
        return fail("bad event slot");
    CUDA_OK(cudaSetDevice(ctx->device));
    CUDA_OK(cudaEventSynchronize(ctx->user_ev[slot1]));
    CUDA_OK(cudaEventElapsedTime(ms, ctx->user_ev[slot0], ctx->user_ev[slot1]));
    return 0;
}

extern "C" int awb_ctx_sync(awb_ctx *ctx)
{
    CUDA_OK(cudaSetDevice(ctx->device));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

extern "C" void awb_batch_destroy(awb_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    if (b->arena) {
        // work queued on the streams may still touch the arena (and the copies
        // read host arrays owned by this batch)
        cudaStreamSynchronize(b->ctx->copy_stream);
        cudaStreamSynchronize(b->ctx->stream);
        if (b->arena == b->ctx->arena_cache)
            b->ctx->arena_busy = false;
        else
            cudaFree(b->arena);
    }
    for (size_t i = 0; i < b->kev.size(); i++)
        cudaEventDestroy(b->kev[i]);
    if (b->d_rec) cudaFree(b->d_rec);
    if (b->d_rec_off) cudaFree(b->d_rec_off);
    if (b->d_rec_info) cudaFree(b->d_rec_info);
    if (b->d_rng) cudaFree(b->d_rng);
    if (b->d_phase) cudaFree(b->d_phase);
    if (b->d_phase_off) cudaFree(b->d_phase_off);
    delete b;
}

// Shape of the register-resident forward kernel (awb_forward_fast.cuh) for a
// batch: TMAX (time rows), U (consecutive states per compute thread), compute
// threads NT.  U is chosen so that a window needs at most four compute warps
// where it can (few warps per window = several windows per SM); a branch of more
// than 32 states (possible with more than 33 time points) needs U >= 2.
struct FastShape {
    int tmax, U, NT, threads;
    int bookw;          // 1: a warp of its own keeps the books (one window per SM)
    bool ok;
};

// dense: more CTAs than SMs in the launch -- few warps per window, several
// windows per SM; else: one window per SM, many warps per window (the per-site
// chain is shorter with one state per thread)
static FastShape batch_fast_shape(const awb_batch *b, bool dense = true)
{
    FastShape f;
    const int Tm1 = b->maxT - 1;
    f.tmax = Tm1 <= 20 ? 20 : (Tm1 <= 40 ? 40 : 64);
    if (b->maxNS <= 128 && b->maxcnt <= 32)
        f.U = 1;
    else if (b->maxNS <= 256)
        f.U = 2;
    else
        f.U = 4;
    if (!dense) {
        if (b->maxNS <= 512 && b->maxcnt <= 32)
            f.U = 1;
        else if (b->maxNS <= 1024)
            f.U = 2;
    }
    if (getenv("AWB_K4_U")) {
        const int u = atoi(getenv("AWB_K4_U"));
        if ((u == 1 && b->maxcnt <= 32 && b->maxNS <= 512) || u == 2 || u == 4)
            f.U = u;
    }
    f.NT = (b->maxNS + 32 * f.U - 1) / (32 * f.U) * 32;
    f.bookw = (dense || getenv("AWB_K4_NOBOOKW")) ? 0 : 1;
    f.threads = f.NT + AWB_FWD_HELPERS + 32 * f.bookw;
    f.ok = !getenv("AWB_FORCE_GENERIC") && !b->lin_unsafe && b->maxcnt <= 64 &&
        f.threads <= 608 &&
        awb_fwd_fast_smem_bytes(f.NT * f.U, f.tmax, b->zcap) <= 200 * 1024;
    return f;
}

static bool batch_fast_path(const awb_batch *b)
{
    return batch_fast_shape(b).ok;
}

extern "C" int awb_batch_create(awb_ctx *ctx, int nproblems,
                                const awb_problem *problems, int flags,
                                awb_batch **out)
{
    if (!ctx) return fail("null context");
    if (nproblems < 1) return fail("need at least one problem");
    CUDA_OK(cudaSetDevice(ctx->device));
    awb_batch *b = new awb_batch;
    b->ctx = ctx;
    b->C = nproblems;
    b->arena = NULL;
    b->d_chains = NULL;
    b->d_err = NULL;
    b->L.resize(nproblems);
    b->P.assign(problems, problems + nproblems);
    b->arena_off.resize(nproblems);
    b->h_chains.resize(nproblems);
    b->maxB = b->maxn = b->maxS = b->maxV = b->maxT = b->maxband = 0;
    b->maxNS = 32;
    b->zcap = 64;
    b->maxcnt = 1;
    b->ms[0] = b->ms[1] = b->ms[2] = 0;
    b->launches = 0;
    b->h2d_bytes = 0;
    b->uploaded = b->setup_done = b->forward_done = false;
    b->rand_uploaded = false;
    b->rand_in_use = false;
    b->lin_unsafe = false;
    b->any_packed = false;
    b->maxnvar = 0;
    b->tables_stale = false;
    b->ckpt = false;
    b->d_rec = NULL;
    b->d_rec_off = NULL;
    b->d_rec_info = b->d_rng = NULL;
    b->recombs_done = false;
    b->d_phase = NULL;
    b->d_phase_off = NULL;
    b->ktimes = false;
    b->kev_used = 0;
    b->k4_bytes = 0;
    b->k4_launches = 0;

    const auto t_create0 = std::chrono::steady_clock::now();
    // host layout of every problem (integer work), one host thread per problem
    {
        std::vector<std::string> errs(nproblems);
        std::vector<char> ok(nproblems, 0);
        const int keep = (flags & AWB_KEEP_DEBUG) ? 1 : 0;
        // checkpointed table: segments of at most 2^23 doubles (64 MiB) per problem
        // (small enough for half a dozen tables per window next to the per-block
        // tables, so that the second pass can rebuild three segments at once)
        long long seg_cap = (flags & AWB_CHECKPOINT) ? (1ll << 23) : 0;
        if (seg_cap && getenv("AWB_SEG_DOUBLES"))
            seg_cap = atoll(getenv("AWB_SEG_DOUBLES"));
        const int nthreads = host_threads(nproblems);
        auto work = [&](int t) {
            for (int c = t; c < nproblems; c += nthreads)
                ok[c] = awb_layout_build(problems[c], keep, b->L[c], errs[c],
                                         seg_cap) ? 1 : 0;
        };
        if (nthreads <= 1) {
            work(0);
        } else {
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; t++) pool.emplace_back(work, t);
            for (auto &th : pool) th.join();
        }
        for (int c = 0; c < nproblems; c++) {
            if (!ok[c]) {
                std::string msg = "problem " + std::to_string(c) + ": " + errs[c];
                delete b;
                return fail(msg);
            }
        }
    }
    for (int c = 0; c < nproblems; c++) {
        const AwbLayout &L = b->L[c];
        if (L.B > b->maxB) b->maxB = L.B;
        if (L.n > b->maxn) b->maxn = L.n;
        if (L.maxS > b->maxS) b->maxS = L.maxS;
        if (L.V > b->maxV) b->maxV = L.V;
        if (L.T > b->maxT) b->maxT = L.T;
        if (L.maxband > b->maxband) b->maxband = L.maxband;
        if (L.maxNS > b->maxNS) b->maxNS = L.maxNS;
        if (L.zcap > b->zcap) b->zcap = L.zcap;
        if (L.maxcnt > b->maxcnt) b->maxcnt = L.maxcnt;
        if (L.lin_unsafe) b->lin_unsafe = true;
        if (L.packed) {
            b->any_packed = true;
            if (problems[c].nvar > b->maxnvar) b->maxnvar = problems[c].nvar;
        }
        for (size_t i = 0; i < L.copies.size(); i++)
            b->h2d_bytes += (int64_t) L.copies[i].bytes;
    }
    b->ckpt = (flags & AWB_CHECKPOINT) != 0;
    b->maxseg = 1;
    b->maxsegsites = b->maxn;
    if (b->ckpt) {
        if (!batch_fast_path(b)) {
            delete b;
            return fail("AWB_CHECKPOINT needs a state space the fast forward kernel covers");
        }
        b->maxsegsites = 1;
        for (int c = 0; c < nproblems; c++) {
            if (b->L[c].nseg > b->maxseg) b->maxseg = b->L[c].nseg;
            if (b->L[c].seg_sites > b->maxsegsites) b->maxsegsites = b->L[c].seg_sites;
        }
    }
    // the tmatrix2 band is only read by the generic forward kernel
    b->with_band = !batch_fast_path(b) || (flags & AWB_KEEP_DEBUG);
    b->windows_bytes = 0;
    b->nslots = 1;
    b->nrot = 1;
    b->bound = false;
    if (getenv("AWB_VERBOSE"))
        fprintf(stderr, "awb_batch_create: layout %.1f ms\n",
                std::chrono::duration<double, std::milli>(
                    std::chrono::steady_clock::now() - t_create0).count());
    *out = b;
    return 0;
}

// Second half of the batch's construction, at the first awb_batch_upload: the
// device arena (the context's cached one when it is free), the chain records,
// the pinned staging copies.  awb_batch_create itself is host-only work, so the
// layout of the next batch can be made while the previous batch still runs on
// (and owns the arena of) the device.
static int batch_bind(awb_batch *b)
{
    awb_ctx *ctx = b->ctx;
    const int nproblems = b->C;
    const auto t_create1 = std::chrono::steady_clock::now();
    // arena bytes of every window.  Checkpointed table: as many segment tables
    // per window as the device has room for (AWB_RESIDENT_SEGS overrides) --
    // the traceback then finds the last ones still resident and rebuilds fewer
    int nslots = 1;
    if (b->ckpt) {
        size_t base = 0, slot = 0;
        for (int c = 0; c < nproblems; c++) {
            base += awb_align(awb_layout_place_slots(b->L[c], 1, b->with_band));
            slot += awb_layout_slot_bytes(b->L[c]);
        }
        size_t free_b = 0, total_b = 0;
        // (a batch next to another live one gets a plain allocation: one table)
        if (!ctx->arena_busy && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            free_b += ctx->arena_cap;                   // ours to re-use or replace
            const size_t margin = (size_t) 8 << 30;     // kernels' local memory, NCCL, ...
            if (free_b > base + margin && slot > 0)
                nslots = 1 + (int) ((free_b - base - margin) / slot);
        } else {
            cudaGetLastError();
        }
        if (getenv("AWB_RESIDENT_SEGS"))
            nslots = atoi(getenv("AWB_RESIDENT_SEGS"));
        if (nslots > b->maxseg) nslots = b->maxseg;
        if (nslots < 1) nslots = 1;
    }
    size_t total = 0, chains_off = 0, err_off = 0;
    for (;;) {
        total = 0;
        for (int c = 0; c < nproblems; c++) {
            b->arena_off[c] = total;
            if (b->ckpt)
                total += awb_align(awb_layout_place_slots(b->L[c], nslots, b->with_band));
            else
                total += awb_align(b->with_band ? b->L[c].total_bytes
                                                : b->L[c].bytes_before_band);
        }
        b->windows_bytes = total;
        b->nslots = nslots;
        // as many segments side by side as the dense forward kernel puts CTAs on
        // an SM (and as there are tables); fewer than three tables: one
        b->nrot = 1;
        if (nslots >= 3) {
            const FastShape f = batch_fast_shape(b, true);
            const int regs = (f.threads <= 192 && f.U <= 2) ? 80 : 96;
            int want = 4;
            for (; want > 1; want--)
                if ((want * (f.threads / 32) + 3) / 4 * 32 * regs <= 16384)
                    break;
            if (want < 2) want = 2;
            b->nrot = want < nslots ? want : nslots;
        }
        // chain records and the error word live at the tail of the arena
        chains_off = total;
        total += awb_align(sizeof(AwbChain) * nproblems);
        err_off = total;
        total += awb_align(sizeof(int));
        b->arena_bytes = total;
        cudaError_t e = cudaSuccess;
        if (!ctx->arena_busy) {
            if (ctx->arena_cap < total) {
                if (ctx->arena_cache) cudaFree(ctx->arena_cache);
                ctx->arena_cache = NULL;
                ctx->arena_cap = 0;
                e = cudaMalloc((void **) &ctx->arena_cache, total);
                if (e == cudaSuccess)
                    ctx->arena_cap = total;
            }
            if (e == cudaSuccess) {
                b->arena = ctx->arena_cache;
                ctx->arena_busy = true;
            }
        } else {
            e = cudaMalloc((void **) &b->arena, total);
            if (e != cudaSuccess)
                b->arena = NULL;
        }
        if (e == cudaSuccess)
            break;
        cudaGetLastError();
        if (nslots > 1) {
            nslots = 1;         // the estimate of the free memory was too generous
            continue;
        }
        return fail(std::string("cudaMalloc of ") + std::to_string(total) +
                    " bytes failed: " + cudaGetErrorString(e));
    }
    b->d_chains = (AwbChain *) (b->arena + chains_off);
    b->d_err = (int *) (b->arena + err_off);
    for (int c = 0; c < nproblems; c++) {
        awb_layout_bind(b->L[c], b->P[c], b->arena + b->arena_off[c],
                        b->h_chains[c]);
        b->h_chains[c].nrot = b->nrot;
        b->h_chains[c].need_band = batch_fast_path(b) ? 0 : 1;
    }
    b->stage_chains = NULL;
    if (b->arena == ctx->arena_cache) {
        // stage the layout's own arrays (and the chain records) in pinned memory
        std::vector<size_t> soff(nproblems + 1, 0);
        for (int c = 0; c < nproblems; c++) {
            size_t need = 0;
            for (size_t i = 0; i < b->L[c].copies.size(); i++)
                if (b->L[c].copies[i].own)
                    need += awb_align(b->L[c].copies[i].bytes);
            soff[c + 1] = soff[c] + need;
        }
        const size_t need = soff[nproblems] + awb_align(sizeof(AwbChain) * nproblems);
        if (ctx->stage_cap < need) {
            if (ctx->stage) cudaFreeHost(ctx->stage);
            ctx->stage = NULL;
            ctx->stage_cap = 0;
            if (cudaHostAlloc((void **) &ctx->stage, need + need / 8,
                              cudaHostAllocDefault) == cudaSuccess)
                ctx->stage_cap = need + need / 8;
            else
                cudaGetLastError();     // no staging: the copies work without it
        }
        if (ctx->stage) {
            const int nthreads = host_threads(nproblems);
            auto work = [&](int t) {
                for (int c = t; c < nproblems; c += nthreads) {
                    char *dst = ctx->stage + soff[c];
                    std::vector<AwbCopy> merged;
                    for (size_t i = 0; i < b->L[c].copies.size(); i++) {
                        AwbCopy cp = b->L[c].copies[i];
                        if (!cp.own) {
                            merged.push_back(cp);
                            continue;
                        }
                        memcpy(dst, cp.src, cp.bytes);
                        cp.src = dst;
                        dst += awb_align(cp.bytes);
                        // neighbours in the arena are neighbours here: one copy
                        // (the alignment gaps go along)
                        if (!merged.empty() && merged.back().own &&
                            cp.dst_off == awb_align(merged.back().dst_off +
                                                    merged.back().bytes) &&
                            (const char *) cp.src == (const char *) merged.back().src +
                                awb_align(merged.back().bytes))
                            merged.back().bytes = cp.dst_off - merged.back().dst_off + cp.bytes;
                        else
                            merged.push_back(cp);
                    }
                    b->L[c].copies.swap(merged);
                }
            };
            if (nthreads <= 1) {
                work(0);
            } else {
                std::vector<std::thread> pool;
                for (int t = 0; t < nthreads; t++) pool.emplace_back(work, t);
                for (auto &th : pool) th.join();
            }
            b->stage_chains = (AwbChain *) (ctx->stage + soff[nproblems]);
            memcpy(b->stage_chains, b->h_chains.data(), sizeof(AwbChain) * nproblems);
        }
    }
    if (getenv("AWB_VERBOSE")) {
        const auto t_create2 = std::chrono::steady_clock::now();
        fprintf(stderr, "awb_batch_upload: arena (%.2f GB, %d segment tables per window) + bind + staging %.1f ms\n",
                total / 1e9, b->nslots,
                std::chrono::duration<double, std::milli>(t_create2 - t_create1).count());
    }
    b->bound = true;
    return 0;
}

extern "C" int64_t awb_batch_h2d_bytes(const awb_batch *b)
{
    return b->h2d_bytes + (int64_t) sizeof(AwbChain) * b->C;
}

// problems [g0, g1) of upload group g
static void upload_group(const awb_batch *b, int g, int &g0, int &g1)
{
    const int ng = b->C >= 2 * AWB_UPLOAD_GROUPS ? AWB_UPLOAD_GROUPS : 1;
    if (g >= ng) { g0 = g1 = b->C; return; }
    g0 = (int) ((long long) b->C * g / ng);
    g1 = (int) ((long long) b->C * (g + 1) / ng);
}

extern "C" int awb_batch_upload(awb_batch *b)
{
    CUDA_OK(cudaSetDevice(b->ctx->device));
    if (!b->bound && batch_bind(b))
        return 1;
    cudaStream_t st = b->ctx->copy_stream;
    // after whatever the compute stream still has queued on this arena
    CUDA_OK(cudaEventRecord(b->ctx->order_ev, b->ctx->stream));
    CUDA_OK(cudaStreamWaitEvent(st, b->ctx->order_ev, 0));
    CUDA_OK(cudaMemcpyAsync(b->d_chains,
                            b->stage_chains ? b->stage_chains : b->h_chains.data(),
                            sizeof(AwbChain) * b->C, cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(b->d_err, 0, sizeof(int), st));
    // the trees first, group by group (the per-block setup kernels start on a
    // group as soon as its trees are there), then the sequences, which are 2/3
    // of the bytes and are not read before the site-kind kernel
    for (int g = 0; g < AWB_UPLOAD_GROUPS; g++) {
        int g0, g1;
        upload_group(b, g, g0, g1);
        for (int c = g0; c < g1; c++) {
            const AwbLayout &L = b->L[c];
            char *base = b->arena + b->arena_off[c];
            // debug arrays are compared entry by entry, including the entries no
            // kernel writes (block 0 has no switch matrix): start them from zero
            // (band included; not the draws, which awb_batch_upload_rand may
            // already have put there)
            if (L.keep_debug) {
                CUDA_OK(cudaMemsetAsync(base, 0, L.o_rand, st));
                CUDA_OK(cudaMemsetAsync(base + L.o_logz, 0, L.total_bytes - L.o_logz, st));
            }
            for (size_t i = 0; i < L.copies.size(); i++)
                if (L.copies[i].dst_off != L.o_seqs)
                    CUDA_OK(cudaMemcpyAsync(base + L.copies[i].dst_off, L.copies[i].src,
                                            L.copies[i].bytes, cudaMemcpyHostToDevice,
                                            st));
        }
        CUDA_OK(cudaEventRecord(b->ctx->up_ev[g], st));
    }
    for (int c = 0; c < b->C; c++) {
        const AwbLayout &L = b->L[c];
        char *base = b->arena + b->arena_off[c];
        for (size_t i = 0; i < L.copies.size(); i++)
            if (L.copies[i].dst_off == L.o_seqs)
                CUDA_OK(cudaMemcpyAsync(base + L.copies[i].dst_off, L.copies[i].src,
                                        L.copies[i].bytes, cudaMemcpyHostToDevice,
                                        st));
    }
    CUDA_OK(cudaEventRecord(b->ctx->seq_ev, st));
    b->uploaded = true;
    b->setup_done = b->forward_done = false;
    return 0;
}

// ---- kernel launchers (seg: segment of a checkpointed table, else 0)

static int launch_emit(awb_batch *b, int seg, int pass, cudaStream_t st = 0)
{
    if (!st) st = b->ctx->stream;
    const int scratch = (int) (((awb_emit_scratch_bytes(b->maxV) + 15) & ~(size_t) 15) +
                               (((size_t) b->maxV * 16 + 4 + 15) & ~(size_t) 15));
    int wpc = 8;
    while (wpc > 1 && (size_t) wpc * scratch > 160 * 1024)
        wpc >>= 1;
    if ((size_t) wpc * scratch > 200 * 1024)
        return fail("tree too large for the emission kernel's shared memory");
    int gx = (b->maxsegsites + wpc - 1) / wpc;
    const int cap = b->ctx->sm_count * 16;
    if (gx > cap) gx = cap;
    if (b->ckpt && gx * b->C > cap)
        gx = (cap + b->C - 1) / b->C;
    dim3 grid(gx, b->C);
    KTimer kt(b, AWB_K_EMIT, st);
    awb_emit_kernel<<<grid, 32 * wpc, (size_t) wpc * scratch, st>>>(
        b->d_chains, scratch, seg, pass);
    b->launches++;
    return 0;
}

// nsub: segments per chain worked on at once (seg, seg-1, ...): only the second
// pass of a checkpointed table has independent segments
static int launch_forward_fast(awb_batch *b, int seg, int pass, int nsub = 1,
                               cudaStream_t st = 0)
{
    if (!st) st = b->ctx->stream;
    const bool dense = (long long) b->C * nsub > b->ctx->sm_count;
    FastShape f = batch_fast_shape(b, dense);
    if (!f.ok)
        f = batch_fast_shape(b);
    const int threads = f.threads;
    const size_t fsmem = awb_fwd_fast_smem_bytes(f.NT * f.U, f.tmax, b->zcap);
    const dim3 grid(b->C, nsub);
    // CTAs per SM the register files allow (four files of 16 K registers, a
    // CTA's warps dealt out over them)
    int want = 1;
    if (dense) {
        const int regs = (threads <= 192 && f.U <= 2 || getenv("AWB_K4_LEAN")) ? 80 : 96;
        for (want = 4; want > 1; want--)
            if ((want * (threads / 32) + 3) / 4 * 32 * regs <= 16384)
                break;
    }
    int carve = (int) (((fsmem + 1024) * want * 100 + 228 * 1024 - 1) / (228 * 1024)) + 1;
    if (carve > 100) carve = 100;
    static bool said = false;
    const bool verbose = getenv("AWB_VERBOSE") && !said;
    said = said || verbose;
#define AWB_LAUNCH_FAST(TM, NL, MT, UU, MB)                                      \
    do {                                                                         \
        if (fsmem > 48 * 1024)                                                   \
            CUDA_OK(cudaFuncSetAttribute(awb_forward_fast_kernel<TM, NL, MT, UU, MB>, \
                cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fsmem));      \
        /* shared memory for the CTAs that are to share an SM, the rest of the  \
           256 KB stays L1 (the kernel lives on its L1 prefetches) */            \
        CUDA_OK(cudaFuncSetAttribute(awb_forward_fast_kernel<TM, NL, MT, UU, MB>, \
            cudaFuncAttributePreferredSharedMemoryCarveout, carve));             \
        if (verbose) {                                                           \
            int nb = 0;                                                          \
            cudaFuncAttributes fa;                                               \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb,                   \
                awb_forward_fast_kernel<TM, NL, MT, UU, MB>, threads, fsmem);    \
            cudaFuncGetAttributes(&fa, awb_forward_fast_kernel<TM, NL, MT, UU, MB>); \
            fprintf(stderr, "forward kernel <%d,%d,%d,%d,%d>: %d threads, %zu B shared, " \
                    "%d registers, %zu B local, carve-out %d %%: %d CTAs per SM\n", TM, NL, MT, UU, MB, \
                    threads, fsmem, fa.numRegs, (size_t) fa.localSizeBytes, carve, nb); \
        }                                                                        \
        awb_forward_fast_kernel<TM, NL, MT, UU, MB><<<grid, threads, fsmem, st>>>( \
            b->d_chains, seg, pass, b->zcap);                                    \
    } while (0)
    // (MAXTHREADS, registers per thread: 7 warps at 72 / 96 registers = 4 / 3
    // CTAs per SM, 11 warps at 88 = 2)
    // N1 / N2 / N4: levels of the carry scans for U = 1 / 2 / 4 (a branch of
    // maxcnt states touches (maxcnt + U - 2) / U + 1 lanes).
    // Registers per thread: an SM has four register files of 16 K registers and
    // a CTA's warps are dealt out over them, so n CTAs of w warps need
    // ceil(n w / 4) warps' worth of registers per file: 6-warp CTAs (4 compute
    // warps + the scribes) run 3 per SM at 96 registers, 4 at 80.
    const bool lean = getenv("AWB_K4_LEAN") != NULL;
#define AWB_LAUNCH_FAST_T(TM, N1, N2, N4)                                        \
    do {                                                                         \
        if (f.bookw) {                                                           \
            if (f.U == 1) AWB_LAUNCH_FAST(TM, N1, 1, 1, 96);                     \
            else if (f.U == 2) AWB_LAUNCH_FAST(TM, N2, 1, 2, 96);                \
            else AWB_LAUNCH_FAST(TM, N4, 1, 4, 96);                              \
        } else if (threads <= 192) {                                             \
            if (f.U == 1) AWB_LAUNCH_FAST(TM, N1, 0, 1, 80);                     \
            else if (f.U == 2) AWB_LAUNCH_FAST(TM, N2, 0, 2, 80);                \
            else if (lean) AWB_LAUNCH_FAST(TM, N4, 0, 4, 80);                    \
            else AWB_LAUNCH_FAST(TM, N4, 0, 4, 96);                              \
        } else {                                                                 \
            if (f.U == 1) AWB_LAUNCH_FAST(TM, N1, 0, 1, 96);                     \
            else if (f.U == 2) AWB_LAUNCH_FAST(TM, N2, 0, 2, 96);                \
            else AWB_LAUNCH_FAST(TM, N4, 0, 4, 96);                              \
        }                                                                        \
    } while (0)
    if (b->ktimes) {
        // algorithmic bytes of this launch: 8 B per site*state of the segments
        // it computes
        for (int c = 0; c < b->C; c++) {
            const AwbLayout &L = b->L[c];
            for (int y = 0; y < nsub; y++) {
                const int sg = seg - y;
                long long d0, d1;
                if (!b->ckpt) {
                    if (y > 0) continue;
                    d0 = 0;
                    d1 = L.fw_off[L.B];
                } else {
                    if (sg < 0 || sg >= L.nseg) continue;
                    const int R = b->nslots < L.nseg ? b->nslots : L.nseg;
                    if (pass == 1 && sg >= L.nseg - R) continue;     // still resident
                    const int b0 = L.seg_start[sg], b1 = L.seg_start[sg + 1];
                    d0 = L.fw_off[b0];
                    d1 = L.fw_off[b1] + (b1 < L.B ? (L.nstates[b1] > 0 ? L.nstates[b1] : 1) : 0);
                }
                b->k4_bytes += 8.0 * (double) (d1 - d0);
            }
        }
        b->k4_launches++;
    }
    KTimer kt(b, AWB_K_FORWARD, st);
    if (f.tmax == 20) AWB_LAUNCH_FAST_T(20, 5, 4, 3);
    else if (f.tmax == 40) AWB_LAUNCH_FAST_T(40, 5, 5, 4);
    else AWB_LAUNCH_FAST_T(64, 5, 5, 4);
#undef AWB_LAUNCH_FAST_T
#undef AWB_LAUNCH_FAST
    b->launches++;
    return 0;
}

static int launch_traceback(awb_batch *b, int rand_max, int seg)
{
    cudaStream_t st = b->ctx->stream;
    const int maxS1 = b->maxS > 0 ? b->maxS : 1;
    const int maxent = maxS1 + b->maxT + 4;        // capacity per block (awb_layout.h)
    const size_t smem = awb_tb_smem_bytes(maxS1, b->maxT, maxent);
    if (smem > 220 * 1024)
        return fail("state space too large for the traceback kernel's shared memory");
#define AWB_LAUNCH_TB(NV, SPW, VPT) do { \
        CUDA_OK(cudaFuncSetAttribute(awb_traceback_kernel<NV, SPW, VPT>, \
            cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
        awb_traceback_kernel<NV, SPW, VPT><<<b->C, AWB_TB_THREADS, smem, st>>>( \
            b->d_chains, rand_max, maxS1, b->maxT, maxent, seg, \
            b->lin_unsafe ? 1 : 0); } while (0)
    KTimer kt(b, AWB_K_TRACEBACK);
    if (maxS1 <= 128) AWB_LAUNCH_TB(4, 2, 1);
    else if (maxS1 <= 256) AWB_LAUNCH_TB(8, 2, 1);
    else if (maxS1 <= 512) AWB_LAUNCH_TB(16, 2, 1);
    else if (maxS1 <= 1024) AWB_LAUNCH_TB(32, 1, 2);
    else AWB_LAUNCH_TB(64, 1, 4);
#undef AWB_LAUNCH_TB
    b->launches++;
    return 0;
}

extern "C" int awb_batch_setup(awb_batch *b)
{
    if (!b->uploaded) return fail("awb_batch_setup: inputs not uploaded");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    CUDA_OK(cudaEventRecord(b->ctx->ev[0], st));

    int wpc = 8;
    const int sw_scratch = (int) awb_sw_warp_scratch_bytes(b->maxS, b->maxT);
    while (wpc > 1 && (size_t) wpc * sw_scratch > 160 * 1024)
        wpc >>= 1;
    for (int g = 0; g < AWB_UPLOAD_GROUPS; g++) {
        int g0, g1;
        upload_group(b, g, g0, g1);
        if (g1 <= g0)
            continue;
        // the group's inputs have arrived
        CUDA_OK(cudaStreamWaitEvent(st, b->ctx->up_ev[g], 0));
        const AwbChain *chains = b->d_chains + g0;
        const int Cg = g1 - g0;
        {
            dim3 grid((b->maxB + 31) / 32, Cg);
            {
                KTimer kt(b, AWB_K_BLOCK);
                if (b->maxV <= AWB_K1_VCAP_SMALL)
                    awb_block_setup_kernel<AWB_K1_VCAP_SMALL><<<grid, 32, 0, st>>>(chains, b->d_err);
                else if (b->maxV <= AWB_K1_VCAP_MID)
                    awb_block_setup_kernel<AWB_K1_VCAP_MID><<<grid, 32, 0, st>>>(chains, b->d_err);
                else
                    awb_block_setup_kernel<AWB_MAXV><<<grid, 32, 0, st>>>(chains, b->d_err);
            }
            dim3 grid2(b->maxB, Cg);
            KTimer kt(b, AWB_K_TMATRIX);
            awb_tmatrix_kernel<<<grid2, 128, 0, st>>>(chains);
        }
        if (b->maxB > 1) {
            dim3 grid((b->maxB - 1 + wpc - 1) / wpc, Cg);
            KTimer kt(b, AWB_K_SWITCH);
            awb_switch_setup_kernel<<<grid, 32 * wpc, (size_t) wpc * sw_scratch, st>>>(
                chains, b->d_err, sw_scratch);
        }
    }
    {
        // site kinds: the first kernel that reads the sequences
        CUDA_OK(cudaStreamWaitEvent(st, b->ctx->seq_ev, 0));
        dim3 grid((b->maxn + 255) / 256, b->C);
        if (grid.x > 4096) grid.x = 4096;
        KTimer kt(b, AWB_K_KIND);
        awb_kind_kernel<<<grid, 256, 0, st>>>(b->d_chains);
        if (b->any_packed) {
            dim3 grid2((b->maxnvar + 255) / 256, b->C);
            if (grid2.x > 1024) grid2.x = 1024;
            if (grid2.x < 1) grid2.x = 1;
            awb_kind_packed_kernel<<<grid2, 256, 0, st>>>(b->d_chains);
            b->launches++;
        }
    }
    b->launches += b->maxB > 1 ? 4 : 3;
    // variant-site emissions go into the forward table; with a checkpointed
    // table they are recomputed per segment, right before its forward pass
    if (!b->ckpt) {
        if (launch_emit(b, 0, 0))
            return 1;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(b->ctx->ev[1], st));
    b->setup_done = true;
    return 0;
}

extern "C" int awb_batch_forward(awb_batch *b, const double *const *priors)
{
    if (!b->setup_done) return fail("awb_batch_forward: setup has not run");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    if (priors) {
        for (int c = 0; c < b->C; c++) {
            if (!priors[c]) continue;
            const AwbLayout &L = b->L[c];
            const int S1 = L.nstates[0] > 0 ? L.nstates[0] : 1;
            CUDA_OK(cudaMemcpyAsync(b->h_chains[c].fw, priors[c],
                                    sizeof(double) * S1, cudaMemcpyHostToDevice,
                                    st));
        }
    }
    const int NS = ((b->maxS + 31) / 32) * 32;
    // fast path (awb_forward_fast.cuh): node-major threads + scribe warps; else
    // the generic kernel
    const bool fast = batch_fast_path(b);
    CUDA_OK(cudaEventRecord(b->ctx->ev[2], st));
    if (b->ckpt) {
        // checkpointed table, first pass: segment by segment (emissions, then
        // the forward pass of the segment); every segment leaves its last
        // column -- the first of the next segment -- in ckptcol
        for (int s = 0; s < b->maxseg; s++) {
            if (launch_emit(b, s, 0) || launch_forward_fast(b, s, 0))
                return 1;
        }
    } else if (fast) {
        if (launch_forward_fast(b, 0, 0))
            return 1;
    } else {
        // generic kernel: up to 1024 threads, 1 or 2 states per thread; the
        // band goes to shared memory when it fits, else it is read in place
        const int GNS = NS < 1024 ? NS : 1024;
        const int npt = (NS + GNS - 1) / GNS;
        int bandcap = b->maxband;
        size_t smem = awb_fwd_smem_bytes(npt * GNS, b->maxT, bandcap);
        if (smem > 200 * 1024) {
            bandcap = 0;
            smem = awb_fwd_smem_bytes(npt * GNS, b->maxT, 0);
        }
        if (smem > 200 * 1024 || npt > 2)
            return fail("forward kernel needs " + std::to_string(smem) +
                        " bytes of shared memory (limit 200 KiB)");
        if (npt == 1)
            awb_forward_kernel<1><<<b->C, GNS, smem, st>>>(b->d_chains, bandcap);
        else
            awb_forward_kernel<2><<<b->C, GNS, smem, st>>>(b->d_chains, bandcap);
        b->launches++;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(b->ctx->ev[3], st));
    b->forward_done = true;
    b->tables_stale = false;
    return 0;
}

// The random draws go up on the copy stream, next to whatever the compute
// stream still has queued (normally the forward pass); the traceback waits for
// rand_ev.
static int upload_rand_ints(awb_batch *b, const int *const *rand_ints)
{
    cudaStream_t cs = b->ctx->copy_stream;
    if (b->rand_in_use) {
        // an earlier traceback of this batch may still be reading the old draws
        CUDA_OK(cudaEventRecord(b->ctx->order_ev, b->ctx->stream));
        CUDA_OK(cudaStreamWaitEvent(cs, b->ctx->order_ev, 0));
    }
    for (int c = 0; c < b->C; c++)
        CUDA_OK(cudaMemcpyAsync((void *) b->h_chains[c].rand_ints, rand_ints[c],
                                sizeof(int) * b->L[c].n, cudaMemcpyHostToDevice, cs));
    CUDA_OK(cudaEventRecord(b->ctx->rand_ev, cs));
    b->rand_uploaded = true;
    return 0;
}

extern "C" int awb_batch_traceback(awb_batch *b, const int *const *rand_ints,
                                   int rand_max, const int *last_states)
{
    if (!b->forward_done) return fail("awb_batch_traceback: forward has not run");
    if (!rand_ints && !b->rand_uploaded)
        return fail("awb_batch_traceback: rand_ints is required");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    bool chains_dirty = false;
    if (rand_ints && upload_rand_ints(b, rand_ints))
        return 1;
    CUDA_OK(cudaStreamWaitEvent(st, b->ctx->rand_ev, 0));
    b->rand_in_use = true;
    for (int c = 0; c < b->C; c++) {
        const int ls = last_states ? last_states[c] : -1;
        if (ls != b->h_chains[c].last_state) {
            b->h_chains[c].last_state = ls;
            chains_dirty = true;
        }
    }
    if (chains_dirty && b->stage_chains) {
        // keep the pinned copy (used by a later awb_batch_upload) in step
        CUDA_OK(cudaStreamSynchronize(b->ctx->copy_stream));
        memcpy(b->stage_chains, b->h_chains.data(), sizeof(AwbChain) * b->C);
    }
    if (chains_dirty)
        CUDA_OK(cudaMemcpyAsync(b->d_chains, b->h_chains.data(),
                                sizeof(AwbChain) * b->C, cudaMemcpyHostToDevice,
                                st));
    CUDA_OK(cudaEventRecord(b->ctx->ev[4], st));
    if (b->ckpt) {
        // checkpointed table, second pass: from the last segment to the first,
        // rebuild the segment's table from its stored first column, then walk
        // back through it
        // (pass 1: the last segments' tables are still resident from the
        // forward pass and are not rebuilt; pass 2, a further traceback of the
        // same forward pass: the rebuilt segments have taken turns in table 0,
        // which is also the first resident one, so every table is rebuilt)
        const int pass = b->tables_stale ? 2 : 1;
        // With three or more tables per window consecutive segments are rebuilt
        // side by side (independent chains: each starts from its own stored
        // column), which puts several CTAs on every SM -- as many as the dense
        // forward kernel fits there; the traceback then walks through them in
        // order.  A group never straddles the resident boundary of any window
        // (above it that window's tables are still in use; below it all of them
        // are free); segment s lives in table R-1 - s mod nrot (AwbSeg).
        // (Tried: the next group rebuilt on a second stream while the traceback
        // walks the current one, tables rotating with twice the period: 1 % --
        // the traceback's CTAs fill the SMs' shared memory, the rebuild waits.)
        const int G = b->nrot;
        std::vector<char> boundary(b->maxseg + 1, 0);
        for (int c = 0; c < b->C; c++) {
            const int ns = b->L[c].nseg;
            const int R = b->nslots < ns ? b->nslots : ns;
            boundary[ns - R] = 1;          // segments >= ns - R are resident
        }
        for (int s = b->maxseg - 1; s >= 0;) {
            const bool res_s = s >= b->maxseg - b->nslots;
            int nsub = 1;
            if (!res_s && G > 1 && !getenv("AWB_NO_PAIRS")) {
                // down to the next multiple of G, or to a window's boundary
                while (nsub < G && (s - nsub + 1) % G != 0 && s - nsub >= 0 &&
                       !boundary[s - nsub + 1])
                    nsub++;
            }
            for (int y = 0; y < nsub; y++)
                if (launch_emit(b, s - y, pass))
                    return 1;
            if (launch_forward_fast(b, s, pass, nsub))
                return 1;
            for (int y = 0; y < nsub; y++)
                if (launch_traceback(b, rand_max, s - y))
                    return 1;
            s -= nsub;
        }
        if (b->maxseg > b->nslots)
            b->tables_stale = true;
    } else {
        if (launch_traceback(b, rand_max, 0))
            return 1;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(b->ctx->ev[5], st));
    return 0;
}

extern "C" int awb_batch_upload_rand(awb_batch *b, const int *const *rand_ints)
{
    CUDA_OK(cudaSetDevice(b->ctx->device));
    if (!b->bound && batch_bind(b))
        return 1;
    return upload_rand_ints(b, rand_ints);
}

extern "C" int awb_batch_sync(awb_batch *b)
{
    CUDA_OK(cudaSetDevice(b->ctx->device));
    CUDA_OK(cudaStreamSynchronize(b->ctx->copy_stream));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    int err = 0;
    CUDA_OK(cudaMemcpy(&err, b->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (err)
        return fail("device-side setup error code " + std::to_string(err));
    return 0;
}

extern "C" int awb_batch_timings(awb_batch *b, float *setup_ms, float *forward_ms,
                                 float *traceback_ms)
{
    CUDA_OK(cudaSetDevice(b->ctx->device));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    float v = 0;
    if (setup_ms) {
        *setup_ms = 0;
        if (cudaEventElapsedTime(&v, b->ctx->ev[0], b->ctx->ev[1]) == cudaSuccess)
            *setup_ms = v;
    }
    if (forward_ms) {
        *forward_ms = 0;
        if (cudaEventElapsedTime(&v, b->ctx->ev[2], b->ctx->ev[3]) == cudaSuccess)
            *forward_ms = v;
    }
    if (traceback_ms) {
        *traceback_ms = 0;
        if (cudaEventElapsedTime(&v, b->ctx->ev[4], b->ctx->ev[5]) == cudaSuccess)
            *traceback_ms = v;
    }
    cudaGetLastError();
    return 0;
}

extern "C" int awb_batch_kernel_times(awb_batch *b, int enable)
{
    CUDA_OK(cudaSetDevice(b->ctx->device));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    b->ktimes = enable != 0;
    b->kev_used = 0;
    b->kclass.clear();
    b->k4_bytes = 0;
    b->k4_launches = 0;
    return 0;
}

extern "C" int awb_batch_get_kernel_times(awb_batch *b, float *ms, double *forward_bytes,
                                          int *forward_launches)
{
    CUDA_OK(cudaSetDevice(b->ctx->device));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    for (int k = 0; k < AWB_K_NCLASS; k++)
        ms[k] = 0;
    for (size_t i = 0; i < b->kclass.size(); i++) {
        float v = 0;
        if (cudaEventElapsedTime(&v, b->kev[2 * i], b->kev[2 * i + 1]) == cudaSuccess)
            ms[b->kclass[i]] += v;
    }
    cudaGetLastError();
    if (forward_bytes) *forward_bytes = b->k4_bytes;
    if (forward_launches) *forward_launches = b->k4_launches;
    // (the events are reused: the next read covers what is launched from here on)
    b->kev_used = 0;
    b->kclass.clear();
    b->k4_bytes = 0;
    b->k4_launches = 0;
    return 0;
}

extern "C" double awb_batch_states_sites(const awb_batch *b, int i)
{
    return b->L[i].states_sites;
}

extern "C" int64_t awb_batch_fw_doubles(const awb_batch *b, int i)
{
    return b->L[i].fw_off[b->L[i].B];
}

extern "C" int awb_batch_nsites(const awb_batch *b, int i) { return b->L[i].n; }

extern "C" int awb_batch_kernel_launches(const awb_batch *b) { return b->launches; }
extern "C" int awb_batch_segments(const awb_batch *b) { return b->maxseg; }
extern "C" int awb_batch_forward_kernel(const awb_batch *b)
{
    return batch_fast_path(b) ? 1 : 0;
}
extern "C" int awb_batch_resident_segments(const awb_batch *b)
{
    return b->bound ? b->nslots : 1;
}

extern "C" int awb_batch_get_path(awb_batch *b, int i, int *path)
{
    CUDA_OK(cudaSetDevice(b->ctx->device));
    if (!b->bound) return fail("awb_batch_get_path: nothing has been uploaded or computed");
    CUDA_OK(cudaMemcpyAsync(path, b->h_chains[i].path, sizeof(int) * b->L[i].n,
                            cudaMemcpyDeviceToHost, b->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    return 0;
}

extern "C" int awb_batch_get_logz(awb_batch *b, int i, double *logz)
{
    CUDA_OK(cudaSetDevice(b->ctx->device));
    if (!b->bound) return fail("awb_batch_get_logz: nothing has been uploaded or computed");
    CUDA_OK(cudaMemcpyAsync(logz, b->h_chains[i].logz, sizeof(double),
                            cudaMemcpyDeviceToHost, b->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    return 0;
}

extern "C" int awb_batch_get_status(awb_batch *b, int i, int *first_bad_site)
{
    CUDA_OK(cudaSetDevice(b->ctx->device));
    if (!b->bound) return fail("awb_batch_get_status: nothing has been uploaded or computed");
    CUDA_OK(cudaMemcpyAsync(first_bad_site, b->h_chains[i].status, sizeof(int),
                            cudaMemcpyDeviceToHost, b->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    return 0;
}

extern "C" int awb_batch_get_fw(awb_batch *b, int i, double *fw)
{
    if (b->ckpt)
        return fail("awb_batch_get_fw: the forward table is not kept with AWB_CHECKPOINT");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    if (!b->bound) return fail("awb_batch_get_fw: nothing has been uploaded or computed");
    const AwbLayout &L = b->L[i];
    CUDA_OK(cudaMemcpyAsync(fw, b->h_chains[i].fw, sizeof(double) * L.fw_off[L.B],
                            cudaMemcpyDeviceToHost, b->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    if (batch_fast_path(b)) {
        // the fast kernel stores the columns as the recursion carries them (a
        // positive scale per row; awb_forward_fast.cuh): hand them out the way
        // the reference stores them, every column after the first summing to 1
        // (sample_thread.cpp:292-294, :386-388)
        for (int blk = 0; blk < L.B; blk++) {
            const int S1 = L.nstates[blk] > 0 ? L.nstates[blk] : 1;
            double *row = fw + L.fw_off[blk];
            for (int r = 0; r < L.block_start[blk + 1] - L.block_start[blk]; r++, row += S1) {
                if (blk == 0 && r == 0)
                    continue;               // the prior, kept as given
                double sum = 0.0;
                for (int j = 0; j < S1; j++)
                    sum += row[j];
                const double inv = 1.0 / sum;
                for (int j = 0; j < S1; j++)
                    row[j] *= inv;
            }
        }
    }
    return 0;
}

extern "C" int awb_batch_get_nstates(awb_batch *b, int i, int *nstates)
{
    memcpy(nstates, b->L[i].nstates.data(), sizeof(int) * b->L[i].B);
    return 0;
}

extern "C" int awb_batch_get_layout(awb_batch *b, int i, int64_t *row_off,
                                    int64_t *fw_off, int64_t *sw1_off)
{
    const AwbLayout &L = b->L[i];
    for (int k = 0; k <= L.B; k++) {
        if (row_off) row_off[k] = L.row_off[k];
        if (fw_off) fw_off[k] = L.fw_off[k];
        if (sw1_off) sw1_off[k] = L.sw1_off[k];
    }
    return 0;
}

extern "C" int awb_batch_get_debug(awb_batch *b, int i, const char *name,
                                   void *dst, int64_t dst_bytes)
{
    size_t off, bytes;
    if (!awb_layout_find(b->L[i], name, off, bytes))
        return fail(std::string("unknown or unavailable array: ") + name);
    if ((int64_t) bytes > dst_bytes)
        return fail("destination too small");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    CUDA_OK(cudaMemcpy(dst, b->arena + b->arena_off[i] + off, bytes,
                       cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int64_t awb_batch_debug_bytes(awb_batch *b, int i, const char *name)
{
    size_t off, bytes;
    if (!awb_layout_find(b->L[i], name, off, bytes))
        return -1;
    return (int64_t) bytes;
}

// ---------------------------------------------------------------- phase probabilities

extern "C" int awb_batch_phase_probs(awb_batch *b)
{
    if (!b->forward_done || !b->rand_in_use)
        return fail("awb_batch_phase_probs: the traceback has not run");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    if (!b->d_phase) {
        b->phase_off.assign(b->C + 1, 0);
        for (int c = 0; c < b->C; c++)
            b->phase_off[c + 1] = b->phase_off[c] + b->L[c].n;
        CUDA_OK(cudaMalloc((void **) &b->d_phase, sizeof(double) * (size_t) b->phase_off[b->C]));
        CUDA_OK(cudaMalloc((void **) &b->d_phase_off, sizeof(long long) * (b->C + 1)));
        CUDA_OK(cudaMemcpyAsync(b->d_phase_off, b->phase_off.data(),
                                sizeof(long long) * (b->C + 1), cudaMemcpyHostToDevice, st));
    }
    const int scratch = (int) (((awb_emit_scratch_bytes(b->maxV) + 15) & ~(size_t) 15) +
                               (((size_t) b->maxV * 16 + 4 + 15) & ~(size_t) 15));
    int wpc = 8;
    while (wpc > 1 && (size_t) wpc * scratch > 160 * 1024)
        wpc >>= 1;
    if ((size_t) wpc * scratch > 200 * 1024)
        return fail("tree too large for the emission kernel's shared memory");
    int gx = (b->maxn + 32 * wpc - 1) / (32 * wpc);
    const int cap = b->ctx->sm_count * 16;
    if (gx * b->C > cap) gx = (cap + b->C - 1) / b->C;
    if (gx < 1) gx = 1;
    dim3 grid(gx, b->C);
    CUDA_OK(cudaFuncSetAttribute(awb_phase_kernel,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    awb_phase_kernel<<<grid, 32 * wpc, (size_t) wpc * scratch, st>>>(
        b->d_chains, scratch, b->d_phase, b->d_phase_off);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(st));
    b->launches++;
    return 0;
}

extern "C" int awb_batch_get_phase_probs(awb_batch *b, int i, double *p)
{
    if (!b->d_phase) return fail("awb_batch_get_phase_probs: nothing computed");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    CUDA_OK(cudaMemcpyAsync(p, b->d_phase + b->phase_off[i], sizeof(double) * b->L[i].n,
                            cudaMemcpyDeviceToHost, b->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    return 0;
}

// ---------------------------------------------------------------- recombination points

// glibc's rand() is random_r on a hidden TYPE_3 state (31 words, r[i] += r[i-3]);
// initstate() hands out the old state array with the rear index stored in front
// of it (glibc stdlib/random_r.c: __initstate_r / __setstate_r).
extern "C" int awb_libc_rand_snapshot(int *state)
{
    static int tmp[32];                 // (a state array has to be word aligned)
    char *old = initstate(1u, (char *) tmp, sizeof(tmp));
    if (!old)
        return fail("initstate failed");
    const int *w = (const int *) old;
    const int type = w[0] % 5, rear = w[0] / 5;
    int rc = 0;
    if (type != 3 || rear < 0 || rear >= 31) {
        rc = fail("libc rand() is not in its default (TYPE_3) state");
    } else {
        for (int i = 0; i < 31; i++)
            state[i] = w[1 + i];
        state[31] = (rear + 3) % 31;     // front index
        state[32] = rear;
        state[33] = type;
    }
    setstate(old);
    return rc;
}

extern "C" void awb_libc_rand_advance(long long ndraws)
{
    for (long long i = 0; i < ndraws; i++)
        rand();
}

extern "C" int awb_rng_draw(int *state, int n, int *out)
{
    AwbRng g;
    for (int i = 0; i < 31; i++)
        g.r[i] = state[i];
    g.f = state[31];
    g.b = state[32];
    for (int i = 0; i < n; i++) {
        const int v = (int) awb_rng_next(g);
        if (out) out[i] = v;
    }
    for (int i = 0; i < 31; i++)
        state[i] = g.r[i];
    state[31] = g.f;
    state[32] = g.b;
    return 0;
}

extern "C" int awb_batch_sample_recombs(awb_batch *b, const int *rng_states,
                                        int rand_max)
{
    if (!b->forward_done || !b->rand_in_use)
        return fail("awb_batch_sample_recombs: the traceback has not run");
    if (!rng_states)
        return fail("awb_batch_sample_recombs: rng_states is required");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    cudaStream_t st = b->ctx->stream;
    if (!b->d_rec) {
        b->rec_off.assign(b->C + 1, 0);
        for (int c = 0; c < b->C; c++)
            b->rec_off[c + 1] = b->rec_off[c] + 3ll * b->L[c].n;
        CUDA_OK(cudaMalloc((void **) &b->d_rec, sizeof(int) * (size_t) b->rec_off[b->C]));
        CUDA_OK(cudaMalloc((void **) &b->d_rec_off, sizeof(long long) * (b->C + 1)));
        CUDA_OK(cudaMalloc((void **) &b->d_rec_info, sizeof(int) * 2 * b->C));
        CUDA_OK(cudaMalloc((void **) &b->d_rng, sizeof(int) * AWB_RNG_WORDS * b->C));
        CUDA_OK(cudaMemcpyAsync(b->d_rec_off, b->rec_off.data(),
                                sizeof(long long) * (b->C + 1), cudaMemcpyHostToDevice, st));
    }
    CUDA_OK(cudaMemcpyAsync(b->d_rng, rng_states, sizeof(int) * AWB_RNG_WORDS * b->C,
                            cudaMemcpyHostToDevice, st));
    {
        KTimer kt(b, AWB_K_RECOMB);
        awb_recomb_kernel<<<b->C, 32, 0, st>>>(b->d_chains, b->d_rng, rand_max, b->d_rec,
                                               b->d_rec_off, b->d_rec_info);
    }
    CUDA_OK(cudaGetLastError());
    // (rng_states and rec_off may be pageable: the copies above must not outlive them)
    CUDA_OK(cudaStreamSynchronize(st));
    b->launches++;
    b->recombs_done = true;
    return 0;
}

extern "C" int awb_batch_get_recomb_count(awb_batch *b, int i, int *nrecombs, int *draws)
{
    if (!b->recombs_done) return fail("awb_batch_get_recomb_count: nothing sampled");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    int info[2];
    CUDA_OK(cudaMemcpyAsync(info, b->d_rec_info + 2 * i, sizeof(info),
                            cudaMemcpyDeviceToHost, b->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    if (nrecombs) *nrecombs = info[0];
    if (draws) *draws = info[1];
    return 0;
}

extern "C" int awb_batch_get_recombs(awb_batch *b, int i, int count, int *pos,
                                     int *node, int *time)
{
    if (!b->recombs_done) return fail("awb_batch_get_recombs: nothing sampled");
    CUDA_OK(cudaSetDevice(b->ctx->device));
    const int cap = b->L[i].n;
    if (count > cap) count = cap;
    if (count <= 0) return 0;
    int *dst[3] = { pos, node, time };
    for (int k = 0; k < 3; k++)
        if (dst[k])
            CUDA_OK(cudaMemcpyAsync(dst[k], b->d_rec + b->rec_off[i] + (size_t) k * cap,
                                    sizeof(int) * count, cudaMemcpyDeviceToHost,
                                    b->ctx->stream));
    CUDA_OK(cudaStreamSynchronize(b->ctx->stream));
    return 0;
}

// ---------------------------------------------------------------- one-shot

static awb_ctx *g_default_ctx = NULL;

static int default_ctx(awb_ctx **ctx)
{
    if (!g_default_ctx) {
        int dev = 0;
        const char *env = getenv("AWB_DEVICE");
        if (env) dev = atoi(env);
        if (awb_ctx_create(dev, &g_default_ctx))
            return 1;
    }
    *ctx = g_default_ctx;
    return 0;
}

extern "C" int awb_thread_sample_cond(const awb_problem *p, const double *prior,
                                      int last_state, const int *rand_ints,
                                      int rand_max, int *path, double *logz)
{
    awb_ctx *ctx;
    if (default_ctx(&ctx)) return 1;
    awb_batch *b = NULL;
    if (awb_batch_create(ctx, 1, p, 0, &b)) return 1;
    const double *priors[1] = { prior };
    int rc = awb_batch_upload(b) || awb_batch_setup(b) ||
        awb_batch_forward(b, prior ? priors : NULL) ||
        awb_batch_traceback(b, &rand_ints, rand_max,
                            last_state >= 0 ? &last_state : NULL) ||
        awb_batch_sync(b);
    int bad = -1;
    if (!rc) rc = awb_batch_get_status(b, 0, &bad);
    if (!rc && bad >= 0)
        rc = fail("forward column " + std::to_string(bad) +
                  " has no positive entry (sample_thread.cpp:443-444)");
    if (!rc && path) rc = awb_batch_get_path(b, 0, path);
    if (!rc && logz) rc = awb_batch_get_logz(b, 0, logz);
    awb_batch_destroy(b);
    return rc;
}

extern "C" int awb_thread_sample_recombs(const awb_problem *p, const double *prior,
                                         int last_state, const int *rand_ints,
                                         int rand_max, const int *rng_state,
                                         int *path, double *logz, int cap,
                                         int *nrecombs, int *pos, int *node,
                                         int *time, int *draws)
{
    awb_ctx *ctx;
    if (default_ctx(&ctx)) return 1;
    awb_batch *b = NULL;
    if (awb_batch_create(ctx, 1, p, 0, &b)) return 1;
    const double *priors[1] = { prior };
    int snap[AWB_RNG_WORDS];
    int rc = 0;
    if (!rng_state) {
        // the caller's libc stream, as it stands after the traceback's draws
        rc = awb_libc_rand_snapshot(snap);
        rng_state = snap;
    }
    rc = rc || awb_batch_upload(b) || awb_batch_setup(b) ||
        awb_batch_forward(b, prior ? priors : NULL) ||
        awb_batch_traceback(b, &rand_ints, rand_max,
                            last_state >= 0 ? &last_state : NULL) ||
        awb_batch_sample_recombs(b, rng_state, rand_max) ||
        awb_batch_sync(b);
    int bad = -1;
    if (!rc) rc = awb_batch_get_status(b, 0, &bad);
    if (!rc && bad >= 0)
        rc = fail("forward column " + std::to_string(bad) +
                  " has no positive entry (sample_thread.cpp:443-444)");
    if (!rc && path) rc = awb_batch_get_path(b, 0, path);
    if (!rc && logz) rc = awb_batch_get_logz(b, 0, logz);
    int n = 0, d = 0;
    if (!rc) rc = awb_batch_get_recomb_count(b, 0, &n, &d);
    if (!rc) rc = awb_batch_get_recombs(b, 0, n < cap ? n : cap, pos, node, time);
    if (!rc && rng_state == snap)
        awb_libc_rand_advance(d);        // leave libc where the reference would
    if (nrecombs) *nrecombs = n;
    if (draws) *draws = d;
    awb_batch_destroy(b);
    return rc;
}

extern "C" int awb_thread_sample(const awb_problem *p, const int *rand_ints,
                                 int rand_max, int *path, double *logz)
{
    return awb_thread_sample_cond(p, NULL, -1, rand_ints, rand_max, path, logz);
}

extern "C" int awb_forward_table(const awb_problem *p, const double *prior,
                                 double *fw, double *logz)
{
    awb_ctx *ctx;
    if (default_ctx(&ctx)) return 1;
    awb_batch *b = NULL;
    if (awb_batch_create(ctx, 1, p, 0, &b)) return 1;
    const double *priors[1] = { prior };
    int rc = awb_batch_upload(b) || awb_batch_setup(b) ||
        awb_batch_forward(b, prior ? priors : NULL) || awb_batch_sync(b);
    if (!rc && fw) rc = awb_batch_get_fw(b, 0, fw);
    if (!rc && logz) rc = awb_batch_get_logz(b, 0, logz);
    awb_batch_destroy(b);
    return rc;
}
