// host_emul.cpp -- TEST HARNESS (CPU, no CUDA).  Compiles the product's
// __host__ __device__ setup/emission workers (argweaver_b200/csrc/awb_setup.cuh,
// awb_emit.cuh) and the host layout code (awb_layout.h) with g++ and runs them
// sequentially over one problem, so their outputs can be compared with the
// oracle on the CPU box before any GPU time is spent.
//
// It also contains a plain sequential walk over the device data structures
// (time matrix + same-branch band + switch CSR + site kinds) that reproduces
// the forward table -- this validates the *tables* the CUDA forward kernel
// consumes.  It is NOT a product path and is never shipped or benchmarked.

#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "argweaver_b200.h"
#include "awb_common.cuh"
#include "awb_emit.cuh"
#include "awb_layout.h"
#include "awb_recomb.cuh"
#include "awb_setup.cuh"

struct Emul {
    AwbLayout L;
    AwbChain ch;
    std::vector<char> arena;
    std::vector<unsigned char> scratch;
    std::string err;
    int code;
};

extern "C" {

void *emul_create(const awb_problem *p)
{
    Emul *e = new Emul;
    e->code = 0;
    if (!awb_layout_build(*p, 1, e->L, e->err)) {
        e->code = -1;
        return e;
    }
    e->arena.assign(e->L.total_bytes, 0);
    for (size_t i = 0; i < e->L.copies.size(); i++)
        memcpy(&e->arena[e->L.copies[i].dst_off], e->L.copies[i].src,
               e->L.copies[i].bytes);
    awb_layout_bind(e->L, *p, e->arena.data(), e->ch);
    // (the fast forward kernel's batches run K1 without the generic kernel's
    // tables: AWB_EMUL_NO_BAND=1 takes that branch; emul_forward needs them)
    if (getenv("AWB_EMUL_NO_BAND") && atoi(getenv("AWB_EMUL_NO_BAND")))
        e->ch.need_band = 0;
    e->scratch.assign(awb_emit_scratch_bytes(p->nnodes) + 64, 0);
    return e;
}

// the forward kernel's thread map of one block (first-fit decreasing, or node
// order): returns the slots used, fills tmap[0..cap) when given
int emul_pack_branches(const short *cnt, const short *nfirst, int V, int in_order,
                       unsigned short *tmap, int cap)
{
    static AwbRowLists<AWB_MAXV> rl;
    const int n = in_order ? awb_pack_place_in_order<AWB_MAXV>(cnt, V, &rl)
                           : awb_pack_place<AWB_MAXV>(cnt, V, &rl);
    if (!in_order && n != awb_pack_branches(cnt, V))
        return -1;                      // count-only mode disagrees
    if (tmap && n <= cap)
        awb_pack_emit<AWB_MAXV>(rl, cnt, nfirst, tmap, 0, cap);
    return n;
}

const char *emul_error(void *h) { return ((Emul *) h)->err.c_str(); }
int emul_code(void *h) { return ((Emul *) h)->code; }

int emul_setup(void *h)
{
    Emul *e = (Emul *) h;
    const AwbChain &ch = e->ch;
    for (int i = 0; i < ch.nsites; i++)
        awb_site_kind(ch, i);
    for (int b = 0; b < ch.ntrees; b++) {
        int rc = awb_block_setup(ch, b);
        if (rc) { e->code = 100 + rc; return e->code; }
    }
    for (int b = 1; b < ch.ntrees; b++) {
        int rc = awb_switch_setup(ch, b);
        if (rc) { e->code = 200 + rc; return e->code; }
    }
    for (int i = 0; i < ch.nsites; i++)
        if (i > 0 && ch.kind[i] == AWB_SITE_VARIANT) {
            const int b = awb_find_block(ch, i);
            const size_t o = (size_t) b * ch.nnodes;
            awb_emit_site(ch, i, b, 0, 1, e->scratch.data(), ch.ptrees + o,
                          ch.ages + o, ch.child0 + o, ch.child1 + o, ch.order + o,
                          ch.lstart + (size_t) b * (ch.nnodes + 2));
        }
    return 0;
}

// sequential walk over the device tables; overwrites ch.fw with the forward
// table (same semantics as the CUDA forward kernel)
int emul_forward(void *h, double *logz_out)
{
    Emul *e = (Emul *) h;
    const AwbChain &ch = e->ch;
    const int T = ch.model.ntimes;
    std::vector<double> col(ch.maxS), col2(ch.maxS), F(T), R(T);
    double logz = 0.0;
    int Sprev1 = 0;
    for (int b = 0; b < ch.ntrees; b++) {
        const int S = ch.nstates[b];
        const int S1 = S > 0 ? S : 1;
        const long long r0 = ch.row_off[b];
        double *fw = ch.fw + ch.fw_off[b];
        const int blen = ch.blocklens[b];
        const int s0 = ch.block_start[b];
        const double *tm = ch.tmatrix + (size_t) b * T * T;
        const double *band = ch.band + ch.band_off[b];
        for (int i = 0; i < blen; i++) {
            const int site = s0 + i;
            if (site == 0) {
                for (int k = 0; k < S1; k++) col[k] = fw[k];
                continue;     // prior column stays unnormalised
            }
            // emission of this site
            const int kd = ch.kind[site];
            if (i == 0) {
                // switch (sample_thread.cpp:345-389)
                for (int k = 0; k < S1; k++) {
                    double sum = 0.0;
                    const int st = ch.sw_start[r0 + k], cn = ch.sw_cnt[r0 + k];
                    for (int q = 0; q < cn; q++)
                        sum += col[ch.sw_src[ch.ent_off[b] + st + q]] *
                            ch.sw_prob[ch.ent_off[b] + st + q];
                    col2[k] = sum;
                }
            } else if (S == 0) {
                col2[0] = col[0];
            } else {
                for (int a = 0; a < T; a++) F[a] = 0.0;
                for (int k = 0; k < S; k++) F[ch.st_time[r0 + k]] += col[k];
                for (int bb = 0; bb < T - 1; bb++) {
                    double sum = 0.0;
                    for (int a = 0; a < T - 1; a++) sum += tm[a * T + bb] * F[a];
                    R[bb] = sum;
                }
                for (int k = 0; k < S; k++) {
                    double sum = R[ch.st_time[r0 + k]];
                    const int j1 = ch.band_j1[r0 + k], len = ch.band_len[r0 + k];
                    const double *cf = band + ch.band_boff[r0 + k];
                    for (int q = 0; q < len; q++) sum += cf[q] * col[j1 + q];
                    col2[k] = sum;
                }
            }
            double norm = 0.0;
            for (int k = 0; k < S1; k++) {
                double em = 1.0;
                if (S > 0) {
                    if (kd == AWB_SITE_VARIANT) em = fw[(size_t) i * S1 + k];
                    else if (kd == AWB_SITE_INVARIANT) em = ch.inv_emit[r0 + k];
                }
                col2[k] *= em;
                norm += col2[k];
            }
            for (int k = 0; k < S1; k++) {
                col[k] = col2[k] / norm;
                fw[(size_t) i * S1 + k] = col[k];
            }
            logz += log(norm);
        }
        Sprev1 = S1;
    }
    (void) Sprev1;
    if (logz_out) *logz_out = logz;
    return 0;
}

long long emul_array_bytes(void *h, const char *name)
{
    Emul *e = (Emul *) h;
    size_t off, bytes;
    if (!awb_layout_find(e->L, name, off, bytes)) return -1;
    return (long long) bytes;
}

int emul_get(void *h, const char *name, void *dst, long long dst_bytes)
{
    Emul *e = (Emul *) h;
    size_t off, bytes;
    if (!awb_layout_find(e->L, name, off, bytes)) return -1;
    if ((long long) bytes > dst_bytes) return -2;
    memcpy(dst, &e->arena[off], bytes);
    return 0;
}

void emul_layout(void *h, int *nstates, long long *row_off, long long *fw_off,
                 long long *sw1_off, int *maxS, int *maxband)
{
    Emul *e = (Emul *) h;
    const int B = e->L.B;
    memcpy(nstates, e->L.nstates.data(), B * sizeof(int));
    memcpy(row_off, e->L.row_off.data(), (B + 1) * sizeof(long long));
    memcpy(fw_off, e->L.fw_off.data(), (B + 1) * sizeof(long long));
    memcpy(sw1_off, e->L.sw1_off.data(), (B + 1) * sizeof(long long));
    *maxS = e->L.maxS;
    *maxband = e->L.maxband;
}

// 1: the layout found that exp(lnB) may overflow for this problem
int emul_lin_unsafe(void *h) { return ((Emul *) h)->L.lin_unsafe; }

// largest |value| of K1's linear-domain vectors (inf / nan when they overflow)
double emul_lin_max(void *h)
{
    Emul *e = (Emul *) h;
    const size_t n = (size_t) e->L.B * 7 * e->L.T;
    const double *lin = (const double *) &e->arena[e->L.o_lin];
    double m = 0.0;
    for (size_t i = 0; i < n; i++) {
        if (!(lin[i] == lin[i])) return NAN;
        if (fabs(lin[i]) > m) m = fabs(lin[i]);
    }
    return m;
}

// the recombination sampler (awb_recomb.cuh) over a given path, after emul_setup
int emul_sample_recombs(void *h, const int *path, const int *rng_state, int rand_max,
                        int cap, int *pos, int *node, int *time, int *info)
{
    Emul *e = (Emul *) h;
    const AwbChain &ch = e->ch;
    memcpy(ch.path, path, sizeof(int) * ch.nsites);
    AwbRng rng;
    for (int i = 0; i < 31; i++) rng.r[i] = rng_state[i];
    rng.f = rng_state[31];
    rng.b = rng_state[32];
    awb_sample_recombs(ch, rng, rand_max, pos, node, time, cap, info, -1);
    return 0;
}

// unphased data: P(phasing as given | path state) at the heterozygous variant
// sites (awb_emit.cuh awb_phase_prob), -1 elsewhere; after emul_setup
int emul_phase_probs(void *h, const int *path, double *out)
{
    Emul *e = (Emul *) h;
    const AwbChain &ch = e->ch;
    for (int i = 0; i < ch.nsites; i++) {
        out[i] = -1.0;
        if (ch.kind[i] != AWB_SITE_VARIANT || !awb_site_het(ch, i))
            continue;
        const int b = awb_find_block(ch, i);
        if (ch.nstates[b] == 0)
            continue;
        const size_t o = (size_t) b * ch.nnodes;
        out[i] = awb_phase_prob(ch, i, b, path[i], 0, 1, e->scratch.data(), ch.ptrees + o,
                                ch.ages + o, ch.child0 + o, ch.child1 + o, ch.order + o,
                                ch.lstart + (size_t) b * (ch.nnodes + 2));
    }
    return 0;
}

void emul_destroy(void *h) { delete (Emul *) h; }

} // extern "C"

// host layout alone (timed by scripts/layout_time.py); returns 0 on success
extern "C" int emul_layout_only(const awb_problem *p, int ckpt)
{
    AwbLayout L;
    std::string err;
    return awb_layout_build(*p, 0, L, err, ckpt ? (1ll << 24) : 0) ? 0 : -1;
}

// thread slots of the fast forward kernel (max over blocks) for a problem
extern "C" int emul_layout_maxns(const awb_problem *p)
{
    AwbLayout L;
    std::string err;
    return awb_layout_build(*p, 0, L, err, 0) ? L.maxNS : -1;
}
