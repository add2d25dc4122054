// dsmc_sim.cpp -- synthetic input generator (host only, no CUDA).
//
// Produces what `arg-sim` + a partially built ARG give the reference's thread
// sampler: an ARG over `nleaves` sequences as (tree, SPR, blocklen) triples with
// the reference's node-numbering invariant, plus nleaves+1 aligned sequences
// (the last one is the sequence to be threaded).  Simulation happens directly at
// compressed-site resolution (rho, mu are per compressed site), in discretised
// time, following the structure of the reference's DSMC simulator
// (argweaver/sim.py:132-263 sample_dsmc_sprs, :297-323 sample_arg_mutations);
// it is a from-scratch restatement, not a translation, and is only used to
// make inputs for tests and benchmarks.
//
// Node numbering (reference local_tree.h:767-776, local_tree.cpp:223-314):
// leaves are 0..nleaves-1; across an SPR every node keeps its index except the
// broken node (parent of the recombining branch), whose index is reused by the
// new re-coalescence node.

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {               // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    int below(int n) { return (int) (uniform() * n) % n; }
    double expo(double rate) { return -log(1.0 - uniform()) / rate; }
    int poisson(double mean) {
        if (mean <= 0) return 0;
        if (mean > 30) {            // normal approximation is fine for inputs
            double u1 = uniform(), u2 = uniform();
            double z = sqrt(-2 * log(1 - u1)) * cos(6.283185307179586 * u2);
            int v = (int) floor(mean + sqrt(mean) * z + 0.5);
            return v < 0 ? 0 : v;
        }
        double L = exp(-mean), p = 1.0;
        int k = 0;
        do { k++; p *= uniform(); } while (p > L);
        return k - 1;
    }
};

struct Tree {
    int V;
    std::vector<int> parent, age, c0, c1;
    int root;
    void rebuild_children() {
        c0.assign(V, -1);
        c1.assign(V, -1);
        root = -1;
        for (int i = 0; i < V; i++) {
            int p = parent[i];
            if (p == -1) root = i;
            else if (c0[p] == -1) c0[p] = i;
            else c1[p] = i;
        }
    }
};

struct Sim {
    int nleaves, nsites, ntimes, V;
    std::vector<double> times, popsizes;
    std::vector<int> ptrees, ages, sprs, blocklens;
    std::vector<unsigned char> seqs;     // [nleaves+1][nsites]
    int ntrees;
};

// nearest time-grid index, never the top time point
int discretize(const std::vector<double> &times, double t)
{
    int n = (int) times.size();
    int best = 0;
    double bd = fabs(times[0] - t);
    for (int i = 1; i < n - 1; i++) {
        double d = fabs(times[i] - t);
        if (d < bd) { bd = d; best = i; }
    }
    return best;
}

// coalescent tree with piecewise-constant population size, ages on the grid
void sample_tree(Tree &t, int nleaves, const std::vector<double> &times,
                 const std::vector<double> &popsizes, Rng &rng)
{
    const int ntimes = (int) times.size();
    t.V = 2 * nleaves - 1;
    t.parent.assign(t.V, -1);
    t.age.assign(t.V, 0);
    std::vector<int> live;
    for (int i = 0; i < nleaves; i++) live.push_back(i);
    double now = 0.0;
    int interval = 0;
    int next = nleaves;
    while (live.size() > 1) {
        int k = (int) live.size();
        double rate = k * (k - 1) / 2.0 / (2.0 * popsizes[std::min(interval, ntimes - 1)]);
        double wait = rng.expo(rate);
        double bound = (interval + 1 < ntimes) ? times[interval + 1] : INFINITY;
        if (now + wait > bound) {
            now = bound;
            interval++;
            continue;
        }
        now += wait;
        int a = rng.below(k);
        int b = rng.below(k - 1);
        if (b >= a) b++;
        int na = live[a], nb = live[b];
        int node = next++;
        t.parent[na] = node;
        t.parent[nb] = node;
        int ag = discretize(times, now);
        ag = std::max(ag, std::max(t.age[na], t.age[nb]));
        t.age[node] = ag;
        if (a < b) std::swap(a, b);
        live.erase(live.begin() + a);
        live.erase(live.begin() + b);
        live.push_back(node);
    }
    t.rebuild_children();
}

// reference local_tree.cpp:223-314 semantics (index-stable SPR)
void apply_spr(Tree &t, int recomb_node, int coal_node, int coal_time)
{
    if (recomb_node == t.root) return;
    int recoal = t.parent[recomb_node];
    int sib = (t.c0[recoal] == recomb_node) ? t.c1[recoal] : t.c0[recoal];
    int broke_parent = t.parent[recoal];

    // detach the broken node: sibling takes its place
    t.parent[sib] = broke_parent;
    if (coal_node == recoal) {
        // re-coalescing on the branch above the broken node == above sibling
        coal_node = sib;
    }
    // insert recoal between coal_node and its parent
    int cp = t.parent[coal_node];
    t.parent[recoal] = cp;
    t.parent[coal_node] = recoal;
    t.parent[recomb_node] = recoal;
    t.age[recoal] = coal_time;
    t.rebuild_children();
}

void lineage_counts(const Tree &t, int ntimes, std::vector<int> &nbranches)
{
    nbranches.assign(ntimes, 0);
    for (int i = 0; i < t.V; i++) {
        int p = t.parent[i];
        int pa = (p == -1) ? ntimes - 2 : t.age[p];
        for (int j = t.age[i]; j < pa; j++) nbranches[j]++;
        if (p == -1) nbranches[pa]++;
    }
    nbranches[ntimes - 1] = 1;
}

bool branch_at(const Tree &t, int node, int time, int ntimes)
{
    int p = t.parent[node];
    int top = (p == -1) ? ntimes - 2 : t.age[p];
    return t.age[node] <= time && time <= top;
}

void collect_leaves(const Tree &t, int node, std::vector<int> &out)
{
    std::vector<int> st(1, node);
    while (!st.empty()) {
        int x = st.back();
        st.pop_back();
        if (t.c0[x] == -1) out.push_back(x);
        else { st.push_back(t.c0[x]); st.push_back(t.c1[x]); }
    }
}

Sim *simulate(int nleaves, int nsites, int ntimes, const double *times_,
              const double *popsizes_, double rho, double mu, uint64_t seed)
{
    Sim *S = new Sim;
    S->nleaves = nleaves;
    S->nsites = nsites;
    S->ntimes = ntimes;
    S->V = 2 * nleaves - 1;
    S->times.assign(times_, times_ + ntimes);
    S->popsizes.assign(popsizes_, popsizes_ + ntimes);
    const std::vector<double> &times = S->times;
    const int V = S->V;
    Rng rng(seed);

    std::vector<double> coal_steps(2 * ntimes, 0.0);   // model.cpp:9-23
    {
        std::vector<double> t2(2 * ntimes + 1, times[ntimes - 1]);
        for (int i = 0; i < ntimes - 1; i++) {
            t2[2 * i] = times[i];
            t2[2 * i + 1] = sqrt((times[i + 1] + 1.0) * (times[i] + 1.0));
        }
        for (int i = 0; i < 2 * ntimes - 2; i++) coal_steps[i] = t2[i + 1] - t2[i];
        coal_steps[2 * ntimes - 2] = INFINITY;
    }
    const double mintime = times[1] * 0.1;

    Tree t;
    if (nleaves == 1) {
        t.V = 1; t.parent.assign(1, -1); t.age.assign(1, 0); t.rebuild_children();
    } else {
        sample_tree(t, nleaves, times, S->popsizes, rng);
    }

    S->seqs.assign((size_t) (nleaves + 1) * nsites, 'A');
    static const char bases[4] = { 'A', 'C', 'G', 'T' };
    for (int i = 0; i < nsites; i++) {
        unsigned char b = bases[rng.below(4)];
        for (int j = 0; j <= nleaves; j++) S->seqs[(size_t) j * nsites + i] = b;
    }

    // state of the extra (to-be-threaded) lineage: (node, time index)
    int th_node = 0, th_time = 0;
    auto pick_thread_state = [&](const Tree &tr) {
        std::vector<int> nb;
        lineage_counts(tr, ntimes, nb);
        int j = 0;
        while (j < ntimes - 2) {
            double A = coal_steps[2 * j] * nb[j];
            if (j > 0) A += coal_steps[2 * j - 1] * nb[j - 1];
            double pc = 1.0 - exp(-A / (2.0 * S->popsizes[j]));
            if (rng.uniform() < pc) break;
            j++;
        }
        std::vector<int> cand;
        for (int x = 0; x < tr.V; x++)
            if (branch_at(tr, x, j, ntimes)) cand.push_back(x);
        th_time = j;
        th_node = cand[rng.below((int) cand.size())];
    };

    int pos = 0;
    int spr[4] = { -1, -1, -1, -1 };
    std::vector<int> nbranches, leaves;
    bool first = true;
    while (pos < nsites) {
        // block length: sim.py sample_next_recomb (minlen 1)
        double treelen = 0.0;
        for (int i = 0; i < t.V; i++)
            if (t.parent[i] != -1) treelen += times[t.age[t.parent[i]]] - times[t.age[i]];
        int blocklen;
        if (t.V == 1) {
            blocklen = nsites - pos;
        } else {
            double rate = std::max(rho * treelen, rho);
            blocklen = 1 + (int) floor(rng.expo(rate));
            if (blocklen > nsites - pos) blocklen = nsites - pos;
        }

        // record the tree
        for (int i = 0; i < V; i++) {
            S->ptrees.push_back(t.parent[i]);
            S->ages.push_back(t.age[i]);
        }
        for (int q = 0; q < 4; q++) S->sprs.push_back(spr[q]);
        S->blocklens.push_back(blocklen);

        // mutations on this block (sim.py:297-323), JC-style base change
        for (int node = 0; node < t.V; node++) {
            if (t.parent[node] == -1) continue;
            double blen = std::max(times[t.age[t.parent[node]]] - times[t.age[node]], mintime);
            int nm = rng.poisson(mu * blen * blocklen);
            if (nm == 0) continue;
            leaves.clear();
            collect_leaves(t, node, leaves);
            for (int q = 0; q < nm; q++) {
                int site = pos + rng.below(blocklen);
                unsigned char cur = S->seqs[(size_t) leaves[0] * nsites + site];
                unsigned char nb2;
                do { nb2 = bases[rng.below(4)]; } while (nb2 == cur);
                for (size_t l = 0; l < leaves.size(); l++)
                    S->seqs[(size_t) leaves[l] * nsites + site] = nb2;
            }
        }

        // the extra sequence: copy the leaf the thread currently joins above,
        // plus private mutations on its own branch
        if (first || !branch_at(t, th_node, th_time, ntimes) || rng.uniform() < 0.25)
            pick_thread_state(t);
        first = false;
        {
            leaves.clear();
            collect_leaves(t, th_node, leaves);
            int src = leaves[rng.below((int) leaves.size())];
            unsigned char *dst = &S->seqs[(size_t) nleaves * nsites];
            memcpy(dst + pos, &S->seqs[(size_t) src * nsites + pos], blocklen);
            double blen = std::max(times[th_time], mintime) * 2.0;
            int nm = rng.poisson(mu * blen * blocklen);
            for (int q = 0; q < nm; q++) {
                int site = pos + rng.below(blocklen);
                unsigned char nb2;
                do { nb2 = bases[rng.below(4)]; } while (nb2 == dst[site]);
                dst[site] = nb2;
            }
        }

        pos += blocklen;
        if (pos >= nsites) break;

        // sample the next SPR (sim.py:177-236); retry no-op moves
        for (;;) {
            lineage_counts(t, ntimes, nbranches);
            const int root_age = t.age[t.root];
            // recombination time index ~ nbranches[i] * time_step[i], i <= root age
            double tot = 0.0;
            for (int i = 0; i <= root_age; i++)
                tot += nbranches[i] * (times[i + 1] - times[i]);
            double u = rng.uniform() * tot;
            int rt = 0;
            for (int i = 0; i <= root_age; i++) {
                u -= nbranches[i] * (times[i + 1] - times[i]);
                rt = i;
                if (u <= 0) break;
            }
            std::vector<int> cand;
            for (int x = 0; x < t.V; x++)
                if (x != t.root && branch_at(t, x, rt, ntimes)) cand.push_back(x);
            if (cand.empty()) continue;
            int rnode = cand[rng.below((int) cand.size())];

            // re-coalescence time
            int j = rt;
            int last_kj = nbranches[std::max(j - 1, 0)];
            while (j < ntimes - 2) {
                int kj = nbranches[j];
                if (branch_at(t, rnode, j, ntimes) && t.age[t.parent[rnode]] > j) kj--;
                if (kj < 1) kj = 1;
                double A = coal_steps[2 * j] * kj;
                if (j > rt) A += coal_steps[2 * j - 1] * last_kj;
                double pc = 1.0 - exp(-A / (2.0 * S->popsizes[j]));
                if (rng.uniform() < pc) break;
                j++;
                last_kj = kj;
            }
            const int ct = j;

            // re-coalescence branch: not in the pruned subtree's top, not (parent, parent age)
            std::vector<char> excl(t.V, 0);
            {
                std::vector<int> st(1, rnode);
                while (!st.empty()) {
                    int x = st.back();
                    st.pop_back();
                    excl[x] = 1;
                    if (t.age[x] == ct && t.c0[x] != -1) {
                        st.push_back(t.c0[x]);
                        st.push_back(t.c1[x]);
                    }
                }
            }
            const int rparent = t.parent[rnode];
            cand.clear();
            for (int x = 0; x < t.V; x++) {
                if (excl[x]) continue;
                if (!branch_at(t, x, ct, ntimes)) continue;
                if (x == rparent && ct == t.age[rparent]) continue;
                cand.push_back(x);
            }
            if (cand.empty()) continue;
            int cnode = cand[rng.below((int) cand.size())];

            Tree t2 = t;
            apply_spr(t2, rnode, cnode, ct);
            if (t2.parent == t.parent && t2.age == t.age) continue;   // no-op move
            spr[0] = rnode; spr[1] = rt; spr[2] = cnode; spr[3] = ct;
            t = t2;
            break;
        }
    }
    S->ntrees = (int) S->blocklens.size();
    return S;
}

} // namespace

extern "C" {

void *awb_sim_new(int nleaves, int nsites, int ntimes, const double *times,
                  const double *popsizes, double rho, double mu, uint64_t seed)
{
    return simulate(nleaves, nsites, ntimes, times, popsizes, rho, mu, seed);
}

int awb_sim_ntrees(void *h) { return ((Sim *) h)->ntrees; }
int awb_sim_nnodes(void *h) { return ((Sim *) h)->V; }

void awb_sim_copy(void *h, int *ptrees, int *ages, int *sprs, int *blocklens,
                  unsigned char *seqs)
{
    Sim *S = (Sim *) h;
    memcpy(ptrees, S->ptrees.data(), S->ptrees.size() * sizeof(int));
    memcpy(ages, S->ages.data(), S->ages.size() * sizeof(int));
    memcpy(sprs, S->sprs.data(), S->sprs.size() * sizeof(int));
    memcpy(blocklens, S->blocklens.data(), S->blocklens.size() * sizeof(int));
    memcpy(seqs, S->seqs.data(), S->seqs.size());
}

void awb_sim_free(void *h) { delete (Sim *) h; }

} // extern "C"
