"""Host logic and the product's __host__ __device__ setup workers, run on the
CPU through tests/host_emul.cpp and compared with the oracle: state
enumeration, lineage counts, transition vectors, switch matrices, emissions,
and the forward table reproduced from the device tables."""

import numpy as np
import pytest

import emul_lib as el
import oracle_lib as ol
from argweaver_b200 import sim
from conftest import golden_files, load_golden
from helpers import assert_close, block_starts

CASES = [(8, 600, 20, False, 1), (8, 600, 20, True, 2), (20, 1500, 20, False, 3),
         (20, 1500, 20, True, 4), (12, 800, 30, False, 5), (6, 500, 10, True, 6),
         (2, 200, 20, False, 7), (3, 200, 20, True, 8), (30, 800, 40, False, 9),
         # few states spread over many time rows (the forward kernel's padded
         # time-major column is sized by the host layout)
         (4, 400, 40, False, 13), (6, 400, 64, True, 14), (3, 300, 64, False, 15)]


def check_problem(d):
    o = ol.run_oracle(d)
    e = el.Emul(d).setup()
    lay = e.layout()
    B = len(o["nstates"])
    T = int(np.ravel(d["ntimes"])[0])
    for k in ["nstates", "row_off", "fw_off", "sw1_off"]:
        assert np.array_equal(lay[k], o[k]), k
    rows = o["row_off"]
    stn, stt = e.get("st_node"), e.get("st_time")
    for b in range(B):
        S = o["nstates"][b]
        s = o["states"][o["state_off"][b]:o["state_off"][b] + S]
        assert np.array_equal(stn[rows[b]:rows[b] + S], s[:, 0])
        assert np.array_equal(stt[rows[b]:rows[b] + S], s[:, 1])
    lin = e.get("lineages").reshape(B, 3, T)
    assert np.array_equal(lin[:, 0], o["nbranches"])
    assert np.array_equal(lin[:, 1], o["nrecombs"])
    assert np.array_equal(lin[:, 2], o["ncoals"])
    tv = e.get("tmvec").reshape(B, 9, T)
    for k, nm in enumerate(ol.TM_NAMES):
        assert_close(tv[:, k], o[nm], nm, rtol=1e-12)
    assert np.array_equal(e.get("tm_minage"), o["tm_minage"])
    assert np.array_equal(e.get("sw_determ"), o["sw_determ"])
    assert_close(e.get("sw_determprob"), o["sw_determprob"], "determprob", 1e-12)
    assert_close(e.get("sw_recombrow"), o["sw_recombrow"], "recombrow", 1e-12)
    assert_close(e.get("sw_recoalrow"), o["sw_recoalrow"], "recoalrow", 1e-12)
    assert np.array_equal(e.get("sw_recombsrc")[1:], o["sw_recombsrc"][1:])
    assert np.array_equal(e.get("sw_recoalsrc")[1:], o["sw_recoalsrc"][1:])

    # time-major permutation is a permutation sorted by time, stable by state
    perm = e.get("perm")
    for b in range(B):
        S = o["nstates"][b]
        if S == 0:
            continue
        p = perm[rows[b]:rows[b] + S].astype(int)
        assert sorted(p) == list(range(S))
        t = stt[rows[b]:rows[b] + S][p]
        assert np.all(np.diff(t) >= 0)
        same = np.diff(t) == 0
        assert np.all(np.diff(p)[same] > 0)

    # emissions: variant rows live in the forward table slab, invariant
    # sites use inv_emit, masked sites emit 1
    kind, fw0, inv = e.get("kind"), e.get("fw"), e.get("inv_emit")
    bs = block_starts(d["blocklens"])
    for b in range(B):
        S = o["nstates"][b]
        if S == 0:
            continue
        for i in range(max(bs[b], 1), bs[b + 1]):
            lo = o["fw_off"][b] + (i - bs[b]) * S
            ref = o["emit"][lo:lo + S]
            if kind[i] == 1:
                mine = fw0[lo:lo + S]
            elif kind[i] == 0:
                mine = inv[rows[b]:rows[b] + S]
            else:
                mine = np.ones(S)
            assert_close(mine, ref, "emit site %d" % i, rtol=1e-12)

    logz = e.forward()
    assert_close(e.get("fw"), o["fw"], "fw", rtol=1e-11)
    assert abs(logz - o["logZ"]) <= 1e-9 * abs(o["logZ"])


@pytest.mark.parametrize("k,n,T,internal,seed", CASES)
def test_setup_workers_generated(k, n, T, internal, seed):
    check_problem(sim.simulate_problem(k, n, ntimes=T, seed=seed,
                                       internal=internal))


def _raw(e, name):
    nb = el.lib().emul_array_bytes(e.h, name.encode())
    out = np.empty(nb, np.uint8)
    assert el.lib().emul_get(e.h, name.encode(), out.ctypes.data, nb) == 0
    return out


@pytest.mark.parametrize("k,n,T,internal,seed",
                         [(6, 700, 20, 0, 1), (9, 900, 20, 1, 2), (33, 6000, 20, 0, 3),
                          (20, 6000, 40, 1, 4), (12, 4000, 64, 0, 5)])
def test_block_setup_without_generic_tables(k, n, T, internal, seed, monkeypatch):
    # batches on the fast forward kernel run K1 with need_band = 0 (no perm, no
    # band, one running slot counter per time row): every table the fast path
    # reads is byte-identical to the full run's
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal)
    monkeypatch.delenv("AWB_EMUL_NO_BAND", raising=False)
    a = el.Emul(d).setup()
    monkeypatch.setenv("AWB_EMUL_NO_BAND", "1")
    b = el.Emul(d).setup()
    for name in ["st_node", "st_time", "inv_emit", "tmvec", "rowstart", "tmap", "iperm",
                 "lin", "node_first", "node_cnt", "child0", "child1", "order", "root",
                 "lineages", "treelen", "tm_minage", "tmatrix", "sw_start", "sw_cnt",
                 "sw_src", "sw_prob"]:
        assert np.array_equal(_raw(a, name), _raw(b, name)), name
    assert not np.array_equal(_raw(a, "perm"), _raw(b, "perm")) or k <= 6


@pytest.mark.parametrize("maxlen,in_order", [(19, 0), (39, 0), (64, 0), (19, 1), (64, 1)])
def test_thread_map_packing(maxlen, in_order):
    # the forward kernel's thread map: every state exactly once, a branch in
    # consecutive slots of one warp (33..64 states: two whole warps from an even
    # one), everything else empty, nothing written past the map
    import ctypes as C
    L = el.lib()
    L.emul_pack_branches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int]
    rs = np.random.RandomState(maxlen + in_order)
    for rep in range(300):
        V = int(rs.randint(1, 400))
        cnt = np.where(rs.rand(V) < 0.25, 0, rs.randint(1, maxlen + 1, V)).astype(np.int16)
        over = np.cumsum(cnt) > 2048                      # at most 2048 states
        cnt[over] = 0
        S = int(cnt.sum())
        if S == 0:
            continue
        nfirst = np.where(cnt > 0, np.cumsum(cnt) - cnt, -1).astype(np.int16)
        n = L.emul_pack_branches(cnt.ctypes.data, nfirst.ctypes.data, V, in_order, None, 0)
        assert n >= S and n > 0 and (in_order or n % 32 == 0)
        cap = (n + 31) // 32 * 32 + 32 * int(rs.randint(0, 3))
        tmap = np.full(cap + 8, 0x1234, np.uint16)
        assert L.emul_pack_branches(cnt.ctypes.data, nfirst.ctypes.data, V, in_order,
                                    tmap.ctypes.data, cap) == n
        assert np.all(tmap[cap:] == 0x1234)
        m = tmap[:cap]
        used = m[m != 0xFFFF]
        assert sorted(used.tolist()) == list(range(S))
        slot_of = np.full(S, -1)
        slot_of[m[m != 0xFFFF]] = np.nonzero(m != 0xFFFF)[0]
        for i in np.nonzero(cnt > 0)[0]:
            sl = slot_of[nfirst[i]:nfirst[i] + cnt[i]]
            assert np.all(np.diff(sl) == 1)
            if cnt[i] > 32:
                assert sl[0] % 64 == 0
                lo = sl[0]
                assert np.all(m[lo + cnt[i]:lo + 64] == 0xFFFF)
            else:
                assert sl[0] // 32 == sl[-1] // 32


@pytest.mark.parametrize("k,T,popsize", [(30, 64, 200.), (24, 64, 100.)])
def test_setup_workers_skewed_time_rows(k, T, popsize):
    # all lineages coalesce in the first few of 63 time rows: one wide row and
    # many one-state rows, the worst case for the forward kernel's padded
    # time-major column (awb_scribe_plan sizes it on the host)
    check_problem(sim.simulate_problem(k, 300, ntimes=T, seed=5, popsize=popsize))


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1])
def test_setup_workers_golden(path):
    check_problem(load_golden(path))


def test_masked_sites():
    d = sim.simulate_problem(6, 300, seed=21)
    d["seqs"] = d["seqs"].copy()
    d["seqs"][:, 40:60] = ord("N")       # fully masked run
    d["seqs"][2, 100:120] = ord("N")     # partially missing data
    check_problem(d)


def test_layout_rejects_bad_input():
    d = sim.simulate_problem(6, 100, seed=22)
    bad = dict(d)
    bad["ages"] = d["ages"].copy()
    bad["ages"][0, -1] = 19               # node at the top time point
    with pytest.raises(ValueError):
        el.Emul(bad)
    bad = dict(d)
    bad["mappings"] = d["mappings"].copy()
    if bad["mappings"].shape[0] > 1:
        bad["mappings"][1, 0] = -1 if bad["mappings"][1, 0] != -1 else 0
        with pytest.raises(ValueError):
            el.Emul(bad)


def test_layout_rejects_malformed_trees():
    """one-child nodes, leaves not listed first, equal-age parent cycles and
    SPR times beyond the time grid are refused on the host (they would corrupt
    the host walk or index device tables out of bounds)"""
    for internal in (False, True):
        d = sim.simulate_problem(6, 100, seed=23, internal=internal)
        V = d["ptrees"].shape[1]
        nl = (V + 1) // 2
        # a node with exactly one child: move one child of some internal node
        # under a leaf
        bad = dict(d)
        bad["ptrees"] = d["ptrees"].copy()
        kid = int(np.nonzero(d["ptrees"][0] >= nl)[0][0])
        bad["ptrees"][0, kid] = 0 if kid != 0 else 1
        with pytest.raises(ValueError):
            el.Emul(bad)
        # an equal-age cycle between two internal nodes (each keeps two children
        # by swapping a child for the other node)
        bad = dict(d)
        bad["ptrees"] = d["ptrees"].copy()
        bad["ages"] = d["ages"].copy()
        a, b = nl, nl + 1
        bad["ptrees"][0, a] = b
        bad["ptrees"][0, b] = a
        bad["ages"][0, a] = bad["ages"][0, b] = 3
        with pytest.raises(ValueError):
            el.Emul(bad)
        # SPR coalescence time beyond the grid
        if d["sprs"].shape[0] > 1:
            bad = dict(d)
            bad["sprs"] = d["sprs"].copy()
            bad["sprs"][1, 3] = int(np.ravel(d["ntimes"])[0]) + 5
            with pytest.raises(ValueError):
                el.Emul(bad)


@pytest.mark.parametrize("popsize,expect", [(1e4, False), (200., False), (40., True)])
def test_layout_flags_linear_domain_overflow(popsize, expect):
    """exp(lnB) overflows once the cumulative coalescent rate passes ~709
    (ntimes=20, maxtime 2e5, popsize <= 60): the host bound must flag every
    problem whose linear-domain vectors are not finite (the batch then avoids
    the kernels that read them), and must not flag ordinary ones"""
    d = sim.simulate_problem(8, 300, ntimes=20, seed=31, popsize=popsize)
    e = el.Emul(d)
    assert e.lin_unsafe() == expect
    e.setup()
    m = e.lin_max()
    if not e.lin_unsafe():
        assert np.isfinite(m)
    if not np.isfinite(m):
        assert e.lin_unsafe()
    # the tables the generic path uses stay finite and match the oracle
    check_problem(d)
