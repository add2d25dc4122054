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
