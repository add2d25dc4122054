"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the
reference's golden vectors.  Gates: integer structures bit-exact; forward
tables, emissions, transition/switch tables and logZ within 1e-9 relative;
sampled paths identical for identical rand() draws."""

import numpy as np
import pytest

import oracle_lib as ol
from argweaver_b200 import api, sim
from conftest import golden_files, load_golden
from helpers import assert_close, first_divergence

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def run_gpu(d, rand_ints, keep_debug=True):
    b = api.Batch([d], keep_debug=keep_debug)
    b.upload().setup().forward().traceback([rand_ints]).sync()
    return b


def compare_with_oracle(d, rand_ints, o=None):
    if o is None:
        o = ol.run_oracle(d, rand_ints)
    b = run_gpu(d, rand_ints)
    B = len(o["nstates"])
    T = int(np.ravel(d["ntimes"])[0])
    assert np.array_equal(b.nstates(), o["nstates"])
    lay = b.layout()
    for k in ["row_off", "fw_off", "sw1_off"]:
        assert np.array_equal(lay[k], o[k]), k
    rows = o["row_off"]
    stn, stt = b.debug("st_node"), b.debug("st_time")
    for blk in range(B):
        S = o["nstates"][blk]
        s = o["states"][o["state_off"][blk]:o["state_off"][blk] + S]
        assert np.array_equal(stn[rows[blk]:rows[blk] + S], s[:, 0])
        assert np.array_equal(stt[rows[blk]:rows[blk] + S], s[:, 1])
    lin = b.debug("lineages").reshape(B, 3, T)
    assert np.array_equal(lin[:, 0], o["nbranches"])
    assert np.array_equal(lin[:, 1], o["nrecombs"])
    assert np.array_equal(lin[:, 2], o["ncoals"])
    tv = b.debug("tmvec").reshape(B, 9, T)
    for k, nm in enumerate(ol.TM_NAMES):
        assert_close(tv[:, k], o[nm], nm, RTOL)
    assert np.array_equal(b.debug("sw_determ"), o["sw_determ"])
    assert_close(b.debug("sw_determprob"), o["sw_determprob"], "determprob", RTOL)
    assert_close(b.debug("sw_recombrow"), o["sw_recombrow"], "recombrow", RTOL)
    assert_close(b.debug("sw_recoalrow"), o["sw_recoalrow"], "recoalrow", RTOL)
    assert np.array_equal(b.debug("sw_recombsrc")[1:], o["sw_recombsrc"][1:])
    assert np.array_equal(b.debug("sw_recoalsrc")[1:], o["sw_recoalsrc"][1:])

    assert b.status() == -1
    assert_close(b.fw(), o["fw"], "fw", RTOL)
    assert abs(b.logz() - o["logZ"]) <= RTOL * abs(o["logZ"])
    path = b.path()
    div = first_divergence(path, o["path"])
    assert div is None, "path diverges from the oracle at site %d" % div
    b.close()
    return o


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1])
def test_golden_vectors(path):
    g = load_golden(path)
    o = compare_with_oracle(g, g["rand_ints"])
    # and directly against what the reference itself produced
    b = run_gpu(g, g["rand_ints"], keep_debug=False)
    assert_close(b.fw(), g["fw"], "fw vs reference", RTOL)
    assert np.array_equal(b.path(), g["path"])
    b.close()
    assert np.array_equal(o["path"], g["path"])


CASES = [(8, 2000, 20, False, 1), (8, 2000, 20, True, 2), (20, 5000, 20, False, 3),
         (20, 5000, 20, True, 4), (12, 3000, 30, False, 5), (6, 1500, 10, True, 6),
         (2, 500, 20, False, 7), (3, 500, 20, True, 8), (40, 3000, 40, False, 9),
         (50, 4000, 20, True, 10),
         # BASELINE config 4 shape: > 1024 states per block (two states per thread
         # in the generic forward kernel, multi-pass rows in the traceback)
         (100, 1500, 40, False, 11), (100, 1200, 40, True, 12),
         # few states spread over many time rows (the forward kernel's padded
         # time-major column is sized by the host layout)
         (4, 800, 40, False, 13), (6, 800, 64, True, 14), (3, 600, 64, False, 15),
         (12, 800, 64, False, 16)]


@pytest.mark.parametrize("k,n,T,internal,seed", CASES)
def test_generated_problems(k, n, T, internal, seed, libc_rand):
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal)
    compare_with_oracle(d, libc_rand(100 + seed, n))


@pytest.mark.parametrize("k,T,popsize", [(30, 64, 200.), (24, 64, 100.)])
def test_skewed_time_rows(k, T, popsize, libc_rand):
    # all lineages coalesce in the first few of 63 time rows (see
    # test_host_logic.test_setup_workers_skewed_time_rows)
    d = sim.simulate_problem(k, 1000, ntimes=T, seed=5, popsize=popsize)
    compare_with_oracle(d, libc_rand(77, 1000))


def test_generic_forward_kernel_forced(libc_rand, monkeypatch):
    """The generic kernel (normally only used for state spaces the fast kernel
    does not cover) on a shape the fast kernel would take."""
    monkeypatch.setenv("AWB_FORCE_GENERIC", "1")
    d = sim.simulate_problem(20, 3000, ntimes=20, seed=21, internal=True)
    compare_with_oracle(d, libc_rand(121, 3000))


def test_masked_and_missing_data(libc_rand):
    d = sim.simulate_problem(6, 600, seed=21)
    d["seqs"] = d["seqs"].copy()
    d["seqs"][:, 40:60] = ord("N")
    d["seqs"][2, 100:120] = ord("N")
    compare_with_oracle(d, libc_rand(5, 600))


def test_prior_given_and_last_state_given(libc_rand):
    """cond_sample_arg_thread semantics (sample_thread.cpp:700-865): one-hot
    first column supplied by the caller, last state fixed"""
    d = sim.simulate_problem(10, 800, seed=31)
    r = libc_rand(9, 800)
    run = ol.OracleRun(d).setup()
    S0 = run.outputs()["nstates"][0]
    prior = np.zeros(S0)
    prior[S0 // 2] = 1.0
    run.forward(prior)
    last = int(run.outputs()["nstates"][-1] // 3)
    run.traceback(r, last_state=last)
    o = run.outputs()
    b = api.Batch([d])
    b.upload().setup().forward([prior]).traceback([r], last_states=[last]).sync()
    assert_close(b.fw(), o["fw"], "fw", RTOL)
    assert np.array_equal(b.path(), o["path"])
    assert b.path()[-1] == last
    b.close()


def test_batch_of_independent_chains(libc_rand):
    """several problems of different shapes in one batch = independent windows"""
    specs = [(8, 700, 20, False, 41), (12, 900, 20, True, 42), (5, 300, 20, False, 43),
             (16, 1100, 20, True, 44)]
    ds = [sim.simulate_problem(k, n, ntimes=T, seed=s, internal=i)
          for (k, n, T, i, s) in specs]
    rs = [libc_rand(s, n) for (k, n, T, i, s) in specs]
    b = api.Batch(ds)
    b.upload().setup().forward().traceback(rs).sync()
    for c, (d, r) in enumerate(zip(ds, rs)):
        o = ol.run_oracle(d, r)
        assert_close(b.fw(c), o["fw"], "fw chain %d" % c, RTOL)
        assert np.array_equal(b.path(c), o["path"])
        assert abs(b.logz(c) - o["logZ"]) <= RTOL * abs(o["logZ"])
    b.close()


def test_full_size_properties(libc_rand):
    """BASELINE config 2 size (k=20, 1 Mb at c=10 -> 1e5 sites): properties that
    do not need the oracle -- columns sum to 1, idempotence, logZ finite.  (The
    comparison with the reference at this and the larger sizes is in
    test_gpu_at_size.py.)"""
    n = 100000
    d = sim.simulate_problem(20, n, seed=51)
    r = libc_rand(77, n)
    b = api.Batch([d])
    b.upload().setup().forward().traceback([r]).sync()
    fw = b.fw()
    lay = b.layout()
    ns = np.maximum(b.nstates(), 1)
    # every column after the first is normalised
    sums = np.add.reduceat(fw, np.concatenate(
        [lay["fw_off"][blk] + np.arange(bl) * ns[blk]
         for blk, bl in enumerate(d["blocklens"])]))
    assert np.all(np.abs(sums[1:] - 1.0) < 1e-12)
    assert np.isfinite(b.logz())
    p1 = b.path()
    assert np.all((p1 >= 0) & (p1 < np.repeat(ns, d["blocklens"])))
    # idempotence: same inputs, same draws -> same table and path
    b.upload().setup().forward().traceback([r]).sync()
    assert np.array_equal(b.fw(), fw)
    assert np.array_equal(b.path(), p1)
    b.close()


def _truncate(d, nblocks=None, blocklens=None):
    """The first blocks of a generated problem, optionally with new block lengths."""
    q = dict(d)
    B = len(d["blocklens"]) if nblocks is None else nblocks
    for key in ("ptrees", "ages", "sprs", "mappings", "subtree_roots"):
        if key in q:
            q[key] = np.ascontiguousarray(d[key][:B])
    bl = np.asarray(d["blocklens"][:B] if blocklens is None else blocklens, np.int32)
    q["blocklens"] = bl
    return q


@pytest.mark.parametrize("internal", [False, True])
def test_degenerate_shapes(libc_rand, internal):
    """ragged / minimal inputs: a one-site window, a window of one block, runs of
    one-site blocks (every site is a breakpoint)"""
    d = sim.simulate_problem(9, 4000, ntimes=12, seed=61, internal=internal)
    B = len(d["blocklens"])
    assert B >= 6
    # one site, one block
    compare_with_oracle(_truncate(d, 1, [1]), libc_rand(1, 1))
    # one block, many sites
    compare_with_oracle(_truncate(d, 1, [257]), libc_rand(2, 257))
    # two sites in two blocks
    compare_with_oracle(_truncate(d, 2, [1, 1]), libc_rand(3, 2))
    # every block one site long, then a long one
    nb = min(B, 40)
    bl = [1] * (nb - 1) + [70]
    compare_with_oracle(_truncate(d, nb, bl), libc_rand(4, sum(bl)))


def test_window_inside_longer_sequences(libc_rand):
    """start_coord > 0: the window is a slice of longer sequence rows
    (arg-sample --region, sequences.cpp:218-247)"""
    d = sim.simulate_problem(7, 900, ntimes=16, seed=71)
    pad = 123
    q = dict(d)
    rng = np.random.default_rng(5)
    left = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=(d["seqs"].shape[0], pad))
    q["seqs"] = np.ascontiguousarray(np.concatenate([left, d["seqs"], left], axis=1))
    q["start_coord"] = np.int32(pad)
    o = ol.run_oracle(d, libc_rand(6, 900))
    b = run_gpu(q, libc_rand(6, 900))
    assert_close(b.fw(), o["fw"], "fw", RTOL)
    assert np.array_equal(b.path(), o["path"])
    b.close()


def test_errors_are_reported_not_computed(libc_rand):
    """malformed inputs fail at the ABI with a message (no partial results)"""
    d = sim.simulate_problem(6, 300, seed=81)
    bad = dict(d)
    bad["blocklens"] = d["blocklens"].copy()
    bad["blocklens"][-1] += 50                       # blocks run past the sequences
    with pytest.raises(api.AwbError):
        api.Batch([bad])
    bad = dict(d)
    bad["ages"] = d["ages"].copy()
    bad["ages"][0, -1] = 0                           # root younger than its children
    with pytest.raises(api.AwbError):
        api.Batch([bad])
    b = api.Batch([d])
    with pytest.raises(api.AwbError):
        b.forward()                                  # setup has not run
    b.close()


@pytest.mark.parametrize("k,n,T,internal,seed,segd",
                         [(8, 3000, 20, False, 91, 20000), (20, 4000, 20, True, 92, 60000),
                          (12, 1500, 30, False, 93, 4000), (50, 3000, 20, True, 94, 200000),
                          (6, 800, 10, True, 95, 1 << 21)])
@pytest.mark.parametrize("resident", ["1", "2", "3", None])
def test_checkpointed_table(k, n, T, internal, seed, segd, resident, libc_rand,
                            monkeypatch):
    """AWB_CHECKPOINT: the forward table is rebuilt segment by segment from
    stored columns; path and logZ must equal the oracle's (and the whole-table
    mode's).  `resident`: segment tables per window -- the last ones of the
    forward pass stay resident and are not rebuilt (None: as many as fit, here
    all of them)."""
    monkeypatch.setenv("AWB_SEG_DOUBLES", str(segd))
    if resident is not None:
        monkeypatch.setenv("AWB_RESIDENT_SEGS", resident)
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal)
    r = libc_rand(200 + seed, n)
    o = ol.run_oracle(d, r)
    b = api.Batch([d], checkpoint=True)
    b.upload().setup().forward().traceback([r]).sync()
    assert b.status() == -1
    assert abs(b.logz() - o["logZ"]) <= RTOL * abs(o["logZ"])
    div = first_divergence(b.path(), o["path"])
    assert div is None, "path diverges from the oracle at site %d" % div
    with pytest.raises(api.AwbError):
        b.fw()
    b.close()


def test_checkpointed_batch_with_given_prior_and_last_state(libc_rand, monkeypatch):
    monkeypatch.setenv("AWB_SEG_DOUBLES", "30000")
    monkeypatch.setenv("AWB_RESIDENT_SEGS", "2")
    specs = [(8, 900, 20, False, 41), (12, 1300, 20, True, 42), (5, 300, 20, False, 43)]
    ds = [sim.simulate_problem(k, n, ntimes=T, seed=s, internal=i)
          for (k, n, T, i, s) in specs]
    rs = [libc_rand(s, n) for (k, n, T, i, s) in specs]
    full = api.Batch(ds)
    S0 = [int(full.nstates(c)[0]) for c in range(len(ds))]
    SL = [int(full.nstates(c)[-1]) for c in range(len(ds))]
    priors = []
    for s0 in S0:
        p = np.zeros(max(s0, 1))
        p[max(s0, 1) // 2] = 1.0
        priors.append(p)
    last = [max(s, 1) // 3 for s in SL]
    full.upload().setup().forward(priors).traceback(rs, last_states=last).sync()
    ck = api.Batch(ds, checkpoint=True)
    ck.upload().setup().forward(priors).traceback(rs, last_states=last).sync()
    for c in range(len(ds)):
        assert np.array_equal(ck.path(c), full.path(c))
        assert abs(ck.logz(c) - full.logz(c)) <= RTOL * abs(full.logz(c))
    full.close()
    ck.close()


@pytest.mark.parametrize("resident", ["1", "3"])
def test_checkpointed_equals_whole_table_at_size(resident, libc_rand, monkeypatch):
    """BASELINE config 3 shape (k=50, T=20), 3e5 sites, leaf and subtree
    threading: the checkpointed table (several 64 MiB segments, 1 or 3 of them
    resident) gives the same paths and logZ as the whole table."""
    monkeypatch.setenv("AWB_RESIDENT_SEGS", resident)
    n = 300000
    ds = [sim.simulate_problem(50, n, seed=71 + i, internal=bool(i)) for i in range(2)]
    rs = [libc_rand(300 + i, n) for i in range(2)]
    full = api.Batch(ds)
    full.upload().setup().forward().traceback(rs).sync()
    ck = api.Batch(ds, checkpoint=True)
    ck.upload().setup().forward().traceback(rs).sync()
    for c in range(2):
        assert full.status(c) == -1 and ck.status(c) == -1
        assert np.array_equal(ck.path(c), full.path(c))
        assert abs(ck.logz(c) - full.logz(c)) <= RTOL * abs(full.logz(c))
    full.close()
    ck.close()


def test_sample_thread_stream_equals_single_batches(libc_rand):
    """api.sample_thread_stream (the next batch's host layout overlaps the
    device work of the current one) gives what one Batch per batch gives."""
    specs = [[(8, 1200, 20, False, 51), (12, 900, 20, True, 52)],
             [(6, 700, 16, True, 53)],
             [(20, 1500, 20, False, 54), (5, 400, 12, False, 55), (9, 800, 20, True, 56)]]
    batches = []
    for group in specs:
        ds = [sim.simulate_problem(k, n, ntimes=T, seed=s, internal=i)
              for (k, n, T, i, s) in group]
        rs = [libc_rand(s, n) for (k, n, T, i, s) in group]
        batches.append((ds, rs))
    got = list(api.sample_thread_stream(batches, checkpoint=True))
    assert len(got) == len(batches)
    for (ds, rs), (paths, logz) in zip(batches, got):
        b = api.Batch(ds)
        b.upload().setup().forward().traceback(rs).sync()
        for c in range(len(ds)):
            assert np.array_equal(paths[c], b.path(c))
            assert abs(logz[c] - b.logz(c)) <= RTOL * abs(b.logz(c))
        b.close()
    # a consumer that stops early leaves nothing behind
    g = api.sample_thread_stream(batches, checkpoint=True)
    next(g)
    g.close()


@pytest.mark.parametrize("popsize", [200., 100., 40.])
@pytest.mark.parametrize("internal", [False, True])
def test_small_population_sizes(popsize, internal, libc_rand):
    """cumulative coalescent rates near and beyond the double exponent range
    (ntimes=20, maxtime 2e5): popsize 200 still runs the fast forward kernel with
    linear-domain vectors ~1e150; for 100 and 40 the host bound routes the batch
    to the generic kernel and the traceback to the closed-form transitions (the
    reference only ever forms exp(lnE2[b] + lnB[a]), which stays finite)"""
    d = sim.simulate_problem(8, 1500, ntimes=20, seed=131, popsize=popsize,
                             internal=internal)
    compare_with_oracle(d, libc_rand(31, 1500))


def test_checkpointed_table_second_traceback(libc_rand, monkeypatch):
    """several tracebacks of one forward pass (upload_rand + traceback, twice):
    after the first one the rebuilt segments have overwritten the first
    resident table, so the second one must rebuild every segment"""
    monkeypatch.setenv("AWB_SEG_DOUBLES", "30000")
    for resident in ("1", "2", "3"):
        monkeypatch.setenv("AWB_RESIDENT_SEGS", resident)
        d = sim.simulate_problem(10, 3000, ntimes=20, seed=141)
        b = api.Batch([d], checkpoint=True)
        b.upload().setup().forward()
        assert b.segments()[0] > 3
        for seed in (7, 8, 9):
            r = libc_rand(seed, 3000)
            o = ol.run_oracle(d, r)
            b.traceback([r]).sync()
            div = first_divergence(b.path(), o["path"])
            assert div is None, ("traceback with draws %d diverges at site %d"
                                 % (seed, div))
        b.close()


@pytest.mark.parametrize("k,n,T,internal,seed", [
    (100, 1500, 40, False, 11),     # BASELINE config 4 shape: four states per thread
    (100, 1200, 40, True, 12),
    (4, 800, 40, False, 13),        # branches of up to 39 states: two register sets
    (12, 800, 64, False, 16),
    (150, 1000, 20, False, 17),     # > 928 states at 20 time points
    (40, 3000, 40, True, 18)])
def test_fast_kernel_covers_large_and_long_shapes(k, n, T, internal, seed, libc_rand):
    """state spaces beyond one state per thread and branches longer than a warp
    run on the register-resident forward kernel (several register sets per
    thread, awb_forward_fast.cuh), whole table and checkpointed"""
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal)
    r = libc_rand(100 + seed, n)
    o = ol.run_oracle(d, r)
    b = run_gpu(d, r, keep_debug=False)
    assert b.fast_path() == "fast"
    assert_close(b.fw(), o["fw"], "fw", RTOL)
    assert abs(b.logz() - o["logZ"]) <= RTOL * abs(o["logZ"])
    assert first_divergence(b.path(), o["path"]) is None
    b.close()


@pytest.mark.parametrize("resident", ["1", "2", "5"])
def test_checkpointed_config4_shape(resident, libc_rand, monkeypatch):
    """k=100, ntimes=40 with a checkpointed table (refused before the fast
    kernel covered the shape)"""
    monkeypatch.setenv("AWB_SEG_DOUBLES", "400000")
    monkeypatch.setenv("AWB_RESIDENT_SEGS", resident)
    d = sim.simulate_problem(100, 4000, ntimes=40, seed=41, internal=True)
    r = libc_rand(141, 4000)
    o = ol.run_oracle(d, r)
    b = api.Batch([d], checkpoint=True)
    b.upload().setup().forward().traceback([r]).sync()
    assert b.segments()[0] >= 6
    assert abs(b.logz() - o["logZ"]) <= RTOL * abs(o["logZ"])
    div = first_divergence(b.path(), o["path"])
    assert div is None, "path diverges at site %d" % div
    b.close()


@pytest.mark.parametrize("resident", ["3", "4", "6"])
def test_checkpointed_pairs_windows_of_different_length(resident, libc_rand, monkeypatch):
    """with three or more tables per window the traceback rebuilds two
    consecutive segments side by side; windows with different segment counts
    in one batch (a segment pair of the longest window straddles the resident
    boundary of a shorter one), repeated tracebacks included"""
    monkeypatch.setenv("AWB_SEG_DOUBLES", "25000")
    monkeypatch.setenv("AWB_RESIDENT_SEGS", resident)
    specs = [(10, 3000, 20, False, 61), (12, 1700, 20, True, 62), (6, 2300, 20, False, 63),
             (10, 2900, 20, True, 64), (8, 400, 20, False, 65)]
    ds = [sim.simulate_problem(k, n, ntimes=T, seed=s, internal=i)
          for (k, n, T, i, s) in specs]
    ck = api.Batch(ds, checkpoint=True)
    ck.upload().setup().forward()
    assert ck.segments()[0] > 8
    for rep in range(2):
        rs = [libc_rand(10 * rep + s, n) for (k, n, T, i, s) in specs]
        ck.traceback(rs).sync()
        for c, d in enumerate(ds):
            o = ol.run_oracle(d, rs[c])
            div = first_divergence(ck.path(c), o["path"])
            assert div is None, "window %d diverges at site %d" % (c, div)
            assert abs(ck.logz(c) - o["logZ"]) <= RTOL * abs(o["logZ"])
    ck.close()
