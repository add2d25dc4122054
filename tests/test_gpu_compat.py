"""Reference-compatible symbols (include/argweaver_b200.h, second half) against
the UNMODIFIED reference library oracle/_ref/libargweaver.so: the same ctypes
calls (argweaver/argweaverc.py:19-349) are made on both libraries with the same
arguments and, for the sampling entry points, the same libc srand() seed."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from helpers import assert_close

pytestmark = pytest.mark.gpu

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libargweaver.so")

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)
c_char_pp = C.POINTER(C.c_char_p)


def rows_int(a):
    a = np.ascontiguousarray(a, np.int32)
    ptrs = (c_int_p * a.shape[0])(*[a[i].ctypes.data_as(c_int_p)
                                     for i in range(a.shape[0])])
    return ptrs, a


def rows_double(a):
    a = np.ascontiguousarray(a, np.float64)
    ptrs = (c_double_p * a.shape[0])(*[a[i].ctypes.data_as(c_double_p)
                                        for i in range(a.shape[0])])
    return ptrs, a


def rows_char(seqs):
    bufs = [C.create_string_buffer(bytes(bytearray(r)), len(r) + 1) for r in seqs]
    ptrs = (C.c_char_p * len(bufs))(*[C.cast(b, C.c_char_p) for b in bufs])
    return ptrs, bufs


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


def bind(lib):
    lib.arghmm_new_trees.restype = C.c_void_p
    lib.arghmm_new_trees.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, c_int_p,
                                     C.c_int, C.c_int, C.c_int]
    lib.delete_local_trees.argtypes = [C.c_void_p]
    lib.get_local_trees_ntrees.argtypes = [C.c_void_p]
    lib.get_local_trees_nnodes.argtypes = [C.c_void_p]
    lib.arghmm_get_nstates.argtypes = [C.c_void_p, C.c_int, C.c_bool, c_int_p]
    lib.get_state_spaces.restype = C.POINTER(c_int_p)
    lib.get_state_spaces.argtypes = [C.c_void_p, C.c_int, C.c_bool]
    lib.delete_state_spaces.argtypes = [C.POINTER(c_int_p), C.c_int]
    lib.arghmm_forward_alg.restype = C.POINTER(c_double_p)
    lib.arghmm_forward_alg.argtypes = [
        C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_double, C.c_double,
        C.c_void_p, C.c_int, C.c_int, C.c_bool, c_double_p, C.c_bool, C.c_bool]
    lib.delete_forward_matrix.argtypes = [C.POINTER(c_double_p), C.c_int]
    lib.arghmm_sample_arg_thread_internal.argtypes = [
        C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_double, C.c_double,
        C.c_void_p, C.c_int, C.c_int, c_int_p]
    lib.arghmm_sample_posterior.restype = c_int_p
    lib.arghmm_sample_posterior.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, c_int_p, C.c_int, C.c_int,
        c_double_p, C.c_int, c_double_p, C.c_double, C.c_double, C.c_void_p,
        C.c_int, C.c_int, c_int_p]
    lib.new_emissions.restype = C.POINTER(c_double_p)
    lib.new_emissions.argtypes = [c_int_p, C.c_int, c_int_p, C.c_int, c_int_p,
                                  C.c_void_p, C.c_int, C.c_int, c_double_p,
                                  C.c_int, C.c_double]
    lib.delete_emissions.argtypes = [C.POINTER(c_double_p), C.c_int]
    lib.new_transition_probs.restype = C.POINTER(c_double_p)
    lib.new_transition_probs.argtypes = [
        C.c_int, c_int_p, c_int_p, C.c_double, c_int_p, C.c_int, C.c_int,
        c_double_p, c_double_p, c_int_p, c_int_p, c_int_p, c_double_p, C.c_double]
    lib.new_transition_probs_switch.restype = C.POINTER(c_double_p)
    lib.new_transition_probs_switch.argtypes = [
        c_int_p, c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p,
        c_int_p, C.c_double, C.c_double, c_int_p, C.c_int, c_int_p, C.c_int,
        C.c_int, c_double_p, c_double_p, c_int_p, c_int_p, c_int_p, c_double_p,
        C.c_double]
    lib.delete_transition_probs.argtypes = [C.POINTER(c_double_p), C.c_int]
    lib.forward_alg.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.backward_alg.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.forward_step.argtypes = [c_double_p, c_double_p, C.c_int, C.c_int,
                                 C.c_void_p, c_double_p]
    lib.sample_hmm_posterior.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         c_int_p]
    lib.sample_hmm_posterior_step.restype = C.c_int
    lib.sample_hmm_posterior_step.argtypes = [C.c_int, C.c_void_p, c_double_p, C.c_int]
    return lib


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libargweaver.so not built (make -C oracle ref)")
    from argweaver_b200 import api
    mine = bind(C.CDLL(api.lib()._name))
    ref = bind(C.CDLL(REF_SO))
    return mine, ref


@pytest.fixture(scope="module")
def libc():
    return C.CDLL("libc.so.6")


def problem(k, nsites, ntimes, seed, internal):
    from argweaver_b200 import sim
    return sim.simulate_problem(k, nsites, ntimes, seed=seed, internal=internal)


def new_trees(lib, p):
    pt, _a = rows_int(p["ptrees"])
    ag, _b = rows_int(p["ages"])
    sp, _c = rows_int(p["sprs"])
    bl = np.ascontiguousarray(p["blocklens"], np.int32)
    t = lib.arghmm_new_trees(pt, ag, sp, iptr(bl), len(bl), p["ptrees"].shape[1],
                             int(p.get("start_coord", 0)))
    assert t
    return t


def matrix(ptr, nrows, ncols_fn):
    out = []
    for i in range(nrows):
        n = ncols_fn(i)
        out.append(np.ctypeslib.as_array(ptr[i], shape=(n,)).copy())
    return out


@pytest.mark.parametrize("internal", [False, True])
def test_trees_and_state_spaces(libs, internal):
    p = problem(7, 300, 12, 5, internal)
    pe = problem(7, 300, 12, 5, False)
    ntimes = len(p["times"])
    nsites = int(np.sum(p["blocklens"]))
    res = []
    for lib in libs:
        t = new_trees(lib, p)
        ntrees = lib.get_local_trees_ntrees(t)
        nnodes = lib.get_local_trees_nnodes(t)
        ns = np.zeros(nsites, np.int32)
        lib.arghmm_get_nstates(t, ntimes, internal, iptr(ns))
        lib.delete_local_trees(t)
        # get_state_spaces ignores `internal` (states.cpp:229-252): always the
        # external state space of each tree, so it is queried on plain trees
        t = new_trees(lib, pe)
        nse = np.zeros(nsites, np.int32)
        lib.arghmm_get_nstates(t, ntimes, False, iptr(nse))
        sp = lib.get_state_spaces(t, ntimes, internal)
        starts = np.concatenate([[0], np.cumsum(pe["blocklens"])[:-1]])
        states = [np.ctypeslib.as_array(sp[b], shape=(int(nse[starts[b]]) * 2,)).copy()
                  for b in range(ntrees)]
        lib.delete_state_spaces(sp, ntrees)
        lib.delete_local_trees(t)
        res.append((ntrees, nnodes, ns, states))
    (nt0, nn0, ns0, st0), (nt1, nn1, ns1, st1) = res
    assert nt0 == nt1 and nn0 == nn1
    assert np.array_equal(ns0, ns1)
    for a, b in zip(st0, st1):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("internal,prior_given", [(False, False), (True, False),
                                                  (False, True)])
def test_arghmm_forward_alg(libs, internal, prior_given):
    p = problem(8, 800, 14, 11, internal)
    ntimes = len(p["times"])
    nsites = int(np.sum(p["blocklens"]))
    seqs, _keep = rows_char(p["seqs"])
    times = np.ascontiguousarray(p["times"], np.float64)
    pops = np.ascontiguousarray(p["popsizes"], np.float64)
    out = []
    for lib in libs:
        t = new_trees(lib, p)
        ns = np.zeros(nsites, np.int32)
        lib.arghmm_get_nstates(t, ntimes, internal, iptr(ns))
        prior = None
        if prior_given:
            rng = np.random.default_rng(3)
            prior = rng.random(max(int(ns[0]), 1))
            prior /= prior.sum()
        fw = lib.arghmm_forward_alg(
            t, dptr(times), ntimes, dptr(pops), float(p["rho"]), float(p["mu"]),
            seqs, p["seqs"].shape[0], p["seqs"].shape[1], prior_given,
            dptr(prior) if prior_given else None, internal, False)
        rows = matrix(fw, nsites, lambda i: max(int(ns[i]), 1))
        lib.delete_forward_matrix(fw, nsites)
        lib.delete_local_trees(t)
        out.append(rows)
    for i, (a, b) in enumerate(zip(*out)):
        assert_close(a, b, name="fw row %d" % i)


def test_arghmm_sample_arg_thread_internal(libs, libc):
    p = problem(8, 1500, 14, 21, True)
    ntimes = len(p["times"])
    nsites = int(np.sum(p["blocklens"]))
    seqs, _keep = rows_char(p["seqs"])
    times = np.ascontiguousarray(p["times"], np.float64)
    pops = np.ascontiguousarray(p["popsizes"], np.float64)
    paths, after = [], []
    for lib in libs:
        t = new_trees(lib, p)
        path = np.zeros(nsites, np.int32)
        libc.srand(4242)
        lib.arghmm_sample_arg_thread_internal(
            t, dptr(times), ntimes, dptr(pops), float(p["rho"]), float(p["mu"]),
            seqs, p["seqs"].shape[0], p["seqs"].shape[1], iptr(path))
        after.append(libc.rand())
        lib.delete_local_trees(t)
        paths.append(path)
    assert np.array_equal(paths[0], paths[1])
    # the caller's rand() stream ends where the reference's does
    assert after[0] == after[1]


def _sample_posterior(lib, libc, p, seed):
    ntimes = len(p["times"])
    nsites = int(np.sum(p["blocklens"]))
    seqs, _keep = rows_char(p["seqs"])
    times = np.ascontiguousarray(p["times"], np.float64)
    pops = np.ascontiguousarray(p["popsizes"], np.float64)
    pt, _a = rows_int(p["ptrees"])
    ag, _b = rows_int(p["ages"])
    sp, _c = rows_int(p["sprs"])
    bl = np.ascontiguousarray(p["blocklens"], np.int32)
    path = np.zeros((nsites, 2), np.int32)
    libc.srand(seed)
    lib.arghmm_sample_posterior(
        pt, ag, sp, iptr(bl), len(bl), p["ptrees"].shape[1], dptr(times),
        ntimes, dptr(pops), float(p["rho"]), float(p["mu"]), seqs,
        p["seqs"].shape[0], p["seqs"].shape[1], iptr(path))
    return path


def test_arghmm_sample_posterior_one_tree(libs, libc):
    """Against the reference on a single local tree.  (With several trees the
    reference's (node,time) conversion loop shadows `end`
    (sample_thread.cpp:958-961) and rewrites the first sites of the path with
    every block's states; this library converts each site with its own block.)"""
    p = problem(6, 900, 10, 31, False)
    one = dict(p)
    nsites = int(np.sum(p["blocklens"]))
    one["ptrees"] = p["ptrees"][:1]
    one["ages"] = p["ages"][:1]
    one["sprs"] = p["sprs"][:1]
    one["blocklens"] = np.array([nsites], np.int32)
    paths = [_sample_posterior(lib, libc, one, 99) for lib in libs]
    assert np.array_equal(paths[0], paths[1])


def test_arghmm_sample_posterior_many_trees(libs, libc, libc_rand):
    """Several local trees: the (node,time) path equals the flat ABI's state
    path (itself pinned to the reference's stochastic_traceback) converted with
    the reference's own state enumeration."""
    from argweaver_b200 import api
    p = problem(6, 900, 10, 31, False)
    mine, ref = libs
    got = _sample_posterior(mine, libc, p, 99)
    nsites = int(np.sum(p["blocklens"]))
    ipath, _logz = api.sample_thread(p, libc_rand(99, nsites))
    t = new_trees(ref, p)
    ntimes = len(p["times"])
    ntrees = ref.get_local_trees_ntrees(t)
    ns = np.zeros(nsites, np.int32)
    ref.arghmm_get_nstates(t, ntimes, False, iptr(ns))
    sp = ref.get_state_spaces(t, ntimes, False)
    starts = np.concatenate([[0], np.cumsum(p["blocklens"])])
    want = np.zeros((nsites, 2), np.int32)
    for b in range(ntrees):
        S = int(ns[starts[b]])
        st = np.ctypeslib.as_array(sp[b], shape=(S * 2,)).reshape(S, 2)
        want[starts[b]:starts[b + 1]] = st[ipath[starts[b]:starts[b + 1]]]
    ref.delete_state_spaces(sp, ntrees)
    ref.delete_local_trees(t)
    assert np.array_equal(got, want)


def test_new_emissions(libs):
    p = problem(9, 400, 12, 41, False)
    ntimes = len(p["times"])
    times = np.ascontiguousarray(p["times"], np.float64)
    seqs, _keep = rows_char(p["seqs"])
    seqlen = p["seqs"].shape[1]
    ptree = np.ascontiguousarray(p["ptrees"][0], np.int32)
    ages = np.ascontiguousarray(p["ages"][0], np.int32)
    nnodes = len(ptree)
    mine, ref = libs
    # states of tree 0 from the reference-compatible state enumeration
    t = new_trees(mine, p)
    ns = np.zeros(int(np.sum(p["blocklens"])), np.int32)
    mine.arghmm_get_nstates(t, ntimes, False, iptr(ns))
    S = int(ns[0])
    sp = mine.get_state_spaces(t, ntimes, False)
    states = np.ctypeslib.as_array(sp[0], shape=(S * 2,)).copy()
    mine.delete_state_spaces(sp, mine.get_local_trees_ntrees(t))
    mine.delete_local_trees(t)
    states = np.ascontiguousarray(states[::-1].reshape(S, 2)[:, ::-1])  # reversed order
    out = []
    for lib in libs:
        e = lib.new_emissions(iptr(states), S, iptr(ptree), nnodes, iptr(ages),
                              seqs, p["seqs"].shape[0], seqlen, dptr(times),
                              ntimes, float(p["mu"]))
        out.append(np.array(matrix(e, seqlen, lambda i: S)))
        lib.delete_emissions(e, seqlen)
    assert_close(out[0], out[1], name="emissions")


def test_new_transition_probs(libs):
    p = problem(9, 400, 12, 43, False)
    ntimes = len(p["times"])
    times = np.ascontiguousarray(p["times"], np.float64)
    pops = np.ascontiguousarray(p["popsizes"], np.float64)
    steps = np.ascontiguousarray(np.diff(np.append(times, times[-1] * 2)))
    mine, ref = libs
    t = new_trees(mine, p)
    ntrees = mine.get_local_trees_ntrees(t)
    sp = mine.get_state_spaces(t, ntimes, False)
    nsv = np.zeros(int(np.sum(p["blocklens"])), np.int32)
    mine.arghmm_get_nstates(t, ntimes, False, iptr(nsv))
    starts = np.concatenate([[0], np.cumsum(p["blocklens"])[:-1]])
    states = [np.ctypeslib.as_array(sp[b], shape=(int(nsv[starts[b]]) * 2,)).copy()
              for b in range(ntrees)]
    mine.delete_state_spaces(sp, ntrees)
    mine.delete_local_trees(t)
    nnodes = p["ptrees"].shape[1]
    dummy = np.zeros(ntimes, np.int32)
    for b in (0, 1, min(3, ntrees - 1)):
        ptree = np.ascontiguousarray(p["ptrees"][b], np.int32)
        ages = np.ascontiguousarray(p["ages"][b], np.int32)
        S = len(states[b]) // 2
        out = []
        for lib in libs:
            m = lib.new_transition_probs(nnodes, iptr(ptree), iptr(ages), 0.0,
                                         iptr(states[b]), S, ntimes, dptr(times),
                                         dptr(steps), iptr(dummy), iptr(dummy),
                                         iptr(dummy), dptr(pops), float(p["rho"]))
            out.append(np.array(matrix(m, S, lambda i: S)))
            lib.delete_transition_probs(m, S)
        assert_close(np.exp(out[0]), np.exp(out[1]), name="transition block %d" % b)
        if b > 0:
            lp = np.ascontiguousarray(p["ptrees"][b - 1], np.int32)
            la = np.ascontiguousarray(p["ages"][b - 1], np.int32)
            S1 = len(states[b - 1]) // 2
            spr = [int(x) for x in p["sprs"][b]]
            out = []
            for lib in libs:
                m = lib.new_transition_probs_switch(
                    iptr(ptree), iptr(lp), nnodes, spr[0], spr[1], spr[2], spr[3],
                    iptr(ages), iptr(la), 0.0, 0.0, iptr(states[b - 1]), S1,
                    iptr(states[b]), S, ntimes, dptr(times), dptr(steps),
                    iptr(dummy), iptr(dummy), iptr(dummy), dptr(pops),
                    float(p["rho"]))
                out.append(np.array(matrix(m, S1, lambda i: S)))
                lib.delete_transition_probs(m, S1)
            assert_close(np.exp(out[0]), np.exp(out[1]), name="switch block %d" % b)


def random_hmm(n, S, seed):
    rng = np.random.default_rng(seed)
    trans = rng.random((S, S)) + 1e-3
    trans /= trans.sum(1, keepdims=True)
    emit = rng.random((n, S)) + 1e-3
    return np.log(trans), np.log(emit)


def test_dense_hmm_forward_backward(libs):
    n, S = 40, 37
    trans, emit = random_hmm(n, S, 7)
    tp, _t = rows_double(trans)
    ep, _e = rows_double(emit)
    res = []
    for lib in libs:
        fw = np.zeros((n, S))
        fw[0] = np.log(1.0 / S) + emit[0]
        fp, fwa = rows_double(fw)
        lib.forward_alg(n, S, tp, ep, fp)
        bw = np.zeros((n, S))
        bp, bwa = rows_double(bw)
        lib.backward_alg(n, S, tp, ep, bp)
        col2 = np.zeros(S)
        lib.forward_step(dptr(fwa[3]), dptr(col2), S, S, tp, dptr(_e[4]))
        res.append((fwa.copy(), bwa.copy(), col2))
    assert_close(res[0][0], res[1][0], name="forward_alg", rtol=1e-9)
    assert_close(res[0][1], res[1][1], name="backward_alg", rtol=1e-9)
    assert_close(res[0][2], res[1][2], name="forward_step", rtol=1e-9)


def test_dense_hmm_sampling(libs, libc):
    n, S = 60, 21
    trans, emit = random_hmm(n, S, 9)
    tp, _t = rows_double(trans)
    ep, _e = rows_double(emit)
    mine, ref = libs
    fw = np.zeros((n, S))
    fw[0] = np.log(1.0 / S) + emit[0]
    fp, fwa = rows_double(fw)
    ref.forward_alg(n, S, tp, ep, fp)
    paths, steps = [], []
    for lib in libs:
        path = np.zeros(n, np.int32)
        path[n - 1] = 5
        libc.srand(17)
        lib.sample_hmm_posterior(n, S, tp, fp, iptr(path))
        steps.append(lib.sample_hmm_posterior_step(S, tp, dptr(fwa[10]), 3))
        paths.append(path)
    assert np.array_equal(paths[0], paths[1])
    assert steps[0] == steps[1]
