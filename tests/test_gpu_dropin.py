"""The reference with its threading path on the device: oracle/_ref/arg-sample-b200
and oracle/_ref/libargweaver_dropin.so are the reference's own objects with the
L2 wrappers of sample_thread.cpp:578-875 replaced by the adapter
oracle/sample_thread_b200.cpp over libargweaver_b200.so (INTEGRATION.md section
2).  Whole MCMC runs with the same seed must reproduce the reference binary's
.stats rows (prior / likelihood / joint / recombs / noncompats / arglen of every
iteration, arg-sample.cpp:444-488): every sampled thread has to be identical
for that, hundreds of them in sequence, each conditioned on the ARG the
previous ones built."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from argweaver_b200 import sim

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref")
ARG_SAMPLE = os.path.join(REF, "arg-sample")
ARG_SAMPLE_B200 = os.path.join(REF, "arg-sample-b200")
SIM1 = os.path.join(ROOT, "tests", "golden", "sim1.sites")

needs_binaries = pytest.mark.skipif(
    not (os.path.exists(ARG_SAMPLE) and os.path.exists(ARG_SAMPLE_B200)),
    reason="oracle/_ref/arg-sample[-b200] not built (make -C oracle ref)")


def run_sampler(binary, sites, out, extra, env=None):
    cmd = [binary, "-s", sites, "-N", "10000", "-r", "1.6e-8", "-m", "1.8e-8",
           "--ntimes", "20", "--maxtime", "200e3", "-x", "1", "-q",
           "-o", out] + extra
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, env=e)
    assert r.returncode == 0, (cmd, r.stderr[-3000:])
    with open(out + ".stats") as f:
        return f.read().splitlines(), r


def assert_same_stats(a, b):
    assert len(a) == len(b) and len(a) > 2
    for i, (x, y) in enumerate(zip(a, b)):
        assert x == y, "row %d differs:\n  reference: %s\n  device:    %s" % (i, x, y)


@needs_binaries
def test_config1_quick_start_stats_identical(tmp_path):
    """BASELINE configs[0]: the README quick start on the reference's own
    examples/sim1/sim1.sites (copied to tests/golden as a data fixture):
    arg-sample -c 10 -n 100 -x 1.  SURVEY appendix C recorded the reference's
    last row on this machine's glibc."""
    extra = ["-c", "10", "-n", "100"]
    ref, _ = run_sampler(ARG_SAMPLE, SIM1, str(tmp_path / "ref"), extra)
    dev, r = run_sampler(ARG_SAMPLE_B200, SIM1, str(tmp_path / "dev"), extra,
                         env={"AWB_ADAPTER_REPORT": "1"})
    assert_same_stats(ref, dev)
    last = ref[-1].split("\t")
    assert last[:2] == ["resample", "100"]
    assert abs(float(last[2]) - (-1989.998121)) < 1e-5
    assert abs(float(last[3]) - (-140342.029327)) < 1e-5
    assert last[5] == "151"
    # the threads did run on the device
    assert "device thread samples" in r.stderr
    n_dev = int(r.stderr.split("device thread samples:")[1].split()[0])
    n_ref = int(r.stderr.split("reference fallbacks:")[1].split()[0])
    assert n_dev >= 100 and n_ref == 0


@needs_binaries
@pytest.mark.parametrize("k,n,iters", [(12, 20000, 12), (20, 50000, 4)])
def test_generated_sites_stats_identical(k, n, iters, tmp_path):
    """generated data (k sequences, n compressed sites at -c 10): leaf
    threading while the ARG is built, then subtree re-threading every
    iteration (resample_arg_all / resample_arg_mcmc_all)"""
    (_t, _p, _r, _m, _pt, _ag, _sp, _bl, seqs) = sim.simulate_arg(
        k - 1, n, 20, seed=500 + k)
    sites = str(tmp_path / "gen.sites")
    sim.write_sites(sites, seqs, compress=10)
    extra = ["-c", "10", "-n", str(iters)]
    ref, _ = run_sampler(ARG_SAMPLE, sites, str(tmp_path / "ref"), extra)
    dev, _ = run_sampler(ARG_SAMPLE_B200, sites, str(tmp_path / "dev"), extra)
    assert_same_stats(ref, dev)


@needs_binaries
def test_infsites_run_stats_identical(tmp_path):
    """--infsites (ArgModel::infsites_penalty = 1e-100, arg-sample.cpp:1077-1079):
    the penalised emissions (emit.cpp:457-589, :848-862) run on the device"""
    extra = ["-c", "10", "-n", "20", "--infsites"]
    ref, _ = run_sampler(ARG_SAMPLE, SIM1, str(tmp_path / "ref"), extra)
    dev, r = run_sampler(ARG_SAMPLE_B200, SIM1, str(tmp_path / "dev"), extra,
                         env={"AWB_ADAPTER_REPORT": "1"})
    assert_same_stats(ref, dev)
    assert int(r.stderr.split("reference fallbacks:")[1].split()[0]) == 0


@needs_binaries
def test_unphased_run_stats_identical(tmp_path):
    """--unphased (ArgModel::unphased): emissions averaged over the two phasings
    of the individual being threaded, PhaseProbs filled from the device, the
    reference's own sample_phase flipping alleles where it calls it -- the data
    itself changes from iteration to iteration, so the .stats rows only agree
    if every phase probability and every draw does"""
    extra = ["-c", "10", "-n", "20", "--unphased"]
    ref, _ = run_sampler(ARG_SAMPLE, SIM1, str(tmp_path / "ref"), extra)
    dev, r = run_sampler(ARG_SAMPLE_B200, SIM1, str(tmp_path / "dev"), extra,
                         env={"AWB_ADAPTER_REPORT": "1"})
    assert_same_stats(ref, dev)
    assert int(r.stderr.split("reference fallbacks:")[1].split()[0]) == 0
    assert int(r.stderr.split("device thread samples:")[1].split()[0]) >= 20


@needs_binaries
def test_host_recombination_sampler_gives_the_same_run(tmp_path):
    """AWB_ADAPTER_HOST_RECOMBS=1: the reference's own sample_recombinations on
    the path the device returns -- same .stats as with the device sampler"""
    extra = ["-c", "10", "-n", "10"]
    ref, _ = run_sampler(ARG_SAMPLE, SIM1, str(tmp_path / "ref"), extra)
    dev, _ = run_sampler(ARG_SAMPLE_B200, SIM1, str(tmp_path / "dev"), extra,
                         env={"AWB_ADAPTER_HOST_RECOMBS": "1"})
    assert_same_stats(ref, dev)


@needs_binaries
def test_forced_fallback_is_the_reference(tmp_path):
    """AWB_ADAPTER_FORCE_REFERENCE routes every call to the renamed reference
    bodies: the binary is then the reference, bit for bit"""
    extra = ["-c", "10", "-n", "5"]
    ref, _ = run_sampler(ARG_SAMPLE, SIM1, str(tmp_path / "ref"), extra)
    dev, r = run_sampler(ARG_SAMPLE_B200, SIM1, str(tmp_path / "dev"), extra,
                         env={"AWB_ADAPTER_FORCE_REFERENCE": "1",
                              "AWB_ADAPTER_REPORT": "1"})
    assert_same_stats(ref, dev)
    assert int(r.stderr.split("device thread samples:")[1].split()[0]) == 0


# ---- the C exports of the drop-in library (argweaverc.py:596-603)

c_int_p = C.POINTER(C.c_int)
c_double_p = C.POINTER(C.c_double)


def _rows_int(a):
    a = np.ascontiguousarray(a, np.int32)
    return (c_int_p * a.shape[0])(*[a[i].ctypes.data_as(c_int_p)
                                     for i in range(a.shape[0])]), a


def _bind(path):
    lib = C.CDLL(path)
    lib.arghmm_new_trees.restype = C.c_void_p
    lib.arghmm_new_trees.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, c_int_p,
                                     C.c_int, C.c_int, C.c_int]
    lib.delete_local_trees.argtypes = [C.c_void_p]
    lib.get_local_trees_ntrees.argtypes = [C.c_void_p]
    lib.get_local_trees_nnodes.argtypes = [C.c_void_p]
    lib.get_local_trees_ptrees.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_void_p, c_int_p]
    lib.arghmm_sample_thread.restype = C.c_void_p
    lib.arghmm_sample_thread.argtypes = [
        C.c_void_p, c_double_p, C.c_int, c_double_p, C.c_double, C.c_double,
        C.c_void_p, C.c_int, C.c_int]
    return lib


def _sample_thread(lib, libc, d, seed):
    pt, _a = _rows_int(d["ptrees"])
    ag, _b = _rows_int(d["ages"])
    sp, _c = _rows_int(d["sprs"])
    bl = np.ascontiguousarray(d["blocklens"], np.int32)
    trees = lib.arghmm_new_trees(pt, ag, sp, bl.ctypes.data_as(c_int_p), len(bl),
                                 d["ptrees"].shape[1], 0)
    seqs = np.ascontiguousarray(d["seqs"], np.uint8)
    bufs = [C.create_string_buffer(bytes(bytearray(r)), len(r) + 1) for r in seqs]
    ptrs = (C.c_char_p * len(bufs))(*[C.cast(b, C.c_char_p) for b in bufs])
    times = np.ascontiguousarray(d["times"], np.float64)
    pops = np.ascontiguousarray(d["popsizes"], np.float64)
    libc.srand(seed)
    lib.arghmm_sample_thread(trees, times.ctypes.data_as(c_double_p), len(times),
                             pops.ctypes.data_as(c_double_p), float(d["rho"]),
                             float(d["mu"]), ptrs, len(bufs), seqs.shape[1])
    B = lib.get_local_trees_ntrees(trees)
    V = lib.get_local_trees_nnodes(trees)
    optr = np.zeros((B, V), np.int32)
    oage = np.zeros((B, V), np.int32)
    ospr = np.zeros((B, 4), np.int32)
    obl = np.zeros(B, np.int32)
    p1, _d = _rows_int(optr)
    p2, _e = _rows_int(oage)
    p3, _f = _rows_int(ospr)
    lib.get_local_trees_ptrees(trees, p1, p2, p3, obl.ctypes.data_as(c_int_p))
    lib.delete_local_trees(trees)
    return _d, _e, _f, obl


def test_dropin_library_arghmm_sample_thread():
    """arghmm_sample_thread (sample_thread.cpp:1010) through ctypes on the
    drop-in library and on the reference library, same srand() seed: the ARG
    with the new chromosome threaded in must be the same ARG"""
    so_ref = os.path.join(REF, "libargweaver.so")
    so_dev = os.path.join(REF, "libargweaver_dropin.so")
    if not (os.path.exists(so_ref) and os.path.exists(so_dev)):
        pytest.skip("drop-in library not built")
    libc = C.CDLL("libc.so.6")
    ref = _bind(so_ref)
    dev = _bind(so_dev)
    for k, n, seed in ((6, 1500, 3), (16, 6000, 4)):
        d = sim.simulate_problem(k, n, ntimes=20, seed=700 + k)
        a = _sample_thread(ref, libc, d, seed)
        b = _sample_thread(dev, libc, d, seed)
        assert a[0].shape == b[0].shape, (a[0].shape, b[0].shape)
        for x, y, name in zip(a, b, ("ptrees", "ages", "sprs", "blocklens")):
            assert np.array_equal(x, y), name
        assert a[0].shape[1] == d["ptrees"].shape[1] + 2
