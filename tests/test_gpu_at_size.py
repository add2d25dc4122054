"""Parity AT SIZE against the reference itself (oracle/_ref/ref_bench = the
unmodified arghmm_forward_alg + stochastic_traceback, sample_thread.cpp:394-460,
522-569) on the BASELINE shapes: forward rows within 1e-9 relative, sampled
paths identical for the same libc rand() draws, for the whole-table mode and
for the checkpointed table with its default 64 MiB segments -- the path the
benchmark runs.  logZ (which the reference does not compute) is checked against
the pinned C oracle."""

import numpy as np
import pytest

import oracle_lib as ol
import ref_lib
from argweaver_b200 import api, sim
from helpers import assert_close, first_divergence

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_lib.available(),
                                 reason="oracle/_ref/ref_bench not built")]

RTOL = 1e-9


def check_against_reference(d, n, seed, libc_rand, stride, checkpoint_too=True,
                            oracle_logz=False, expect_segments=None):
    ref = ref_lib.run_reference(d, rand_seed=seed, fw_stride=stride)
    r = libc_rand(seed, n)
    full = api.Batch([d])
    full.upload().setup().forward().traceback([r]).sync()
    assert full.status() == -1
    assert np.array_equal(full.nstates(), ref["nstates"])
    lay = full.layout()
    mine = ref_lib.rows_of(full.fw(), lay["fw_off"], ref["nstates"],
                           d["blocklens"], ref["fw_sites"])
    assert_close(mine, ref["fw"], "forward rows vs the reference", RTOL)
    div = first_divergence(full.path(), ref["path"])
    assert div is None, ("whole table: path leaves the reference's at site %d "
                         "of %d" % (div, n))
    logz = full.logz()
    full.close()
    if oracle_logz:
        o = ol.run_oracle(d, r)
        assert abs(logz - o["logZ"]) <= RTOL * abs(o["logZ"])
        assert np.array_equal(o["path"], ref["path"])
    if checkpoint_too:
        ck = api.Batch([d], checkpoint=True)
        ck.upload().setup().forward().traceback([r]).sync()
        assert ck.status() == -1
        if expect_segments is not None:
            assert ck.segments()[0] >= expect_segments
        div = first_divergence(ck.path(), ref["path"])
        assert div is None, ("checkpointed table: path leaves the reference's at "
                             "site %d of %d" % (div, n))
        assert abs(ck.logz() - logz) <= RTOL * abs(logz)
        ck.close()


@pytest.mark.parametrize("internal", [False, True])
def test_config2_full_size(internal, libc_rand):
    """BASELINE configs[1]: k=20, L=1 Mb, c=10 (1e5 sites), ntimes=20"""
    n = 100000
    d = sim.simulate_problem(20, n, ntimes=20, seed=201 + int(internal),
                             internal=internal)
    check_against_reference(d, n, 301 + int(internal), libc_rand, stride=10,
                            oracle_logz=True)


@pytest.mark.parametrize("internal", [False, True])
def test_config3_full_size(internal, libc_rand):
    """BASELINE configs[2] = the bench workload: k=50, L=10 Mb, c=10 (1e6
    sites), ntimes=20, leaf and subtree threading; the checkpointed run has the
    bench's ~27 segments per window"""
    n = 1000000
    d = sim.simulate_problem(50, n, ntimes=20, seed=211 + int(internal),
                             internal=internal)
    check_against_reference(d, n, 311 + int(internal), libc_rand, stride=500,
                            oracle_logz=not internal, expect_segments=20)


@pytest.mark.parametrize("internal", [False, True])
def test_config4_shape(internal, libc_rand):
    """BASELINE configs[3] shape: k=100, ntimes=40 (> 1 000 states per block),
    1e5 sites"""
    n = 100000
    d = sim.simulate_problem(100, n, ntimes=40, seed=221 + int(internal),
                             internal=internal)
    check_against_reference(d, n, 321 + int(internal), libc_rand, stride=200,
                            checkpoint_too=False)


def test_batch_of_full_size_windows_checkpointed(libc_rand):
    """several full-size windows in one checkpointed batch (what one bench step
    is, with fewer windows): every window's path equals the reference's"""
    n = 400000
    ds = [sim.simulate_problem(50, n, ntimes=20, seed=231 + i, internal=bool(i & 1))
          for i in range(4)]
    rs = [libc_rand(331 + i, n) for i in range(4)]
    ck = api.Batch(ds, checkpoint=True)
    ck.upload().setup().forward().traceback(rs).sync()
    for i, d in enumerate(ds):
        ref = ref_lib.run_reference(d, rand_seed=331 + i, fw_stride=100000)
        div = first_divergence(ck.path(i), ref["path"])
        assert div is None, "window %d diverges at site %d" % (i, div)
    ck.close()
