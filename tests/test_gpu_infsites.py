"""Infinite-sites penalty (ArgModel::infsites_penalty, arg-sample --infsites;
emit.cpp:457-589, :848-862) on the device against the UNMODIFIED reference
(oracle/_ref/ref_bench with the penalty set on its ArgModel): forward rows
within 1e-9, sampled paths identical for the same libc rand() draws.
SURVEY 8f N-4."""
import numpy as np
import pytest

import ref_lib
from argweaver_b200 import api, sim
from helpers import assert_close, first_divergence

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_lib.available(),
                                 reason="oracle/_ref/ref_bench not built")]


@pytest.mark.parametrize("k,n,T,internal,seed,penalty", [
    (8, 4000, 20, False, 1, 1e-2), (8, 4000, 20, True, 2, 1e-2),
    (20, 6000, 20, False, 3, 1e-100), (20, 6000, 20, True, 4, 0.5),
    (50, 5000, 20, True, 5, 1e-3), (30, 3000, 40, False, 6, 1e-2)])
def test_infsites_against_the_reference(k, n, T, internal, seed, penalty, libc_rand):
    # a high mutation rate: recurrent mutations, so that the penalty bites
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal, mu=2.5e-7)
    d["infsites_penalty"] = np.float64(penalty)
    ref = ref_lib.run_reference(d, rand_seed=40 + seed, fw_stride=7)
    r = libc_rand(40 + seed, n)
    b = api.Batch([d])
    b.upload().setup().forward().traceback([r]).sync()
    lay = b.layout()
    mine = ref_lib.rows_of(b.fw(), lay["fw_off"], ref["nstates"], d["blocklens"],
                           ref["fw_sites"])
    assert_close(mine, ref["fw"], "forward rows vs the reference")
    assert first_divergence(b.path(), ref["path"]) is None
    # and the penalty did change something
    d0 = dict(d)
    d0["infsites_penalty"] = np.float64(1.0)
    b0 = api.Batch([d0])
    b0.upload().setup().forward().sync()
    assert b0.logz() > b.logz()
    b0.close()
    b.close()


def test_infsites_with_missing_data_and_packed_columns(libc_rand):
    d = sim.simulate_problem(10, 3000, seed=9, mu=2.5e-7)
    seqs = d["seqs"].copy()
    seqs[3, 200:400] = ord("N")
    seqs[9, 1000:1100] = ord("N")       # the sequence being threaded
    d["seqs"] = seqs
    d["infsites_penalty"] = np.float64(1e-2)
    ref = ref_lib.run_reference(d, rand_seed=77, fw_stride=5)
    r = libc_rand(77, 3000)
    q = dict(d)
    q.update(api.pack_columns(seqs))
    del q["seqs"]
    b = api.Batch([q])
    b.upload().setup().forward().traceback([r]).sync()
    lay = b.layout()
    mine = ref_lib.rows_of(b.fw(), lay["fw_off"], ref["nstates"], d["blocklens"],
                           ref["fw_sites"])
    assert_close(mine, ref["fw"], "forward rows vs the reference")
    assert first_divergence(b.path(), ref["path"]) is None
    b.close()
