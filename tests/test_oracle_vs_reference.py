"""The C oracle against the reference itself on GENERATED problems (the golden
fixtures cover reference-built ARGs; these cover the simulator's, including
the subtree-maintree form used for internal threading).  CPU only; skipped
where oracle/_ref is not built."""

import numpy as np
import pytest

import oracle_lib as ol
import ref_lib
from argweaver_b200 import sim
from helpers import assert_close

pytestmark = pytest.mark.skipif(not ref_lib.available(),
                                reason="oracle/_ref/ref_bench not built")


@pytest.mark.parametrize("k,n,T,internal,seed,stride",
                         [(8, 3000, 20, False, 1, 1), (8, 3000, 20, True, 2, 1),
                          (20, 4000, 20, True, 3, 7), (30, 2000, 40, False, 4, 5),
                          (12, 2500, 30, True, 5, 1)])
def test_oracle_equals_reference(k, n, T, internal, seed, stride, libc_rand):
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal)
    r = ref_lib.run_reference(d, rand_seed=50 + seed, fw_stride=stride)
    o = ol.run_oracle(d, libc_rand(50 + seed, n))
    assert np.array_equal(r["nstates"], o["nstates"])
    sites = r["fw_sites"]
    assert sites[0] == 0 and sites[-1] == n - 1
    mine = ref_lib.rows_of(o["fw"], o["fw_off"], o["nstates"], d["blocklens"], sites)
    assert_close(mine, r["fw"], "fw rows", 1e-12)
    assert np.array_equal(o["path"], r["path"])
