"""The alignment given as its variant columns only (awb_problem.var_pos /
var_cols, the form a .sites file holds) gives exactly what the
dense rows give: site kinds, forward table, path, logZ."""
import numpy as np
import pytest

import oracle_lib as ol
from argweaver_b200 import api, sim
from helpers import assert_close, first_divergence

pytestmark = pytest.mark.gpu


def packed_problem(d):
    q = dict(d)
    q.update(api.pack_columns(d["seqs"]))
    del q["seqs"]
    return q


@pytest.mark.parametrize("k,n,T,internal,seed", [(8, 3000, 20, False, 1),
                                                 (20, 5000, 20, True, 2),
                                                 (50, 4000, 20, False, 3),
                                                 (100, 1500, 40, True, 4)])
def test_packed_equals_dense_and_oracle(k, n, T, internal, seed, libc_rand):
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal)
    r = libc_rand(seed, n)
    o = ol.run_oracle(d, r)
    dense = api.Batch([d], keep_debug=True)
    dense.upload().setup().forward().traceback([r]).sync()
    pk = api.Batch([packed_problem(d)], keep_debug=True)
    pk.upload().setup().forward().traceback([r]).sync()
    assert pk.h2d_bytes() < dense.h2d_bytes()
    assert np.array_equal(pk.debug("kind"), dense.debug("kind"))
    assert np.array_equal(pk.fw(), dense.fw())
    assert np.array_equal(pk.path(), dense.path())
    assert pk.logz() == dense.logz()
    assert_close(pk.fw(), o["fw"], "fw vs oracle")
    assert first_divergence(pk.path(), o["path"]) is None
    dense.close()
    pk.close()


def test_packed_with_missing_data_and_start_coord(libc_rand):
    """'N' entries inside variant columns, an all-'N' column (masked site), and a
    window that starts inside the alignment"""
    d = sim.simulate_problem(6, 900, seed=21)
    seqs = d["seqs"].copy()
    seqs[2, 100:140] = ord("N")
    seqs[:, 300:303] = ord("N")
    pad = np.full((seqs.shape[0], 50), ord("A"), np.uint8)
    d["seqs"] = np.concatenate([pad, seqs], axis=1)
    d["start_coord"] = np.int32(50)
    r = libc_rand(9, 900)
    o = ol.run_oracle(d, r)
    pk = api.Batch([packed_problem(d)], keep_debug=True)
    pk.upload().setup().forward().traceback([r]).sync()
    assert_close(pk.fw(), o["fw"], "fw vs oracle")
    assert first_divergence(pk.path(), o["path"]) is None
    kind = pk.debug("kind")
    assert (kind[300:303] == 2).all()
    pk.close()


def test_packed_checkpointed_batch(libc_rand, monkeypatch):
    monkeypatch.setenv("AWB_SEG_DOUBLES", "30000")
    specs = [(8, 1500, False, 31), (12, 2500, True, 32), (5, 700, False, 33)]
    ds = [sim.simulate_problem(k, n, seed=s, internal=i) for (k, n, i, s) in specs]
    rs = [libc_rand(s, n) for (k, n, i, s) in specs]
    full = api.Batch(ds)
    full.upload().setup().forward().traceback(rs).sync()
    ck = api.Batch([packed_problem(d) for d in ds], checkpoint=True)
    ck.upload().setup().forward().traceback(rs).sync()
    for c in range(len(ds)):
        assert np.array_equal(ck.path(c), full.path(c))
        assert abs(ck.logz(c) - full.logz(c)) <= 1e-9 * abs(full.logz(c))
    full.close()
    ck.close()


def test_sites_file_to_device(tmp_path, libc_rand):
    """.sites file -> awb_sites_read -> awb_sites_compress -> packed columns ->
    thread sampling, against the dense rows the reference would have made"""
    k, n = 10, 4000
    d = sim.simulate_problem(k, n, seed=41)
    path = str(tmp_path / "x.sites")
    sim.write_sites(path, d["seqs"], compress=10)
    s = api.Sites.read(path)
    assert s.compress(10)
    q = dict(d)
    del q["seqs"]
    q.update(s.packed())
    assert q["seqlen"] >= n
    r = libc_rand(3, n)
    a = api.Batch([q])
    a.upload().setup().forward().traceback([r]).sync()
    # dense rows as make_sequences_from_sites fills them (default 'A')
    d2 = dict(d)
    d2["seqs"] = s.sequences()
    o = ol.run_oracle(d2, r)
    assert_close(a.fw(), o["fw"], "fw vs oracle")
    assert first_divergence(a.path(), o["path"]) is None
    a.close()
    s.close()
