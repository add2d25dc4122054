"""awb_sites_* (.sites reader, -c compression, dense rows; host code of
libargweaver_b200.so) against the UNMODIFIED reference's read_sites /
find_compress_cols / compress_sites / make_sequences_from_sites
(sequences.cpp:173-352, :523-609), run by oracle/_ref/ref_sites.  No GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from argweaver_b200 import api, sim
from argweaver_b200.flatfile import read_awf

REF_SITES = os.path.join(ROOT, "oracle", "_ref", "ref_sites")
SIM1 = os.path.join(ROOT, "tests", "golden", "sim1.sites")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_SITES),
                                reason="oracle/_ref/ref_sites not built")


def reference(path, compress, region=None, tmp="/tmp"):
    out = os.path.join(str(tmp), "ref_sites.awf")
    cmd = [REF_SITES, path, out, str(compress)]
    if region:
        cmd += [str(region[0]), str(region[1])]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return read_awf(out)


def check(path, compress, region, tmp):
    ref = reference(path, compress, region, tmp)
    s = api.Sites.read(path, region)
    assert s.nseqs == int(ref["nseqs"][0])
    assert (s.start, s.end) == (int(ref["read_start"][0]), int(ref["read_end"][0]))
    assert np.array_equal(s.positions(), ref["read_positions"])
    assert np.array_equal(s.columns(), ref["cols"].reshape(s.ncols, s.nseqs))
    ok = s.compress(compress)
    assert ok == bool(ref["compress_ok"][0])
    if ok:
        assert (s.start, s.end) == (int(ref["start"][0]), int(ref["end"][0]))
        assert np.array_equal(s.positions(), ref["positions"])
        assert np.array_equal(s.mapping(), ref["all_sites"])
        assert np.array_equal(s.sequences(), ref["seqs"])
    s.close()
    return ok


@pytest.mark.parametrize("compress", [1, 2, 5, 10, 20, 100])
def test_sim1_sites(compress, tmp_path):
    assert check(SIM1, compress, None, tmp_path)


@pytest.mark.parametrize("region", [(20000, 60000), (0, 404), (403, 50000), (99000, 100000)])
def test_sim1_subregions(region, tmp_path):
    check(SIM1, 10, region, tmp_path)


@pytest.mark.parametrize("k,n,c,seed", [(12, 20000, 10, 3), (20, 50000, 10, 4),
                                        (6, 3000, 4, 5), (30, 8000, 1, 6)])
def test_generated_sites(k, n, c, seed, tmp_path):
    seqs = sim.simulate_arg(k - 1, n, 20, seed=seed)[8]
    path = str(tmp_path / "gen.sites")
    sim.write_sites(path, seqs, compress=c)
    assert check(path, c, None, tmp_path)
    # what goes to the device: the variant columns of the compressed alignment
    s = api.Sites.read(path)
    s.compress(c)
    dense = s.sequences()
    var = np.nonzero((seqs != seqs[0]).any(axis=0))[0]
    assert np.array_equal(dense[:, :n], np.where(
        np.isin(np.arange(n), var)[None, :], seqs, ord("A")))
    s.close()


def test_compression_that_does_not_fit(tmp_path):
    """variant sites denser than one per `compress` bases cannot be compressed
    (find_compress_cols returns false, sequences.cpp:576-578)"""
    path = str(tmp_path / "dense.sites")
    with open(path, "w") as f:
        f.write("NAMES\ta\tb\tc\nREGION\tchr\t1\t200\n")
        for p in range(10, 190, 3):
            f.write("%d\t%s\n" % (p, "ACA" if p % 2 else "GGT"))
    assert not check(path, 10, None, tmp_path)
    assert check(path, 2, None, tmp_path)


def test_malformed_files_are_refused(tmp_path):
    bad = [("NAMES\ta\tb\nREGION\tchr\t1\t100\n5\tAC\n3\tAC\n", "sorted"),
           ("NAMES\ta\tb\nREGION\tchr\t1\t100\n5\tACG\n", "does not match"),
           ("NAMES\ta\tb\nREGION\tchr\t1\t100\n5\tAX\n", "invalid sequence"),
           ("NAMES\ta\tb\nRANGE\tchr\t1\t100\n", "RANGE"),
           ("NAMES\ta\tb\nREGION\tchr\tx\n", "REGION")]
    for i, (text, what) in enumerate(bad):
        path = str(tmp_path / ("bad%d.sites" % i))
        open(path, "w").write(text)
        with pytest.raises(api.AwbError) as e:
            api.Sites.read(path)
        assert what in str(e.value)
        r = subprocess.run([REF_SITES, path, str(tmp_path / "x.awf"), "1"],
                           capture_output=True)
        assert r.returncode != 0            # the reference refuses it too
