"""Unphased emissions (second half; reference emit.cpp:705-742,
:834-842, PhaseProbs sequences.h:174-206): at the sites where the two
haplotypes of one individual differ the emission is the mean over the two
phasings, and after the traceback the probability of the given phasing at the
sampled state is what PhaseProbs::sample_phase draws against.  The reference is
run by oracle/_ref/ref_bench with a PhaseProbs for the same two rows.

CPU part: the product's __host__ __device__ emission code on the host
(tests/host_emul.cpp).  GPU part: the whole chain on the device, including the
recombination points sampled after the phase draws."""

import numpy as np
import pytest

import emul_lib
import ref_lib
from argweaver_b200 import api, sim
from helpers import assert_close

needs_ref = pytest.mark.skipif(not ref_lib.available(),
                               reason="oracle/_ref/ref_bench not built")
RTOL = 1e-9

# k, sites, ntimes, internal, rows of the unphased individual, seed
CASES = [(8, 3000, 20, False, (7, 2), 81),      # the new chromosome and a leaf
         (8, 3000, 20, True, (7, 3), 82),       # the subtree leaf and a leaf
         (12, 6000, 20, True, (2, 5), 83),      # two leaves of the main tree
         (10, 4000, 30, False, (9, 0), 84)]


def problem(k, n, T, internal, rows, seed):
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal)
    d["phase_rows"] = np.array(rows, np.int32)
    return d


@needs_ref
@pytest.mark.parametrize("k,n,T,internal,rows,seed", CASES)
def test_host_emissions_match_reference(k, n, T, internal, rows, seed, libc_rand):
    d = problem(k, n, T, internal, rows, seed)
    ref = ref_lib.run_reference(d, rand_seed=seed, fw_stride=1)
    e = emul_lib.Emul(d).setup()
    e.forward()
    lay = e.layout()
    mine = ref_lib.rows_of(e.get("fw"), lay["fw_off"], ref["nstates"], d["blocklens"],
                           ref["fw_sites"])
    assert_close(mine, ref["fw"], "forward table with unphased emissions", RTOL)
    # the heterozygous sites and the phase probabilities at the reference's path
    pp = e.phase_probs(ref["path"])
    het = np.nonzero(pp >= 0)[0]
    assert np.array_equal(het, ref["phase_pos"])
    assert len(het) > 0
    assert_close(pp[het], ref["phase_p"], "phase probabilities", RTOL)
    # phased run of the same problem differs (the mode is really on)
    d2 = dict(d)
    d2.pop("phase_rows")
    ref2 = ref_lib.run_reference(d2, rand_seed=seed, fw_stride=1)
    assert not np.allclose(ref2["fw"], ref["fw"], rtol=1e-6, atol=0)


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("checkpoint", [False, True])
@pytest.mark.parametrize("k,n,T,internal,rows,seed", CASES)
def test_device_matches_reference(k, n, T, internal, rows, seed, checkpoint, libc_rand,
                                  monkeypatch):
    if checkpoint:
        monkeypatch.setenv("AWB_SEG_DOUBLES", str(1 << 16))
    d = problem(k, n, T, internal, rows, seed)
    ref = ref_lib.run_reference(d, rand_seed=seed, fw_stride=1)
    r = libc_rand(seed, n)
    st = api.libc_rand_snapshot()
    b = api.Batch([d], checkpoint=checkpoint)
    b.upload().setup().forward().traceback([r]).phase_probs().sync()
    if not checkpoint:
        mine = ref_lib.rows_of(b.fw(), b.layout()["fw_off"], ref["nstates"],
                               d["blocklens"], ref["fw_sites"])
        assert_close(mine, ref["fw"], "forward table with unphased emissions", RTOL)
    assert np.array_equal(b.path(), ref["path"])
    pp = b.get_phase_probs()
    het = np.nonzero(pp >= 0)[0]
    assert np.array_equal(het, ref["phase_pos"])
    assert_close(pp[het], ref["phase_p"], "phase probabilities", RTOL)
    # PhaseProbs::sample_phase takes one frand() per heterozygous site; the
    # recombination points follow on the same stream
    api.rng_draw(st, len(het))
    b.sample_recombs([st]).sync()
    pos, node, time, draws = b.recombs()
    assert np.array_equal(pos, ref["recomb_pos"])
    assert np.array_equal(node, ref["recomb_node"])
    assert np.array_equal(time, ref["recomb_time"])
    s2 = st.copy()
    api.rng_draw(s2, draws)
    assert int(api.rng_draw(s2, 1)[0]) == int(ref["next_rand"][0])
    b.close()


def test_bad_phase_rows_are_rejected():
    d = problem(8, 500, 20, False, (7, 7), 85)
    with pytest.raises(ValueError):
        emul_lib.Emul(d)
    d = problem(8, 500, 20, True, (8, 1), 85)       # internal: rows 0..7 only
    with pytest.raises(ValueError):
        emul_lib.Emul(d)
