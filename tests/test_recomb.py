"""Recombination points of a sampled thread: the device sampler
(argweaver_b200/csrc/awb_recomb.cuh) against the reference's
sample_recombinations (recomb.cpp:151-235) run by oracle/_ref/ref_bench right
after its own traceback, on the same libc rand() stream.

CPU part: the generator snapshot against glibc itself, and the sampler's
__host__ __device__ code run on the host over the reference's path.  GPU part:
the whole chain on the device (traceback -> recombination points)."""

import ctypes

import numpy as np
import pytest

import emul_lib
import ref_lib
from argweaver_b200 import api, sim

needs_ref = pytest.mark.skipif(not ref_lib.available(),
                               reason="oracle/_ref/ref_bench not built")


def test_rng_snapshot_is_glibc_rand():
    """awb_libc_rand_snapshot + awb_rng_draw reproduce rand() (TYPE_3 additive
    feedback generator, glibc random_r.c) and leave the stream untouched"""
    libc = ctypes.CDLL("libc.so.6")
    for seed, skip in ((1, 0), (12345, 7), (2**31 + 5, 400)):
        libc.srand(ctypes.c_uint(seed))
        for _ in range(skip):
            libc.rand()
        st = api.libc_rand_snapshot()
        ref = np.array([libc.rand() for _ in range(2000)], np.int32)
        assert np.array_equal(api.rng_draw(st, 2000), ref)
        # ... and the advanced state keeps producing the same stream
        ref2 = np.array([libc.rand() for _ in range(10)], np.int32)
        assert np.array_equal(api.rng_draw(st, 10), ref2)


def state_after(libc_rand, seed, n):
    r = libc_rand(seed, n)
    return r, api.libc_rand_snapshot()


def next_rand_after(st, draws):
    s = st.copy()
    api.rng_draw(s, draws)
    return int(api.rng_draw(s, 1)[0])


CASES = [(8, 3000, 20, False, 41), (8, 3000, 20, True, 42),
         (20, 20000, 20, False, 43), (20, 20000, 20, True, 44),
         (12, 5000, 40, True, 45)]


@needs_ref
@pytest.mark.parametrize("k,n,T,internal,seed", CASES)
def test_host_sampler_matches_reference(k, n, T, internal, seed, libc_rand):
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal)
    ref = ref_lib.run_reference(d, rand_seed=seed, fw_stride=n)
    _, st = state_after(libc_rand, seed, n)
    e = emul_lib.Emul(d).setup()
    pos, node, time, draws = e.sample_recombs(ref["path"], st)
    assert np.array_equal(pos, ref["recomb_pos"])
    assert np.array_equal(node, ref["recomb_node"])
    assert np.array_equal(time, ref["recomb_time"])
    assert len(pos) > 0
    assert next_rand_after(st, draws) == int(ref["next_rand"][0])


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("checkpoint", [False, True])
@pytest.mark.parametrize("k,n,T,internal,seed", CASES + [(50, 200000, 20, True, 46)])
def test_device_sampler_matches_reference(k, n, T, internal, seed, checkpoint,
                                          libc_rand, monkeypatch):
    if checkpoint:
        monkeypatch.setenv("AWB_SEG_DOUBLES", str(max(n * 8, 1 << 16)))
    d = sim.simulate_problem(k, n, ntimes=T, seed=seed, internal=internal)
    ref = ref_lib.run_reference(d, rand_seed=seed, fw_stride=n)
    r, st = state_after(libc_rand, seed, n)
    b = api.Batch([d], checkpoint=checkpoint)
    b.upload().setup().forward().traceback([r]).sample_recombs([st]).sync()
    assert np.array_equal(b.path(), ref["path"])
    pos, node, time, draws = b.recombs()
    assert np.array_equal(pos, ref["recomb_pos"])
    assert np.array_equal(node, ref["recomb_node"])
    assert np.array_equal(time, ref["recomb_time"])
    assert next_rand_after(st, draws) == int(ref["next_rand"][0])
    b.close()


@pytest.mark.gpu
@needs_ref
def test_batch_of_windows_each_with_its_own_stream(libc_rand):
    ds, rs, sts, refs = [], [], [], []
    for i in range(5):
        n = 4000 + 500 * i
        d = sim.simulate_problem(10, n, ntimes=20, seed=60 + i, internal=bool(i & 1))
        r, st = state_after(libc_rand, 70 + i, n)
        ds.append(d); rs.append(r); sts.append(st)
        refs.append(ref_lib.run_reference(d, rand_seed=70 + i, fw_stride=n))
    b = api.Batch(ds)
    b.upload().setup().forward().traceback(rs).sample_recombs(sts).sync()
    for i, ref in enumerate(refs):
        pos, node, time, draws = b.recombs(i)
        assert np.array_equal(pos, ref["recomb_pos"])
        assert np.array_equal(node, ref["recomb_node"])
        assert np.array_equal(time, ref["recomb_time"])
        assert next_rand_after(sts[i], draws) == int(ref["next_rand"][0])
    b.close()
