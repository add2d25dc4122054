"""calc_arg_likelihood / calc_arg_prior (total_prob.cpp:19-42, :262-299) on the
device against the UNMODIFIED reference library, through the same ctypes calls
(arghmm_likelihood, arghmm_prior_prob, arghmm_joint_prob; total_prob.cpp:316-377)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from argweaver_b200 import sim
from test_gpu_compat import bind, new_trees, rows_char, dptr

pytestmark = pytest.mark.gpu

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libargweaver.so")
RTOL = 1e-9


def bind_tp(lib):
    bind(lib)
    lib.arghmm_likelihood.restype = C.c_double
    lib.arghmm_likelihood.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int,
                                      C.c_double, C.c_void_p, C.c_int, C.c_int]
    lib.arghmm_prior_prob.restype = C.c_double
    lib.arghmm_prior_prob.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int,
                                      C.POINTER(C.c_double), C.c_double]
    lib.arghmm_joint_prob.restype = C.c_double
    lib.arghmm_joint_prob.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int,
                                      C.POINTER(C.c_double), C.c_double, C.c_double,
                                      C.c_void_p, C.c_int, C.c_int]
    return lib


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libargweaver.so not built (make -C oracle ref)")
    from argweaver_b200 import api
    return bind_tp(C.CDLL(api.lib()._name)), bind_tp(C.CDLL(REF_SO))


def full_arg(k, n, T, seed, mask=None):
    (times, pops, rho, mu, ptrees, ages, sprs, blocklens, seqs) = sim.simulate_arg(
        k, n, T, seed=seed)
    seqs = seqs[:k].copy()
    if mask:
        for lo, hi in mask:
            seqs[:, lo:hi] = ord("N")
    return dict(times=times, popsizes=pops, rho=rho, mu=mu, ptrees=ptrees, ages=ages,
                sprs=sprs, blocklens=blocklens, seqs=seqs, start_coord=0)


def both(libs, a):
    out = []
    for lib in libs:
        trees = new_trees(lib, a)
        sp, _keep = rows_char(a["seqs"])
        k, n = a["seqs"].shape
        T = len(a["times"])
        lik = lib.arghmm_likelihood(trees, dptr(a["times"]), T, float(a["mu"]), sp, k, n)
        pri = lib.arghmm_prior_prob(trees, dptr(a["times"]), T, dptr(a["popsizes"]),
                                    float(a["rho"]))
        joint = lib.arghmm_joint_prob(trees, dptr(a["times"]), T, dptr(a["popsizes"]),
                                      float(a["mu"]), float(a["rho"]), sp, k, n)
        lib.delete_local_trees(trees)
        out.append((lik, pri, joint))
    return out


@pytest.mark.parametrize("k,n,T,seed", [(2, 400, 20, 1), (3, 800, 20, 2), (8, 5000, 20, 3),
                                        (20, 20000, 20, 4), (12, 3000, 40, 5),
                                        (50, 100000, 20, 6), (100, 20000, 40, 7)])
def test_likelihood_and_prior_match_the_reference(libs, k, n, T, seed):
    a = full_arg(k, n, T, seed)
    mine, ref = both(libs, a)
    for x, y, name in zip(mine, ref, ("likelihood", "prior", "joint")):
        assert abs(x - y) <= RTOL * abs(y), (name, x, y)


def test_masked_columns_and_the_first_invariant_site_rule(libs):
    """every invariant site of a block takes the likelihood of the block's first
    invariant site (emit.cpp:431-442) -- 1.0 when that one is all 'N'"""
    a = full_arg(6, 3000, 20, 11, mask=[(0, 40), (700, 760), (2990, 3000)])
    # a block that starts inside a masked stretch
    mine, ref = both(libs, a)
    for x, y in zip(mine, ref):
        assert abs(x - y) <= RTOL * abs(y), (x, y)
