"""The CPU oracle (oracle/oracle.c) against golden vectors dumped from the
unmodified reference (tests/golden/*.npz, made by tests/golden/make_golden.py).
This is what pins the oracle."""

import numpy as np
import pytest

import oracle_lib as ol
from conftest import golden_files, load_golden
from helpers import assert_close

INT_KEYS = ["nstates", "states", "nbranches", "nrecombs", "ncoals", "tm_minage",
            "sw_determ", "sw_recombsrc", "sw_recoalsrc", "row_off", "fw_off",
            "sw1_off"]
FLT_KEYS = ol.TM_NAMES + ["sw_determprob", "sw_recombrow", "sw_recoalrow",
                          "emit", "fw"]


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1])
def test_oracle_matches_reference(path):
    g = load_golden(path)
    o = ol.run_oracle(g, g["rand_ints"], int(g["rand_max"][0]))
    for k in INT_KEYS:
        assert np.array_equal(o[k], g[k]), k
    for k in FLT_KEYS:
        assert_close(o[k], g[k], k, rtol=1e-12)
    assert np.array_equal(o["path"], g["path"])
    assert o["rand_used"] == int(g["rand_used"][0]) == len(g["path"])
    assert o["first_bad_site"] == -1


def test_golden_covers_edge_cases():
    """zero-state blocks, internal and external mode, several ntimes"""
    modes, zero, ntimes = set(), False, set()
    for path in golden_files():
        g = load_golden(path)
        modes.add(int(g["internal"][0]))
        ntimes.add(int(g["ntimes"][0]))
        zero |= bool((g["nstates"] == 0).any())
    assert modes == {0, 1} and zero and len(ntimes) >= 3


def test_mapping_invariant():
    """Across an SPR only the broken node (parent of the recombining branch)
    maps to -1 and the map is injective; it is NOT always the identity
    elsewhere (ARG surgery renames nodes), so the product carries the
    caller's mapping arrays to the device."""
    nonident = 0
    for path in golden_files():
        g = load_golden(path)
        B, V = g["ptrees"].shape
        for b in range(1, B):
            m = g["mappings"][b]
            broken = g["ptrees"][b - 1][g["sprs"][b][0]]
            assert m[broken] == -1
            rest = np.delete(m, broken)
            assert (rest >= 0).all() and len(set(rest.tolist())) == V - 1
            nonident += int((rest != np.delete(np.arange(V), broken)).any())
    assert nonident > 0
