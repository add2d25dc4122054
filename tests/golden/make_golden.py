"""Regenerates tests/golden/*.npz from the UNMODIFIED reference.

Needs oracle/_ref/ref_dump (``make -C oracle ref``; only possible where
/root/reference exists).  Each fixture holds the flattened inputs of one
thread-sampling call and everything the reference computed for it (see
oracle/ref_dump.cpp).  Run from the repository root:

    python tests/golden/make_golden.py
"""

import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from argweaver_b200 import sim  # noqa: E402
from argweaver_b200.flatfile import read_awf  # noqa: E402

REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")
SIM1 = "/root/reference/examples/sim1/sim1.sites"
OUT = os.path.join(ROOT, "tests", "golden")

# name, sites source, ref_dump options
CASES = [
    ("sim1_external", SIM1, ["--region", "1-6000", "--mode", "external"]),
    ("sim1_internal_leaf", SIM1, ["--region", "1-6000", "--mode", "internal-leaf"]),
    ("sim1_internal_uniform", SIM1,
     ["--region", "1-6000", "--mode", "internal-uniform"]),
    ("sim1_internal_uniform2", SIM1,
     ["--region", "20001-26000", "--mode", "internal-uniform", "--seed", "5"]),
    ("gen_k12_T30_external", ("gen", dict(k=12, nsites=400, ntimes=30, seed=11)),
     ["--ntimes", "30", "--mode", "external", "--seed", "3"]),
    ("gen_k6_T10_internal", ("gen", dict(k=6, nsites=500, ntimes=10, seed=12)),
     ["--ntimes", "10", "--mode", "internal-uniform", "--seed", "4"]),
    ("gen_k16_refined", ("gen", dict(k=16, nsites=300, ntimes=20, seed=13)),
     ["--mode", "internal-leaf", "--seed", "2", "--refine", "8"]),
]


def main():
    if not os.path.exists(REF_DUMP):
        raise SystemExit("build oracle/_ref first: make -C oracle ref")
    tmp = tempfile.mkdtemp()
    for name, src, opts in CASES:
        if isinstance(src, tuple):
            kw = src[1]
            (_, _, _, _, _, _, _, _, seqs) = sim.simulate_arg(
                kw["k"] - 1, kw["nsites"], ntimes=kw["ntimes"], seed=kw["seed"])
            # all k rows are real data: use the generator's k rows as k sequences
            sites = os.path.join(tmp, name + ".sites")
            sim.write_sites(sites, seqs, compress=10)
        else:
            sites = src
        awf = os.path.join(tmp, name + ".awf")
        subprocess.check_call([REF_DUMP, "--sites", sites, "--out", awf] + opts)
        d = read_awf(awf)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "sites", int(d["blocklens"].sum()), "trees", len(d["blocklens"]),
              "nstates", d["nstates"].min(), d["nstates"].max())


if __name__ == "__main__":
    main()
