"""Runs the UNMODIFIED reference (oracle/_ref/ref_bench: its own
arghmm_forward_alg + stochastic_traceback, sample_thread.cpp:394-460,522-569)
on a problem dict and returns its forward rows and sampled path -- test
infrastructure.  The binary is built by oracle/Makefile where /root/reference
exists and travels to the GPU box prebuilt."""

import os
import re
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench")


def available():
    return os.path.exists(REF_BENCH)


def problem_for_file(d):
    """The arrays ref_bench reads (oracle/ref_bench.cpp load())."""
    keys = ("ntimes", "times", "popsizes", "rho", "mu", "seqs", "seqids",
            "new_chrom", "internal", "start_coord", "ptrees", "ages", "sprs",
            "blocklens", "mappings", "subtree_roots")
    q = {}
    for k in keys:
        if k in d and d[k] is not None:
            v = np.asarray(d[k])
            if v.dtype == np.int64 and k not in ():
                v = v.astype(np.int32)
            q[k] = v
    for k in ("ntimes", "new_chrom", "internal", "start_coord"):
        if k in q:
            q[k] = np.asarray(q[k], np.int32).reshape(-1)[:1]
    for k in ("rho", "mu"):
        q[k] = np.asarray(q[k], np.float64).reshape(-1)[:1]
    if d.get("phase_rows") is not None:
        q["phase_rows"] = np.asarray(d["phase_rows"], np.int32).reshape(-1)[:2]
    if "infsites_penalty" in d:
        q["infsites_penalty"] = np.asarray(d["infsites_penalty"],
                                           np.float64).reshape(-1)[:1]
    q["seqs"] = np.ascontiguousarray(q["seqs"], np.uint8)
    return q


def run_reference(d, rand_seed, fw_stride=1):
    """Returns dict(nstates, fw_sites, fw (flat rows of fw_sites), path,
    forward_s, trace_s, states_sites).  The traceback uses libc rand() after
    srand(rand_seed): the same draws as conftest.libc_rand(rand_seed, n)."""
    from argweaver_b200.flatfile import read_awf, write_awf
    with tempfile.TemporaryDirectory() as tmp:
        fin = os.path.join(tmp, "p.awf")
        fout = os.path.join(tmp, "r.awf")
        write_awf(fin, problem_for_file(d))
        out = subprocess.run([REF_BENCH, "--in", fin, "--out", fout,
                              "--rand-seed", str(int(rand_seed)),
                              "--fw-stride", str(int(fw_stride))],
                             capture_output=True, text=True)
        if out.returncode != 0:
            raise RuntimeError("ref_bench failed: " + out.stderr[-2000:])
        r = read_awf(fout)
    kv = dict(re.findall(r"(\w+)=([-0-9.e+]+)", out.stdout))
    r.update({k: float(v) for k, v in kv.items()})
    return r


def rows_of(fw_flat, fw_off, nstates, blocklens, sites):
    """Rows `sites` of a flat forward table laid out block by block
    (fw_off[b] + (i - start_b) * max(S_b, 1)), concatenated."""
    bs = np.concatenate([[0], np.cumsum(blocklens)]).astype(np.int64)
    blk = np.searchsorted(bs, sites, side="right") - 1
    s1 = np.maximum(np.asarray(nstates, np.int64), 1)
    out = []
    for i, b in zip(sites, blk):
        lo = fw_off[b] + (i - bs[b]) * s1[b]
        out.append(fw_flat[lo:lo + s1[b]])
    return np.concatenate(out) if out else np.zeros(0)
