"""Synthetic thread-sampling problems (the role `arg-sim` + a partial ARG play
for the reference).  Wraps the native generator csrc/dsmc_sim.cpp.

A *problem* is a dict of flat arrays -- the same content as ``awb_problem`` in
include/argweaver_b200.h:

  ntimes, times[T], popsizes[T], rho, mu          model (rho, mu per compressed site)
  seqs[nseqs][n] uint8, seqids[nleaves], new_chrom
  internal, minage, start_coord
  ptrees[B][V], ages[B][V], sprs[B][4], mappings[B][V], blocklens[B],
  subtree_roots[B]
"""

import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def _sim_lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_sim())
        _lib.awb_sim_new.restype = C.c_void_p
        _lib.awb_sim_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_double, C.c_double,
                                     C.c_uint64]
        _lib.awb_sim_ntrees.argtypes = [C.c_void_p]
        _lib.awb_sim_nnodes.argtypes = [C.c_void_p]
        _lib.awb_sim_copy.argtypes = [C.c_void_p] * 6
        _lib.awb_sim_free.argtypes = [C.c_void_p]
    return _lib


def get_time_points(ntimes=20, maxtime=200e3, delta=0.01):
    """Log-spaced time grid (reference model.h:21-32)."""
    i = np.arange(ntimes, dtype=np.float64)
    return (np.exp(i / (ntimes - 1) * np.log(1.0 + delta * maxtime)) - 1) / delta


def make_mappings(ptrees, sprs):
    """mapping[b][old] -> new: identity except the broken node (-1).
    Reference local_tree.h:767-776 (make_node_mapping)."""
    B, V = ptrees.shape
    m = np.tile(np.arange(V, dtype=np.int32), (B, 1))
    m[0, :] = -2
    if B > 1:
        rn = sprs[1:, 0]
        broken = ptrees[np.arange(B - 1), rn]
        m[np.arange(1, B), broken] = -1
    return m


def simulate_arg(nleaves, nsites, ntimes=20, maxtime=200e3, popsize=1e4,
                 rho=1.6e-8, mu=1.8e-8, compress=10, seed=1, times=None,
                 popsizes=None):
    """Simulate an ARG over `nleaves` sequences plus nleaves+1 sequences.

    Returns (times, popsizes, rho_c, mu_c, ptrees, ages, sprs, blocklens, seqs).
    """
    lib = _sim_lib()
    if times is None:
        times = get_time_points(ntimes, maxtime)
    times = np.ascontiguousarray(times, np.float64)
    ntimes = len(times)
    if popsizes is None:
        popsizes = np.full(ntimes, float(popsize))
    popsizes = np.ascontiguousarray(popsizes, np.float64)
    rho_c = rho * compress
    mu_c = mu * compress
    h = lib.awb_sim_new(nleaves, nsites, ntimes, times.ctypes.data,
                        popsizes.ctypes.data, rho_c, mu_c, seed)
    try:
        B = lib.awb_sim_ntrees(h)
        V = lib.awb_sim_nnodes(h)
        ptrees = np.empty((B, V), np.int32)
        ages = np.empty((B, V), np.int32)
        sprs = np.empty((B, 4), np.int32)
        blocklens = np.empty(B, np.int32)
        seqs = np.empty((nleaves + 1, nsites), np.uint8)
        lib.awb_sim_copy(h, ptrees.ctypes.data, ages.ctypes.data,
                         sprs.ctypes.data, blocklens.ctypes.data,
                         seqs.ctypes.data)
    finally:
        lib.awb_sim_free(h)
    return times, popsizes, rho_c, mu_c, ptrees, ages, sprs, blocklens, seqs


def simulate_problem(k, nsites, ntimes=20, maxtime=200e3, popsize=1e4,
                     rho=1.6e-8, mu=1.8e-8, compress=10, seed=1,
                     internal=False, times=None, popsizes=None):
    """A thread-sampling problem for `k` sequences.

    external: the ARG holds sequences 0..k-2, sequence k-1 is threaded in.
    internal: the same ARG in subtree-maintree form -- every tree gets the new
    leaf (index k-1) and a virtual root of age ntimes+1 whose child[0] is the
    new leaf (reference thread.cpp:1521-1539), i.e. the state the reference is in
    after removing a leaf thread in resample_arg_leaf.
    """
    (times, popsizes, rho_c, mu_c, ptrees, ages, sprs, blocklens,
     seqs) = simulate_arg(k - 1, nsites, ntimes, maxtime, popsize, rho, mu,
                          compress, seed, times, popsizes)
    B, V = ptrees.shape
    nl = k - 1
    ntimes = len(times)
    d = dict(ntimes=np.int32(ntimes), times=times, popsizes=popsizes,
             rho=np.float64(rho_c), mu=np.float64(mu_c), seqs=seqs,
             minage=np.int32(0), start_coord=np.int32(0), blocklens=blocklens)
    if not internal:
        d.update(internal=np.int32(0), new_chrom=np.int32(k - 1),
                 seqids=np.arange(nl, dtype=np.int32), ptrees=ptrees,
                 ages=ages, sprs=sprs,
                 subtree_roots=np.full(B, -1, np.int32))
        d["mappings"] = make_mappings(ptrees, sprs)
        return d

    # subtree-maintree form: leaves 0..k-2 keep their index, new leaf = k-1,
    # internal nodes shift by one, virtual root = last index.
    V2 = V + 2
    vroot = V2 - 1

    def shift(x):
        x = np.asarray(x)
        return np.where(x >= nl, x + 1, x)

    pt2 = np.full((B, V2), -1, np.int32)
    ag2 = np.zeros((B, V2), np.int32)
    old_idx = np.arange(V)
    new_idx = shift(old_idx)
    par = np.where(ptrees >= 0, shift(ptrees), vroot)
    pt2[:, new_idx] = par
    ag2[:, new_idx] = ages
    pt2[:, nl] = vroot
    ag2[:, nl] = 0
    pt2[:, vroot] = -1
    ag2[:, vroot] = ntimes + 1
    sp2 = sprs.copy()
    valid = sprs[:, 0] >= 0
    sp2[valid, 0] = shift(sprs[valid, 0])
    sp2[valid, 2] = shift(sprs[valid, 2])
    d.update(internal=np.int32(1), new_chrom=np.int32(-1),
             seqids=np.arange(k, dtype=np.int32), ptrees=pt2, ages=ag2,
             sprs=sp2, subtree_roots=np.full(B, nl, np.int32))
    d["mappings"] = make_mappings(pt2, sp2)
    return d


def pack_problem(d):
    """The same problem with its alignment given as variant columns only
    (awb_problem.var_pos / var_cols, the form a .sites file holds): every column
    in which the sequences differ, or which is masked ('N').  The other columns
    read 'A' in every row, as make_sequences_from_sites fills them
    (sequences.cpp:323-352); under Jukes-Cantor an invariant column has the same
    likelihood whatever its base."""
    seqs = np.asarray(d["seqs"])
    var = np.nonzero((seqs != seqs[0]).any(axis=0) | (seqs[0] == ord("N")))[0]
    q = {k: v for k, v in d.items() if k != "seqs"}
    q["var_pos"] = var.astype(np.int32)
    q["var_cols"] = np.ascontiguousarray(seqs[:, var].T)
    q["nseqs"] = np.int32(seqs.shape[0])
    q["seqlen"] = np.int32(seqs.shape[1])
    return q


def write_sites(filename, seqs, compress=10, chrom="chr", names=None):
    """Write the variant columns of `seqs` as a reference .sites file
    (sequences.cpp:136-158): compressed site i sits at bp i*compress+1."""
    k, n = seqs.shape
    if names is None:
        names = ["n%d" % i for i in range(k)]
    var = np.nonzero((seqs != seqs[0]).any(axis=0))[0]
    with open(filename, "w") as f:
        f.write("NAMES\t" + "\t".join(names) + "\n")
        f.write("REGION\t%s\t1\t%d\n" % (chrom, n * compress))
        for i in var:
            f.write("%d\t%s\n" % (i * compress + 1,
                                  seqs[:, i].tobytes().decode()))
