"""ctypes mirror of ``awb_problem`` (include/argweaver_b200.h) and helpers to
build it from a dict of numpy arrays (see argweaver_b200.sim for the keys)."""

import ctypes as C

import numpy as np

_c_int_p = C.POINTER(C.c_int)
_c_dbl_p = C.POINTER(C.c_double)
_c_u8_p = C.POINTER(C.c_ubyte)


class AwbProblem(C.Structure):
    _fields_ = [
        ("ntimes", C.c_int), ("times", _c_dbl_p), ("popsizes", _c_dbl_p),
        ("rho", C.c_double), ("mu", C.c_double),
        ("nseqs", C.c_int), ("seqlen", C.c_int), ("seqs", _c_u8_p),
        ("nleaves", C.c_int), ("seqids", _c_int_p), ("new_chrom", C.c_int),
        ("internal", C.c_int), ("minage", C.c_int),
        ("ntrees", C.c_int), ("nnodes", C.c_int), ("start_coord", C.c_int),
        ("ptrees", _c_int_p), ("ages", _c_int_p), ("sprs", _c_int_p),
        ("mappings", _c_int_p), ("blocklens", _c_int_p),
        ("subtree_roots", _c_int_p),
        ("nvar", C.c_int), ("var_pos", _c_int_p), ("var_cols", _c_u8_p),
        ("default_char", C.c_ubyte), ("infsites_penalty", C.c_double),
        ("unphased", C.c_int), ("phase_row1", C.c_int), ("phase_row2", C.c_int),
    ]


def _scalar(d, key, default=None):
    if key not in d:
        if default is None:
            raise KeyError(key)
        return default
    return np.ravel(d[key])[0]


def normalize(d):
    """Contiguous, correctly typed copies of the input arrays of a problem."""
    q = {}
    q["ntimes"] = int(_scalar(d, "ntimes"))
    q["times"] = np.ascontiguousarray(d["times"], np.float64)
    q["popsizes"] = np.ascontiguousarray(d["popsizes"], np.float64)
    q["rho"] = float(_scalar(d, "rho"))
    q["mu"] = float(_scalar(d, "mu"))
    # dense rows, or ("var_pos", "var_cols", "nseqs", "seqlen"): the variant
    # columns only (awb_problem.var_cols)
    q["var_pos"] = q["var_cols"] = None
    if d.get("var_cols") is not None:
        q["var_pos"] = np.ascontiguousarray(d["var_pos"], np.int32)
        q["var_cols"] = np.ascontiguousarray(d["var_cols"], np.uint8)
        q["nseqs"] = int(_scalar(d, "nseqs"))
        q["seqlen"] = int(_scalar(d, "seqlen"))
        if q["var_cols"].shape != (len(q["var_pos"]), q["nseqs"]):
            raise ValueError("var_cols must be [nvar][nseqs]")
        q["seqs"] = None
    else:
        q["seqs"] = np.ascontiguousarray(d["seqs"], np.uint8)
    q["infsites_penalty"] = float(_scalar(d, "infsites_penalty", 0.0))
    # unphased individual: its two rows in the leaf order, or None
    q["phase_rows"] = None
    if d.get("phase_rows") is not None:
        pr = np.asarray(d["phase_rows"]).reshape(-1)
        q["phase_rows"] = (int(pr[0]), int(pr[1]))
    q["seqids"] = np.ascontiguousarray(d["seqids"], np.int32)
    q["new_chrom"] = int(_scalar(d, "new_chrom"))
    q["internal"] = int(_scalar(d, "internal"))
    q["minage"] = int(_scalar(d, "minage", 0))
    q["start_coord"] = int(_scalar(d, "start_coord", 0))
    q["ptrees"] = np.ascontiguousarray(d["ptrees"], np.int32)
    q["ages"] = np.ascontiguousarray(d["ages"], np.int32)
    q["sprs"] = np.ascontiguousarray(d["sprs"], np.int32)
    q["blocklens"] = np.ascontiguousarray(d["blocklens"], np.int32)
    if q["ptrees"].ndim != 2:
        raise ValueError("ptrees must be [ntrees][nnodes]")
    B, V = q["ptrees"].shape
    q["mappings"] = (np.ascontiguousarray(d["mappings"], np.int32)
                     if "mappings" in d else None)
    # shapes are checked here: past this point only raw pointers reach C
    T = q["ntimes"]
    for key, shape in (("times", (T,)), ("popsizes", (T,)), ("ages", (B, V)),
                       ("sprs", (B, 4)), ("blocklens", (B,)),
                       ("mappings", (B, V))):
        if q[key] is not None and q[key].shape != shape:
            raise ValueError("%s has shape %s, expected %s"
                             % (key, q[key].shape, shape))
    if q["seqs"] is not None and q["seqs"].ndim != 2:
        raise ValueError("seqs must be [nseqs][seqlen]")
    if q["seqids"].ndim != 1:
        raise ValueError("seqids must be one-dimensional")
    if "subtree_roots" in d:
        q["subtree_roots"] = np.ascontiguousarray(d["subtree_roots"], np.int32)
    elif q["internal"] and "child0" in d:
        roots = np.asarray(d["roots"])
        q["subtree_roots"] = np.ascontiguousarray(
            np.asarray(d["child0"])[np.arange(B), roots], np.int32)
    else:
        q["subtree_roots"] = None
    if q["subtree_roots"] is not None and q["subtree_roots"].shape != (B,):
        raise ValueError("subtree_roots has shape %s, expected %s"
                         % (q["subtree_roots"].shape, (B,)))
    return q


def make_problem(d):
    """Return (AwbProblem, keepalive) for a problem dict."""
    q = normalize(d)
    p = AwbProblem()
    p.ntimes = q["ntimes"]
    p.times = q["times"].ctypes.data_as(_c_dbl_p)
    p.popsizes = q["popsizes"].ctypes.data_as(_c_dbl_p)
    p.rho = q["rho"]
    p.mu = q["mu"]
    if q["seqs"] is not None:
        p.nseqs, p.seqlen = q["seqs"].shape
        p.seqs = q["seqs"].ctypes.data_as(_c_u8_p)
    else:
        p.nseqs, p.seqlen = q["nseqs"], q["seqlen"]
        p.seqs = None
        p.nvar = len(q["var_pos"])
        p.var_pos = q["var_pos"].ctypes.data_as(_c_int_p)
        p.var_cols = q["var_cols"].ctypes.data_as(_c_u8_p)
    p.infsites_penalty = q["infsites_penalty"]
    if q["phase_rows"] is not None:
        p.unphased = 1
        p.phase_row1, p.phase_row2 = q["phase_rows"]
    p.nleaves = len(q["seqids"])
    p.seqids = q["seqids"].ctypes.data_as(_c_int_p)
    p.new_chrom = q["new_chrom"]
    p.internal = q["internal"]
    p.minage = q["minage"]
    p.ntrees, p.nnodes = q["ptrees"].shape
    p.start_coord = q["start_coord"]
    p.ptrees = q["ptrees"].ctypes.data_as(_c_int_p)
    p.ages = q["ages"].ctypes.data_as(_c_int_p)
    p.sprs = q["sprs"].ctypes.data_as(_c_int_p)
    p.mappings = (q["mappings"].ctypes.data_as(_c_int_p)
                  if q["mappings"] is not None else None)
    p.blocklens = q["blocklens"].ctypes.data_as(_c_int_p)
    p.subtree_roots = (q["subtree_roots"].ctypes.data_as(_c_int_p)
                       if q["subtree_roots"] is not None else None)
    return p, q
