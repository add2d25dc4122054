"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- test infrastructure.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.  The product package never does.
"""

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_c_int_p = C.POINTER(C.c_int)
_c_dbl_p = C.POINTER(C.c_double)
_c_i64_p = C.POINTER(C.c_int64)
_c_u8_p = C.POINTER(C.c_ubyte)


class OrcProblem(C.Structure):
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
    ]


class OrcResult(C.Structure):
    _fields_ = [
        ("nstates", _c_int_p), ("state_off", _c_i64_p), ("states", _c_int_p),
        ("row_off", _c_i64_p), ("fw_off", _c_i64_p), ("sw1_off", _c_i64_p),
        ("nbranches", _c_int_p), ("nrecombs", _c_int_p), ("ncoals", _c_int_p),
        ("tm", _c_dbl_p * 9), ("tm_minage", _c_int_p),
        ("sw_determ", _c_int_p), ("sw_determprob", _c_dbl_p),
        ("sw_recombrow", _c_dbl_p), ("sw_recoalrow", _c_dbl_p),
        ("sw_recombsrc", _c_int_p), ("sw_recoalsrc", _c_int_p),
        ("emit", _c_dbl_p), ("fw", _c_dbl_p), ("path", _c_int_p),
        ("logZ", C.c_double), ("first_bad_site", C.c_int),
    ]


TM_NAMES = ["tm_D", "tm_E", "tm_lnB", "tm_lnE2", "tm_lnNegG1", "tm_G2",
            "tm_G3", "tm_lnG4", "tm_norecombs"]

_lib = None


def build():
    """Compile oracle/liboracle.so (gcc) if missing or stale."""
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "oracle.c")
    hdr = os.path.join(ORACLE_DIR, "oracle.h")
    if (not os.path.exists(so) or
            os.path.getmtime(so) < max(os.path.getmtime(src),
                                       os.path.getmtime(hdr))):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_result_new.restype = C.POINTER(OrcResult)
        _lib.orc_result_new.argtypes = [C.POINTER(OrcProblem)]
        _lib.orc_result_free.argtypes = [C.POINTER(OrcResult)]
        _lib.orc_setup.argtypes = [C.POINTER(OrcProblem), C.POINTER(OrcResult)]
        _lib.orc_forward.argtypes = [C.POINTER(OrcProblem),
                                     C.POINTER(OrcResult), _c_dbl_p]
        _lib.orc_traceback.restype = C.c_int
        _lib.orc_traceback.argtypes = [C.POINTER(OrcProblem),
                                       C.POINTER(OrcResult), _c_int_p, C.c_int,
                                       C.c_int]
        _lib.orc_get_time.restype = C.c_double
    return _lib


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


def normalize_problem(d):
    """Return a dict of contiguous, correctly-typed input arrays."""
    q = {}
    q["ntimes"] = int(np.ravel(d["ntimes"])[0])
    q["times"] = np.ascontiguousarray(d["times"], np.float64)
    q["popsizes"] = np.ascontiguousarray(d["popsizes"], np.float64)
    q["rho"] = float(np.ravel(d["rho"])[0])
    q["mu"] = float(np.ravel(d["mu"])[0])
    q["seqs"] = np.ascontiguousarray(d["seqs"], np.uint8)
    q["seqids"] = np.ascontiguousarray(d["seqids"], np.int32)
    q["new_chrom"] = int(np.ravel(d["new_chrom"])[0])
    q["internal"] = int(np.ravel(d["internal"])[0])
    q["minage"] = int(np.ravel(d.get("minage", 0))[0])
    q["start_coord"] = int(np.ravel(d.get("start_coord", 0))[0])
    q["ptrees"] = np.ascontiguousarray(d["ptrees"], np.int32)
    q["ages"] = np.ascontiguousarray(d["ages"], np.int32)
    q["sprs"] = np.ascontiguousarray(d["sprs"], np.int32)
    q["mappings"] = np.ascontiguousarray(d["mappings"], np.int32)
    q["blocklens"] = np.ascontiguousarray(d["blocklens"], np.int32)
    B, V = q["ptrees"].shape
    if "subtree_roots" in d:
        sr = np.ascontiguousarray(d["subtree_roots"], np.int32)
    elif q["internal"] and "child0" in d:
        roots = np.asarray(d["roots"])
        sr = np.ascontiguousarray(
            np.asarray(d["child0"])[np.arange(B), roots], np.int32)
    else:
        sr = np.full(B, -1, np.int32)
    q["subtree_roots"] = sr
    return q


class OracleRun(object):
    """Runs the oracle on a problem dict; exposes outputs as numpy arrays."""

    def __init__(self, d):
        self.q = q = normalize_problem(d)
        L = lib()
        p = OrcProblem()
        p.ntimes = q["ntimes"]
        p.times = _ptr(q["times"], _c_dbl_p)
        p.popsizes = _ptr(q["popsizes"], _c_dbl_p)
        p.rho = q["rho"]
        p.mu = q["mu"]
        p.nseqs, p.seqlen = q["seqs"].shape
        p.seqs = _ptr(q["seqs"], _c_u8_p)
        p.nleaves = len(q["seqids"])
        p.seqids = _ptr(q["seqids"], _c_int_p)
        p.new_chrom = q["new_chrom"]
        p.internal = q["internal"]
        p.minage = q["minage"]
        p.ntrees, p.nnodes = q["ptrees"].shape
        p.start_coord = q["start_coord"]
        p.ptrees = _ptr(q["ptrees"], _c_int_p)
        p.ages = _ptr(q["ages"], _c_int_p)
        p.sprs = _ptr(q["sprs"], _c_int_p)
        p.mappings = _ptr(q["mappings"], _c_int_p)
        p.blocklens = _ptr(q["blocklens"], _c_int_p)
        p.subtree_roots = _ptr(q["subtree_roots"], _c_int_p)
        self.p = p
        self.r = L.orc_result_new(C.byref(p))
        self.B = p.ntrees
        self.T = p.ntimes
        self.n = int(q["blocklens"].sum())

    def __del__(self):
        try:
            lib().orc_result_free(self.r)
        except Exception:
            pass

    def setup(self):
        lib().orc_setup(C.byref(self.p), self.r)
        return self

    def forward(self, prior=None):
        pr = None
        if prior is not None:
            prior = np.ascontiguousarray(prior, np.float64)
            pr = _ptr(prior, _c_dbl_p)
        lib().orc_forward(C.byref(self.p), self.r, pr)
        return self

    def traceback(self, rand_ints, rand_max=2147483647, last_state=-1):
        ri = np.ascontiguousarray(rand_ints, np.int32)
        return lib().orc_traceback(C.byref(self.p), self.r,
                                   _ptr(ri, _c_int_p), int(rand_max),
                                   int(last_state))

    # ---- outputs
    def _arr(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

    def outputs(self):
        r = self.r.contents
        B, T = self.B, self.T
        o = {}
        o["nstates"] = self._arr(r.nstates, B, np.int32)
        o["state_off"] = self._arr(r.state_off, B + 1, np.int64)
        o["row_off"] = self._arr(r.row_off, B + 1, np.int64)
        o["fw_off"] = self._arr(r.fw_off, B + 1, np.int64)
        o["sw1_off"] = self._arr(r.sw1_off, B + 1, np.int64)
        ns = int(o["state_off"][-1])
        o["states"] = self._arr(r.states, 2 * ns, np.int32).reshape(ns, 2)
        for nm, ptr in (("nbranches", r.nbranches), ("nrecombs", r.nrecombs),
                        ("ncoals", r.ncoals)):
            o[nm] = self._arr(ptr, B * T, np.int32).reshape(B, T)
        for k, nm in enumerate(TM_NAMES):
            o[nm] = self._arr(r.tm[k], B * T, np.float64).reshape(B, T)
        o["tm_minage"] = self._arr(r.tm_minage, B, np.int32)
        n1 = int(o["sw1_off"][-1])
        o["sw_determ"] = self._arr(r.sw_determ, n1, np.int32)
        o["sw_determprob"] = self._arr(r.sw_determprob, n1, np.float64)
        nr = int(o["row_off"][-1])
        o["sw_recombrow"] = self._arr(r.sw_recombrow, nr, np.float64)
        o["sw_recoalrow"] = self._arr(r.sw_recoalrow, nr, np.float64)
        o["sw_recombsrc"] = self._arr(r.sw_recombsrc, B, np.int32)
        o["sw_recoalsrc"] = self._arr(r.sw_recoalsrc, B, np.int32)
        nf = int(o["fw_off"][-1])
        o["emit"] = self._arr(r.emit, nf, np.float64)
        o["fw"] = self._arr(r.fw, nf, np.float64)
        o["path"] = self._arr(r.path, self.n, np.int32)
        o["logZ"] = float(r.logZ)
        o["first_bad_site"] = int(r.first_bad_site)
        return o


def run_oracle(d, rand_ints=None, rand_max=2147483647):
    """setup + forward (+ traceback if rand_ints given); returns outputs dict."""
    run = OracleRun(d).setup().forward()
    if rand_ints is not None:
        used = run.traceback(rand_ints, rand_max)
        o = run.outputs()
        o["rand_used"] = used
        return o
    return run.outputs()
