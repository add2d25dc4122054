"""Builds and binds tests/host_emul.cpp (CPU emulation of the product's
__host__ __device__ setup workers) -- test infrastructure."""

import ctypes as C
import os
import subprocess

import numpy as np

from argweaver_b200.problem import make_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests")
BUILD = os.path.join(HERE, "_build")
CSRC = os.path.join(ROOT, "argweaver_b200", "csrc")
INC = os.path.join(ROOT, "include")

_lib = None

DTYPES = {
    "st_node": np.int16, "st_time": np.int8, "perm": np.uint16,
    "pslot": np.uint16, "band_j1": np.uint16, "band_len": np.uint8,
    "band_boff": np.int32, "inv_emit": np.float64, "band": np.float64,
    "tmatrix": np.float64, "tmvec": np.float64, "rowstart": np.uint16,
    "pstart": np.uint16, "node_first": np.int16, "node_cnt": np.int16,
    "child0": np.int16, "child1": np.int16, "order": np.int16,
    "root": np.int16, "lineages": np.int32, "treelen": np.float64,
    "tm_minage": np.int32, "sw_start": np.uint16, "sw_cnt": np.uint16,
    "sw_src": np.uint16, "sw_prob": np.float64, "sw_determ": np.int32,
    "sw_determprob": np.float64, "sw_recombrow": np.float64,
    "sw_recoalrow": np.float64, "sw_recombsrc": np.int32,
    "sw_recoalsrc": np.int32, "kind": np.uint8, "fw": np.float64,
    "path": np.int32, "ent_off": np.int64, "band_off": np.int64,
}


def build():
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, "libhost_emul.so")
    srcs = [os.path.join(HERE, "host_emul.cpp")] + [
        os.path.join(CSRC, f) for f in ("awb_common.cuh", "awb_setup.cuh",
                                        "awb_emit.cuh", "awb_layout.h",
                                        "awb_recomb.cuh")]
    srcs.append(os.path.join(INC, "argweaver_b200.h"))
    if (not os.path.exists(so) or
            os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs)):
        subprocess.check_call(
            ["g++", "-O2", "-fPIC", "-shared", "-std=c++11", "-I", CSRC, "-I",
             INC, "-o", so, srcs[0]])
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emul_create.restype = C.c_void_p
        _lib.emul_create.argtypes = [C.c_void_p]
        _lib.emul_error.restype = C.c_char_p
        _lib.emul_error.argtypes = [C.c_void_p]
        _lib.emul_code.argtypes = [C.c_void_p]
        _lib.emul_setup.argtypes = [C.c_void_p]
        _lib.emul_forward.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        _lib.emul_array_bytes.restype = C.c_longlong
        _lib.emul_array_bytes.argtypes = [C.c_void_p, C.c_char_p]
        _lib.emul_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p,
                                  C.c_longlong]
        _lib.emul_layout.argtypes = [C.c_void_p] * 7
        _lib.emul_destroy.argtypes = [C.c_void_p]
        _lib.emul_lin_unsafe.argtypes = [C.c_void_p]
        _lib.emul_lin_max.restype = C.c_double
        _lib.emul_lin_max.argtypes = [C.c_void_p]
        _lib.emul_phase_probs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.emul_sample_recombs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


class Emul(object):
    def __init__(self, d):
        self.p, self.keep = make_problem(d)
        self.h = lib().emul_create(C.byref(self.p))
        if lib().emul_code(self.h) != 0:
            raise ValueError(lib().emul_error(self.h).decode())
        self.B = self.p.ntrees

    def setup(self):
        rc = lib().emul_setup(self.h)
        if rc:
            raise RuntimeError("emul_setup failed with code %d" % rc)
        return self

    def forward(self):
        z = C.c_double()
        lib().emul_forward(self.h, C.byref(z))
        return z.value

    def get(self, name):
        nb = lib().emul_array_bytes(self.h, name.encode())
        if nb < 0:
            raise KeyError(name)
        dt = np.dtype(DTYPES[name])
        out = np.empty(nb // dt.itemsize, dt)
        rc = lib().emul_get(self.h, name.encode(), out.ctypes.data, nb)
        assert rc == 0
        return out

    def sample_recombs(self, path, rng_state, rand_max=2147483647):
        """(pos, node, time, draws) of the recombination sampler run on the host
        over `path` (after setup())."""
        path = np.ascontiguousarray(path, np.int32)
        st = np.ascontiguousarray(rng_state, np.int32)
        cap = len(path)
        pos = np.empty(cap, np.int32)
        node = np.empty(cap, np.int32)
        time = np.empty(cap, np.int32)
        info = np.zeros(2, np.int32)
        lib().emul_sample_recombs(self.h, path.ctypes.data, st.ctypes.data,
                                  int(rand_max), cap, pos.ctypes.data,
                                  node.ctypes.data, time.ctypes.data,
                                  info.ctypes.data)
        n = int(info[0])
        return pos[:n], node[:n], time[:n], int(info[1])

    def phase_probs(self, path):
        path = np.ascontiguousarray(path, np.int32)
        out = np.empty(len(path), np.float64)
        lib().emul_phase_probs(self.h, path.ctypes.data, out.ctypes.data)
        return out

    def lin_unsafe(self):
        return bool(lib().emul_lin_unsafe(self.h))

    def lin_max(self):
        return lib().emul_lin_max(self.h)

    def layout(self):
        B = self.B
        ns = np.empty(B, np.int32)
        ro = np.empty(B + 1, np.int64)
        fo = np.empty(B + 1, np.int64)
        so = np.empty(B + 1, np.int64)
        ms = C.c_int()
        mb = C.c_int()
        lib().emul_layout(self.h, ns.ctypes.data, ro.ctypes.data,
                          fo.ctypes.data, so.ctypes.data, C.byref(ms),
                          C.byref(mb))
        return dict(nstates=ns, row_off=ro, fw_off=fo, sw1_off=so,
                    maxS=ms.value, maxband=mb.value)

    def __del__(self):
        try:
            lib().emul_destroy(self.h)
        except Exception:
            pass
