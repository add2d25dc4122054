"""Python binding of the C ABI (include/argweaver_b200.h) via ctypes.

The host-side names mirror the reference's Python/ctypes layer
(argweaver/argweaverc.py): ``forward_algorithm`` ~ ``argweaver_forward_algorithm``
(:529), ``sample_thread`` ~ the path of ``arghmm_sample_posterior`` (:930 in
sample_thread.cpp).  All compute happens in libargweaver_b200.so on the GPU; the
library has no CPU fallback and raises :class:`AwbError` when CUDA is missing.
"""

import ctypes as C
import os

import numpy as np

from . import build as _build
from .problem import AwbProblem, make_problem

KEEP_DEBUG = 1
CHECKPOINT = 2      # AWB_CHECKPOINT: segment-wise forward table (see the header)
RAND_MAX = 2147483647
RNG_WORDS = 34      # AWB_RNG_WORDS

_lib = None


class AwbError(RuntimeError):
    pass


def lib():
    """Load (building if needed) libargweaver_b200.so."""
    global _lib
    if _lib is None:
        # AWB_LIB: load a specific build (kernel experiments); default = in-tree
        L = C.CDLL(os.environ.get("AWB_LIB") or _build.build_cuda())
        L.awb_last_error.restype = C.c_char_p
        L.awb_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.awb_ctx_destroy.argtypes = [C.c_void_p]
        L.awb_ctx_record.argtypes = [C.c_void_p, C.c_int]
        L.awb_ctx_elapsed_ms.argtypes = [C.c_void_p, C.c_int, C.c_int,
                                         C.POINTER(C.c_float)]
        L.awb_ctx_sync.argtypes = [C.c_void_p]
        L.awb_batch_upload_rand.argtypes = [C.c_void_p, C.c_void_p]
        L.awb_batch_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                       C.POINTER(C.c_void_p)]
        L.awb_batch_destroy.argtypes = [C.c_void_p]
        L.awb_batch_upload.argtypes = [C.c_void_p]
        L.awb_batch_h2d_bytes.restype = C.c_int64
        L.awb_batch_h2d_bytes.argtypes = [C.c_void_p]
        L.awb_batch_setup.argtypes = [C.c_void_p]
        L.awb_batch_forward.argtypes = [C.c_void_p, C.c_void_p]
        L.awb_batch_traceback.argtypes = [C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p]
        L.awb_batch_sync.argtypes = [C.c_void_p]
        L.awb_batch_timings.argtypes = [C.c_void_p] + [C.POINTER(C.c_float)] * 3
        L.awb_batch_states_sites.restype = C.c_double
        L.awb_batch_states_sites.argtypes = [C.c_void_p, C.c_int]
        L.awb_batch_fw_doubles.restype = C.c_int64
        L.awb_batch_fw_doubles.argtypes = [C.c_void_p, C.c_int]
        L.awb_batch_nsites.argtypes = [C.c_void_p, C.c_int]
        L.awb_batch_kernel_launches.argtypes = [C.c_void_p]
        L.awb_batch_segments.argtypes = [C.c_void_p]
        L.awb_batch_resident_segments.argtypes = [C.c_void_p]
        L.awb_batch_forward_kernel.argtypes = [C.c_void_p]
        L.awb_batch_get_path.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.awb_batch_get_logz.argtypes = [C.c_void_p, C.c_int,
                                         C.POINTER(C.c_double)]
        L.awb_batch_get_status.argtypes = [C.c_void_p, C.c_int,
                                           C.POINTER(C.c_int)]
        L.awb_batch_get_fw.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.awb_batch_get_nstates.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.awb_batch_get_layout.argtypes = [C.c_void_p, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        L.awb_batch_get_debug.argtypes = [C.c_void_p, C.c_int, C.c_char_p,
                                          C.c_void_p, C.c_int64]
        L.awb_batch_debug_bytes.restype = C.c_int64
        L.awb_batch_debug_bytes.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        if not hasattr(L, "awb_sites_read"):
            _lib = L            # an older build loaded through AWB_LIB (A/B probes)
            return _lib
        L.awb_sites_read.argtypes = [C.c_char_p, C.c_int, C.c_int,
                                     C.POINTER(C.c_void_p)]
        L.awb_sites_from_columns.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_void_p)]
        L.awb_sites_free.argtypes = [C.c_void_p]
        for nm in ("nseqs", "ncols", "start", "end", "mapping_size"):
            getattr(L, "awb_sites_" + nm).argtypes = [C.c_void_p]
        L.awb_sites_name.restype = C.c_char_p
        L.awb_sites_name.argtypes = [C.c_void_p, C.c_int]
        L.awb_sites_positions.restype = C.POINTER(C.c_int)
        L.awb_sites_positions.argtypes = [C.c_void_p]
        L.awb_sites_columns.restype = C.POINTER(C.c_ubyte)
        L.awb_sites_columns.argtypes = [C.c_void_p]
        L.awb_sites_mapping.restype = C.POINTER(C.c_int)
        L.awb_sites_mapping.argtypes = [C.c_void_p]
        L.awb_sites_compress.argtypes = [C.c_void_p, C.c_int]
        L.awb_sites_to_sequences.argtypes = [C.c_void_p, C.c_void_p, C.c_ubyte]
        L.awb_batch_phase_probs.argtypes = [C.c_void_p]
        L.awb_batch_get_phase_probs.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.awb_batch_kernel_times.argtypes = [C.c_void_p, C.c_int]
        L.awb_batch_get_kernel_times.argtypes = [C.c_void_p, C.c_void_p,
                                                 C.POINTER(C.c_double),
                                                 C.POINTER(C.c_int)]
        L.awb_libc_rand_snapshot.argtypes = [C.c_void_p]
        L.awb_libc_rand_advance.argtypes = [C.c_longlong]
        L.awb_libc_rand_advance.restype = None
        L.awb_rng_draw.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.awb_batch_sample_recombs.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.awb_batch_get_recomb_count.argtypes = [C.c_void_p, C.c_int,
                                                 C.POINTER(C.c_int),
                                                 C.POINTER(C.c_int)]
        L.awb_batch_get_recombs.argtypes = [C.c_void_p, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        L.awb_arg_likelihood.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.awb_arg_prior.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.awb_arg_joint.argtypes = [C.c_void_p, C.POINTER(C.c_double),
                                    C.POINTER(C.c_double)]
        _lib = L
    return _lib


def libc_rand_snapshot():
    """State of this process's libc rand() stream (awb_libc_rand_snapshot)."""
    st = np.empty(RNG_WORDS, np.int32)
    _check(lib().awb_libc_rand_snapshot(st.ctypes.data))
    return st


def rng_draw(state, n):
    """n values of glibc's rand() generator from `state` (advanced in place)."""
    out = np.empty(n, np.int32)
    _check(lib().awb_rng_draw(state.ctypes.data, int(n), out.ctypes.data))
    return out


def _check(rc):
    if rc != 0:
        raise AwbError(lib().awb_last_error().decode())


def device_count():
    return lib().awb_device_count()


class Context(object):
    """One CUDA device + stream (``awb_ctx``)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib().awb_ctx_create(int(device), C.byref(self.h)))
        self.device = device

    def record(self, slot):
        _check(lib().awb_ctx_record(self.h, slot))

    def elapsed_ms(self, slot0, slot1):
        ms = C.c_float()
        _check(lib().awb_ctx_elapsed_ms(self.h, slot0, slot1, C.byref(ms)))
        return ms.value

    def sync(self):
        _check(lib().awb_ctx_sync(self.h))

    def close(self):
        if self.h:
            lib().awb_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


DEBUG_DTYPES = {
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
    "path": np.int32, "fsum": np.float64, "sink": np.float64,
}


class Batch(object):
    """A set of independent thread-sampling problems on one device
    (``awb_batch``).  ``problems`` is a list of problem dicts (see sim.py)."""

    def __init__(self, problems, ctx=None, keep_debug=False, checkpoint=False):
        self.ctx = ctx or default_context()
        self.n = len(problems)
        arr = (AwbProblem * self.n)()
        self._keep = []
        for i, d in enumerate(problems):
            p, keep = make_problem(d)
            arr[i] = p
            self._keep.append(keep)
        self._arr = arr
        self.h = C.c_void_p()
        _check(lib().awb_batch_create(self.ctx.h, self.n, arr,
                                      (KEEP_DEBUG if keep_debug else 0) |
                                      (CHECKPOINT if checkpoint else 0),
                                      C.byref(self.h)))
        self.ntrees = [arr[i].ntrees for i in range(self.n)]

    def close(self):
        if self.h:
            lib().awb_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- pipeline
    def upload(self):
        _check(lib().awb_batch_upload(self.h))
        return self

    def setup(self):
        _check(lib().awb_batch_setup(self.h))
        return self

    def forward(self, priors=None):
        ptr = None
        if priors is not None:
            if len(priors) != self.n:
                raise ValueError("need one prior (or None) per problem")
            self._priors = [None if p is None else
                            np.ascontiguousarray(p, np.float64) for p in priors]
            for i, p in enumerate(self._priors):
                if p is not None and p.size < max(int(self.nstates(i)[0]), 1):
                    raise ValueError("prior of problem %d is shorter than its "
                                     "first state space" % i)
            pa = (C.c_void_p * self.n)()
            for i, p in enumerate(self._priors):
                pa[i] = None if p is None else p.ctypes.data
            ptr = pa
        _check(lib().awb_batch_forward(self.h, ptr))
        return self

    def _rand_ptrs(self, rand_ints):
        self._rand = [np.ascontiguousarray(r, np.int32) for r in rand_ints]
        ra = (C.c_void_p * self.n)()
        for i, r in enumerate(self._rand):
            if len(r) < self.nsites(i):
                raise ValueError("need one rand() draw per site")
            ra[i] = r.ctypes.data
        return ra

    def upload_rand(self, rand_ints):
        _check(lib().awb_batch_upload_rand(self.h, self._rand_ptrs(rand_ints)))
        return self

    def traceback(self, rand_ints=None, rand_max=RAND_MAX, last_states=None):
        ra = None if rand_ints is None else self._rand_ptrs(rand_ints)
        ls = None
        if last_states is not None:
            self._ls = np.ascontiguousarray(last_states, np.int32)
            ls = self._ls.ctypes.data
        _check(lib().awb_batch_traceback(self.h, ra, int(rand_max), ls))
        return self

    def phase_probs(self):
        """Unphased data: evaluate P(phasing as given | sampled state) at the
        heterozygous sites (after traceback)."""
        _check(lib().awb_batch_phase_probs(self.h))
        return self

    def get_phase_probs(self, i=0):
        """[nsites] doubles, -1 where the unphased individual is not heterozygous."""
        out = np.empty(self.nsites(i), np.float64)
        _check(lib().awb_batch_get_phase_probs(self.h, i, out.ctypes.data))
        return out

    def sample_recombs(self, rng_states, rand_max=RAND_MAX):
        """Recombination points of the sampled paths (sample_recombinations,
        recomb.cpp:151-235).  rng_states: [n][RNG_WORDS] int32, one libc rand()
        state per problem (libc_rand_snapshot)."""
        st = np.ascontiguousarray(rng_states, np.int32).reshape(self.n, RNG_WORDS)
        _check(lib().awb_batch_sample_recombs(self.h, st.ctypes.data, int(rand_max)))
        return self

    def recombs(self, i=0):
        """(pos, node, time, draws): the recombination points of problem i and the
        number of rand() values the sampler consumed."""
        n, d = C.c_int(), C.c_int()
        _check(lib().awb_batch_get_recomb_count(self.h, i, C.byref(n), C.byref(d)))
        pos = np.empty(n.value, np.int32)
        node = np.empty(n.value, np.int32)
        time = np.empty(n.value, np.int32)
        _check(lib().awb_batch_get_recombs(self.h, i, n.value, pos.ctypes.data,
                                           node.ctypes.data, time.ctypes.data))
        return pos, node, time, d.value

    def sync(self):
        _check(lib().awb_batch_sync(self.h))
        return self

    def check_status(self):
        """Raise if the forward pass met a column whose norm is not positive
        (the reference asserts there, sample_thread.cpp:443-444,457-458)."""
        for i in range(self.n):
            s = self.status(i)
            if s >= 0:
                raise AwbError("problem %d: forward column %d is not positive"
                               % (i, s))
        return self

    # ---- info
    def nsites(self, i=0):
        return lib().awb_batch_nsites(self.h, i)

    def states_sites(self, i=0):
        return lib().awb_batch_states_sites(self.h, i)

    def total_states_sites(self):
        return sum(self.states_sites(i) for i in range(self.n))

    def h2d_bytes(self):
        return lib().awb_batch_h2d_bytes(self.h)

    def kernel_launches(self):
        return lib().awb_batch_kernel_launches(self.h)

    def segments(self):
        """Checkpointed table: (segments of the longest window, segment tables
        kept per window)."""
        return (lib().awb_batch_segments(self.h),
                lib().awb_batch_resident_segments(self.h))

    def fast_path(self):
        """'fast' or 'generic': the forward kernel this batch's shape selects."""
        return "fast" if lib().awb_batch_forward_kernel(self.h) else "generic"

    def timings(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _check(lib().awb_batch_timings(self.h, C.byref(a), C.byref(b),
                                       C.byref(c)))
        return dict(setup_ms=a.value, forward_ms=b.value, traceback_ms=c.value)

    KERNEL_CLASSES = ("kind", "block_setup", "tmatrix", "switch_setup", "emit",
                      "forward", "traceback", "recomb")

    def kernel_times(self, enable=True):
        """Start (and reset) per-kernel device timing (CUDA events around every
        launch on the batch's stream)."""
        _check(lib().awb_batch_kernel_times(self.h, 1 if enable else 0))
        return self

    def get_kernel_times(self):
        """dict kernel class -> ms since kernel_times(), plus 'forward_bytes'
        (algorithmic bytes of the forward launches) and 'forward_launches'."""
        ms = np.zeros(len(self.KERNEL_CLASSES), np.float32)
        fb, fl = C.c_double(), C.c_int()
        _check(lib().awb_batch_get_kernel_times(self.h, ms.ctypes.data,
                                                C.byref(fb), C.byref(fl)))
        out = {k: float(v) for k, v in zip(self.KERNEL_CLASSES, ms)}
        out["forward_bytes"] = fb.value
        out["forward_launches"] = fl.value
        return out

    # ---- results
    def path(self, i=0, out=None):
        """Sampled state path of problem i; `out`: a caller buffer (e.g. pinned)."""
        if out is None:
            out = np.empty(self.nsites(i), np.int32)
        _check(lib().awb_batch_get_path(self.h, i, out.ctypes.data))
        return out

    def logz(self, i=0):
        z = C.c_double()
        _check(lib().awb_batch_get_logz(self.h, i, C.byref(z)))
        return z.value

    def status(self, i=0):
        s = C.c_int()
        _check(lib().awb_batch_get_status(self.h, i, C.byref(s)))
        return s.value

    def fw(self, i=0):
        out = np.empty(lib().awb_batch_fw_doubles(self.h, i), np.float64)
        _check(lib().awb_batch_get_fw(self.h, i, out.ctypes.data))
        return out

    def nstates(self, i=0):
        out = np.empty(self.ntrees[i], np.int32)
        _check(lib().awb_batch_get_nstates(self.h, i, out.ctypes.data))
        return out

    def layout(self, i=0):
        B = self.ntrees[i]
        ro = np.empty(B + 1, np.int64)
        fo = np.empty(B + 1, np.int64)
        so = np.empty(B + 1, np.int64)
        _check(lib().awb_batch_get_layout(self.h, i, ro.ctypes.data,
                                          fo.ctypes.data, so.ctypes.data))
        return dict(row_off=ro, fw_off=fo, sw1_off=so)

    def debug(self, name, i=0):
        """A named per-block array (needs keep_debug for the sw_* copies)."""
        dt = np.dtype(DEBUG_DTYPES[name])
        nb = lib().awb_batch_debug_bytes(self.h, i, name.encode())
        if nb < 0:
            raise AwbError("unknown or unavailable array: " + name)
        out = np.empty(nb // dt.itemsize, dt)
        _check(lib().awb_batch_get_debug(self.h, i, name.encode(),
                                         out.ctypes.data, out.nbytes))
        return out


def forward_algorithm(problem, prior=None, ctx=None):
    """Forward table of one problem; returns (fw_flat, layout, logZ).

    Counterpart of argweaver_forward_algorithm (argweaverc.py:529-567)."""
    b = Batch([problem], ctx)
    try:
        b.upload().setup().forward(None if prior is None else [prior]).sync()
        b.check_status()
        return b.fw(0), b.layout(0), b.logz(0)
    finally:
        b.close()


class Sites(object):
    """A .sites alignment (``awb_sites``; reference ``Sites``,
    sequences.h:211-268): variant columns only.  ``read`` ~ read_sites,
    ``compress`` ~ find_compress_cols + compress_sites, ``sequences`` ~
    make_sequences_from_sites; ``packed()`` gives the arrays a problem dict takes
    instead of dense ``seqs``."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def read(cls, filename, region=None):
        h = C.c_void_p()
        a, b = region if region else (-1, -1)
        _check(lib().awb_sites_read(filename.encode(), int(a), int(b), C.byref(h)))
        return cls(h)

    @classmethod
    def from_columns(cls, positions, cols, start, end):
        positions = np.ascontiguousarray(positions, np.int32)
        cols = np.ascontiguousarray(cols, np.uint8)
        h = C.c_void_p()
        _check(lib().awb_sites_from_columns(
            cols.shape[1] if cols.ndim == 2 else 0, int(start), int(end),
            len(positions), positions.ctypes.data, cols.ctypes.data, C.byref(h)))
        return cls(h)

    nseqs = property(lambda self: lib().awb_sites_nseqs(self.h))
    ncols = property(lambda self: lib().awb_sites_ncols(self.h))
    start = property(lambda self: lib().awb_sites_start(self.h))
    end = property(lambda self: lib().awb_sites_end(self.h))

    def names(self):
        return [lib().awb_sites_name(self.h, i).decode() for i in range(self.nseqs)]

    def positions(self):
        n = self.ncols
        return (np.ctypeslib.as_array(lib().awb_sites_positions(self.h), (n,)).copy()
                if n else np.zeros(0, np.int32))

    def columns(self):
        n, k = self.ncols, self.nseqs
        return (np.ctypeslib.as_array(lib().awb_sites_columns(self.h), (n * k,))
                .reshape(n, k).copy() if n else np.zeros((0, k), np.uint8))

    def compress(self, compress):
        """False when the alignment cannot be compressed at this level."""
        rc = lib().awb_sites_compress(self.h, int(compress))
        if rc == 1:
            _check(rc)
        return rc == 0

    def mapping(self):
        n = lib().awb_sites_mapping_size(self.h)
        return (np.ctypeslib.as_array(lib().awb_sites_mapping(self.h), (n,)).copy()
                if n > 0 else np.zeros(0, np.int32))

    def sequences(self, default_char="A"):
        out = np.empty((self.nseqs, self.end - self.start), np.uint8)
        _check(lib().awb_sites_to_sequences(self.h, out.ctypes.data,
                                            ord(default_char)))
        return out

    def packed(self):
        """dict(var_pos, var_cols, nseqs, seqlen) relative to the region start"""
        return dict(var_pos=self.positions() - self.start, var_cols=self.columns(),
                    nseqs=self.nseqs, seqlen=self.end - self.start)

    def close(self):
        if self.h:
            lib().awb_sites_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pack_columns(seqs):
    """The variant-columns form of dense rows, as a .sites file holds them
    (sequences.cpp:136-158): the columns whose rows differ -- plus the all-'N'
    ones, which are masked sites.  Columns with one base in every row become the
    default character; the threading HMM does not tell such columns apart
    (emit.cpp:30-58, :807-828).  dict(var_pos, var_cols, nseqs, seqlen)."""
    seqs = np.asarray(seqs, np.uint8)
    var = np.nonzero((seqs != seqs[0]).any(axis=0) |
                     (seqs[0] == ord("N")))[0].astype(np.int32)
    return dict(var_pos=var, var_cols=np.ascontiguousarray(seqs[:, var].T),
                nseqs=seqs.shape[0], seqlen=seqs.shape[1])


def arg_likelihood(arg):
    """calc_arg_likelihood (total_prob.cpp:19-42) of a complete ARG (dict with
    the tree, model and sequence keys of a problem)."""
    d = dict(arg)
    d.setdefault("new_chrom", 0)
    d.setdefault("internal", 0)
    d.setdefault("seqids", np.arange((np.asarray(d["ptrees"]).shape[1] + 1) // 2))
    p, keep = make_problem(d)
    out = C.c_double()
    _check(lib().awb_arg_likelihood(C.byref(p), C.byref(out)))
    return out.value


def arg_prior(arg):
    """calc_arg_prior (total_prob.cpp:262-299)."""
    d = dict(arg)
    d.setdefault("new_chrom", 0)
    d.setdefault("internal", 0)
    d.setdefault("seqids", np.arange((np.asarray(d["ptrees"]).shape[1] + 1) // 2))
    p, keep = make_problem(d)
    out = C.c_double()
    _check(lib().awb_arg_prior(C.byref(p), C.byref(out)))
    return out.value


def sample_thread(problem, rand_ints, rand_max=RAND_MAX, ctx=None):
    """Forward + stochastic traceback; returns (path, logZ)."""
    b = Batch([problem], ctx)
    try:
        b.upload().setup().forward().traceback([rand_ints], rand_max).sync()
        b.check_status()
        return b.path(0), b.logz(0)
    finally:
        b.close()


def sample_thread_stream(batches, ctx=None, checkpoint=True, rand_max=RAND_MAX,
                         out=None):
    """Thread sampling over a stream of batches (a genome-wide run: one batch
    = the windows that fit on the device at once).

    `batches` yields `(problems, rand_ints)` pairs; for each the generator
    yields `(paths, logz)` -- lists with one entry per problem.  The host-only
    part of a batch (`awb_batch_create`: validation and layout) is done for
    batch n+1 while batch n runs on the device, so it is off the critical path;
    the copies and kernels of every batch are as in `Batch`.  `out(i)`, if
    given, returns the (e.g. pinned) int32 buffer for the path of problem i.
    """
    ctx = ctx or default_context()
    it = iter(batches)

    def create():
        try:
            problems, rands = next(it)
        except StopIteration:
            return None
        return Batch(problems, ctx, checkpoint=checkpoint), rands

    cur = create()
    try:
        while cur is not None:
            b, rands = cur
            cur = None
            try:
                b.upload().setup().forward().traceback(rands, rand_max)   # queued
                cur = create()                                            # overlaps
                b.sync()
                b.check_status()
                paths = [b.path(i, out=None if out is None else out(i))
                         for i in range(b.n)]
                logz = [b.logz(i) for i in range(b.n)]
            finally:
                b.close()
            yield paths, logz
    finally:
        if cur is not None:             # the consumer stopped early
            cur[0].close()
