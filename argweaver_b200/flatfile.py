"""Reader/writer for the "AWF1" named-array container.

A flat thread-sampling problem (model, sequences, local trees, SPRs, uniforms)
and, optionally, the outputs computed for it are stored as a list of named
numpy arrays.  The C twin is ``oracle/flatio.h``; the layout is

    b"AWF1" then records:  u32 name_len | name | u8 dtype | u32 ndim |
                           u64 dims[ndim] | raw little-endian data

dtype codes: 0=int32 1=float64 2=uint8 3=int64.
"""

import struct

import numpy as np

_DTYPES = {0: np.dtype("<i4"), 1: np.dtype("<f8"), 2: np.dtype("u1"),
           3: np.dtype("<i8")}
_CODES = {np.dtype("int32"): 0, np.dtype("float64"): 1, np.dtype("uint8"): 2,
          np.dtype("int64"): 3}


def read_awf(filename):
    """Return a dict name -> numpy array for an AWF1 file."""
    with open(filename, "rb") as f:
        buf = f.read()
    if buf[:4] != b"AWF1":
        raise ValueError("%s: not an AWF1 file" % filename)
    pos = 4
    out = {}
    while pos < len(buf):
        (nl,) = struct.unpack_from("<I", buf, pos)
        pos += 4
        name = buf[pos:pos + nl].decode()
        pos += nl
        dt, nd = struct.unpack_from("<BI", buf, pos)
        pos += 5
        dims = struct.unpack_from("<%dQ" % nd, buf, pos)
        pos += 8 * nd
        dtype = _DTYPES[dt]
        count = int(np.prod(dims, dtype=np.int64)) if nd else 1
        arr = np.frombuffer(buf, dtype=dtype, count=count, offset=pos)
        pos += count * dtype.itemsize
        out[name] = arr.reshape(dims).copy()
    return out


def write_awf(filename, arrays):
    """Write a dict name -> array as an AWF1 file."""
    with open(filename, "wb") as f:
        f.write(b"AWF1")
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            if arr.dtype == np.bool_:
                arr = arr.astype(np.uint8)
            code = _CODES[arr.dtype]
            nb = name.encode()
            shape = arr.shape if arr.ndim else (1,)
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<BI", code, len(shape)))
            f.write(struct.pack("<%dQ" % len(shape), *shape))
            f.write(arr.tobytes())


def load_problem(filename):
    """Load an AWF1 or .npz problem file into a dict of arrays."""
    if str(filename).endswith(".npz"):
        with np.load(filename) as z:
            return {k: z[k] for k in z.files}
    return read_awf(filename)
