"""In-tree builds of the native libraries (explicit nvcc / g++ commands).

    libargweaver_b200.so   CUDA sm_100a kernels + the C ABI (include/argweaver_b200.h)
    libawb_sim.so          host-only synthetic input generator (dsmc_sim.cpp)

The libraries are written next to their sources in argweaver_b200/csrc/ so that
they travel with the repository snapshot to the GPU box; nothing is JIT-cached.
"""

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")

CUDA_LIB = os.path.join(CSRC, "libargweaver_b200.so")
SIM_LIB = os.path.join(CSRC, "libawb_sim.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    # internal references bind inside the library: a host program that defines
    # the reference's own symbols of the same names (arg-sample) cannot interpose
    "-Xlinker", "-Bsymbolic",
    "--fmad=true",
]

CUDA_SOURCES = ["awb_api.cu", "awb_compat.cu", "awb_totalprob.cu", "awb_sites.cpp"]
CUDA_DEPS = ["awb_setup.cuh", "awb_forward.cuh", "awb_forward_fast.cuh", "awb_traceback.cuh",
             "awb_emit.cuh", "awb_common.cuh", "awb_layout.h", "awb_recomb.cuh"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def find_nvcc():
    nvcc = shutil.which("nvcc")
    if nvcc:
        return nvcc
    cand = "/usr/local/cuda/bin/nvcc"
    if os.path.exists(cand):
        return cand
    raise RuntimeError("nvcc not found")


def build_sim(force=False, verbose=False):
    src = os.path.join(CSRC, "dsmc_sim.cpp")
    if force or _stale(SIM_LIB, [src]):
        cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++11", "-o", SIM_LIB, src]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return SIM_LIB


def build_cuda(force=False, verbose=False, extra_flags=()):
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, d) for d in CUDA_DEPS] + [
        os.path.join(INCLUDE, "argweaver_b200.h")]
    if force or _stale(CUDA_LIB, deps):
        cmd = [find_nvcc()] + NVCC_FLAGS + list(extra_flags) + [
            "-I", INCLUDE, "-I", CSRC, "-o", CUDA_LIB] + srcs + ["-lcudart"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return CUDA_LIB


def build_all(force=False, verbose=False):
    build_sim(force, verbose)
    build_cuda(force, verbose)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv, verbose=True)
