"""Build the C-ABI shared library in-tree with nvcc for sm_100a.

``python -m peppa_b200.build`` or ``peppa_b200.build.build()``.  Two libraries come out of the same sources:
``peppa_b200/csrc/libpeppa_b200.so`` (the product: no debug exports, no mutable global state) and
``peppa_b200/csrc/libpeppa_b200_measure.so`` (``-DPB2_MEASURE``: the ``pb2_debug_*`` tile-shape / cluster-variant
selectors and knock-out switches used by ``tools/`` and the variant tests).  Both are git-ignored but travel to the
GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
INCLUDE = os.path.join(ROOT, "include")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libpeppa_b200.so")
OBJ_MEASURE = os.path.join(CSRC, "build_measure")
LIB_MEASURE = os.path.join(CSRC, "libpeppa_b200_measure.so")

LIB_SOURCES = ["host_util.cu", "triplet.cu", "rowstats.cu", "sim.cu", "gradgemm.cu", "step.cu", "proj.cu", "collective.cu", "sampler.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-Xptxas", "-v",
    "-I", INCLUDE, "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths, extra=()) -> str:
    # what the object depends on, not where the checkout lives: the GPU box runs a copy under another path, and its
    # prebuilt objects must count as current there
    h = hashlib.sha256()
    h.update(" ".join([f for f in list(NVCC_FLAGS) + list(extra) if not os.path.isabs(f)]).encode())
    for p in sorted(paths, key=os.path.basename):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())
            h.update(f.read())
    return h.hexdigest()


def _compile(src: str, verbose: bool, objdir: str = OBJ, extra=()) -> str:
    obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "peppa_b200.h"))
    stamp = obj + ".sha"
    want = _digest([os.path.join(CSRC, src)] + headers, extra)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(objdir, os.path.splitext(src)[0] + ".log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed on {src}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(stamp, "w") as f:
        f.write(want)
    return obj


def _link(lib: str, objs) -> None:
    newest = max(os.path.getmtime(o) for o in objs)
    if not os.path.exists(lib) or os.path.getmtime(lib) < newest:
        tmp = lib + f".tmp{os.getpid()}"
        subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs, "-ldl"], check=True)
        os.replace(tmp, lib)        # atomic: another rank may be dlopen()ing the old file


FAST_SRC = os.path.join(CSRC, "torch_fast.cpp")
FAST_LIB = os.path.join(CSRC, "_pb2_fast.so")


def build_fast(verbose: bool = False) -> str:
    """The C++ autograd glue for the launch-bound training step (csrc/torch_fast.cpp): g++ against torch's headers and
    the product library (no CUDA code in it).  Rebuilt when its source, the header or torch's version changes."""
    import sysconfig

    import torch
    tdir = os.path.dirname(torch.__file__)
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_pb2_fast",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-DTORCH_API_INCLUDE_EXTENSION_H",
           "-I", os.path.join(tdir, "include"), "-I", os.path.join(tdir, "include", "torch", "csrc", "api", "include"),
           "-I", sysconfig.get_paths()["include"], "-I", "/usr/local/cuda/include", "-I", INCLUDE, FAST_SRC, "-o", FAST_LIB + ".tmp",
           "-L", os.path.join(tdir, "lib"), "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
           "-L", CSRC, "-lpeppa_b200", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + os.path.join(tdir, "lib")]
    # the stamp names what the binary depends on, not where the checkout lives (the GPU box runs a copy elsewhere)
    h = hashlib.sha256(f"{torch.__version__} abi{int(torch._C._GLIBCXX_USE_CXX11_ABI)} py{sys.version_info[:2]} -O2 c++17".encode())
    for p_ in (FAST_SRC, os.path.join(INCLUDE, "peppa_b200.h")):
        with open(p_, "rb") as f:
            h.update(f.read())
    stamp, want = FAST_LIB + ".sha", h.hexdigest()
    if os.path.exists(FAST_LIB) and os.path.exists(stamp) and open(stamp).read() == want:
        return FAST_LIB
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(OBJ, "torch_fast.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed on torch_fast.cpp")
    if verbose:
        sys.stderr.write(r.stderr)
    os.replace(FAST_LIB + ".tmp", FAST_LIB)
    with open(stamp, "w") as f:
        f.write(want)
    return FAST_LIB


def build(verbose: bool = False, measure: bool = True) -> str:
    """Compile (only what changed: per-source hashes) and link; returns the product library's path."""
    jobs = [(s, OBJ, ()) for s in LIB_SOURCES]
    if measure:
        jobs += [(s, OBJ_MEASURE, ("-DPB2_MEASURE",)) for s in LIB_SOURCES]
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(OBJ_MEASURE, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        objs = list(ex.map(lambda j: _compile(j[0], verbose, j[1], j[2]), jobs))
    _link(LIB, objs[:len(LIB_SOURCES)])
    if measure:
        _link(LIB_MEASURE, objs[len(LIB_SOURCES):])
    build_fast(verbose)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
