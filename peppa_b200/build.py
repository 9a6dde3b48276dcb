"""Build the C-ABI shared library (and the stand-alone self-test) in-tree with nvcc for sm_100a.

``python -m peppa_b200.build`` or ``peppa_b200.build.build()``.  The outputs
(``peppa_b200/csrc/libpeppa_b200.so``, ``peppa_b200/csrc/pb2_selftest``) are git-ignored but travel
to the GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
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
SELFTEST = os.path.join(CSRC, "pb2_selftest")

LIB_SOURCES = ["host_util.cu", "triplet.cu", "rowstats.cu", "sim.cu", "gradgemm.cu", "step.cu", "proj.cu"]
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


def _digest(paths) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "peppa_b200.h"))
    stamp = obj + ".sha"
    want = _digest([os.path.join(CSRC, src)] + headers)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(OBJ, os.path.splitext(src)[0] + ".log")
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


def build(verbose: bool = False, selftest: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = list(LIB_SOURCES) + (["selftest.cu"] if selftest and os.path.exists(os.path.join(CSRC, "selftest.cu")) else [])
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = dict(zip(srcs, ex.map(lambda s: _compile(s, verbose), srcs)))
    lib_objs = [objs[s] for s in LIB_SOURCES]
    newest = max(os.path.getmtime(o) for o in lib_objs)
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *lib_objs]
        subprocess.run(cmd, check=True)
    if "selftest.cu" in objs:
        if not os.path.exists(SELFTEST) or os.path.getmtime(SELFTEST) < max(newest, os.path.getmtime(objs["selftest.cu"])):
            cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-o", SELFTEST, objs["selftest.cu"], *lib_objs]
            subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
