"""Build recipe for the C-ABI CUDA library (``libb2dt.so``), sm_100a only.

Every ``csrc/*.cu`` is compiled by plain ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` (in
parallel) and linked in-tree next to this file, so the built ``.so`` travels with the repository snapshot
to the GPU box.  No driver library is linked: ``cuTensorMapEncodeTiled`` is resolved at run time through
``cudaGetDriverEntryPoint``.
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb2dt.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the b2dt CUDA library cannot be built")
    return nvcc


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as fh:
            h.update(p.encode() + b"\0" + fh.read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile and link ``libb2dt.so`` if sources changed.  Returns the library path."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "b2dt.h")]
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()
    hdr_digest = _digest([d for d in deps if not d.endswith(".cu")])

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        tag = obj + ".sha"
        want = _digest([src]) + hdr_digest
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read() == want:
            return obj, ""
        cmd = [nvcc, *ARCH, *FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(tag, "w") as fh:
            fh.write(want)
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    r = subprocess.run([nvcc, *ARCH, "-shared", "-o", LIB, *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
