"""In-tree build of libeasydgl_b200.so (sm_100a only).

``python -m easydgl_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting .so is git-ignored but travels to the GPU box with the
repo snapshot.  Each .cu is compiled to an object in parallel and cached by a content
hash of (flags, source, headers), so an up-to-date build is a no-op.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libeasydgl_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hs.append(os.path.join(ROOT, "include", "easydgl_b200.h"))
    return hs


def _hash(files, extra=""):
    h = hashlib.sha256((" ".join(CFLAGS) + extra).encode())
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())
            h.update(fh.read())
    return h.hexdigest()


def _run(cmd, verbose):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write(res.stderr)


def _compile_one(src, force, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".hash"
    want = _hash([src] + _headers())
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return obj, False
    _run([NVCC] + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src], verbose)
    with open(stamp, "w") as fh:
        fh.write(want)
    return obj, True


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile_one(s, force, verbose), _sources()))
    objs = [o for o, _ in res]
    if force or any(c for _, c in res) or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs, verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
