"""Build libemer_b200.so (sm_100a only) in-tree with nvcc.

The shared library is the product: ``emernerf_b200/lib/libemer_b200.so`` is git-ignored but
travels to the GPU box with the gpurun snapshot.  ``python -m emernerf_b200.build`` (re)builds it;
``__graft_entry__.build()`` calls :func:`build_library`.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libemer_b200.so")
OBJ_DIR = os.path.join(PKG, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                     # FMAs are explicit (fmaf); see csrc/common.cuh
    "-Xcompiler", "-fPIC",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]
# compile-time experiment switches of the kernels (see the comment at each #ifndef in csrc/): passed through
# from the environment so that an A/B build is one command; the default build defines none of them
for _flag in ("EMER_TC_ELECT_ONE", "EMER_WARP_ARRIVE", "EMER_CHAIN_STAGE"):
    if os.environ.get(_flag):
        NVCC_FLAGS.append(f"-D{_flag}={os.environ[_flag]}")


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sources() + sorted(
            os.path.join(d, f) for d in (CSRC, os.path.join(ROOT, "include"))
            for f in os.listdir(d) if f.endswith((".h", ".cuh"))):
        h.update(os.path.relpath(p, ROOT).encode())      # relative: the stamp is valid on any checkout path
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(f for f in NVCC_FLAGS if not os.path.isabs(f)).encode())
    return h.hexdigest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.stamp")
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB_PATH
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
