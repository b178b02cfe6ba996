"""In-tree build of libvr180_b200.so (nvcc, sm_100a only).  No GPU is needed to build.

    python vr180-convert_b200/build.py [--force] [--verbose]

The shared object is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import platform
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libvr180_b200.so"
OBJ_DIR = PKG / "build"
SOURCES = ["kernels.cu", "tiled.cu", "stream.cu", "api.cu", "pipeline.cu", "codec.cu"]
HOST_SOURCES = ["hostcopy.cpp"]  # plain C++ with per-file ISA flags (g++), linked into the same library
HOST_FLAGS = ["-O2", "-std=c++17", "-fPIC"] + (["-mavx2"] if platform.machine() in ("x86_64", "AMD64") else [])
HEADERS = ["chain.cuh", "chain_fast.cuh", "sampler.cuh", "tables.cuh", "common.cuh", "tiled.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libvr180_b200.so cannot be built")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for f in [*(CSRC / s for s in SOURCES), *(CSRC / s for s in HOST_SOURCES), *(CSRC / s for s in HEADERS), INCLUDE / "vr180_b200.h"]:
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS + HOST_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = OBJ_DIR / "fingerprint"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == fp:
        return LIB
    OBJ_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src: str) -> Path:
        obj = OBJ_DIR / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJ_DIR / (Path(src).stem + ".ptxas.log")).write_text(r.stderr)
        if verbose or r.returncode:
            sys.stderr.write(r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed for {src}")
        return obj

    def compile_host(src: str) -> Path:
        obj = OBJ_DIR / (Path(src).stem + ".o")
        cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
        r = subprocess.run([cxx, *HOST_FLAGS, "-c", str(CSRC / src), "-o", str(obj)], capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stderr)
        if r.returncode:
            raise RuntimeError(f"{cxx} failed for {src}")
        return obj

    with ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    objs += [compile_host(s) for s in HOST_SOURCES]
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static", "-ldl",
           "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stderr)
        raise RuntimeError("link failed")
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
