"""Builds diral_b200/libdiral_env.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The shared object is git-ignored but travels to the GPU box with the working-tree snapshot, so the
box never needs to compile.  `python -m diral_b200._build [--force]` rebuilds by hand.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libdiral_env.so")
SOURCES = ["diral_api.cu", "diral_step_group.cu", "diral_step_block.cu", "diral_step_row.cu", "diral_step_pair.cu", "diral_aux.cu", "diral_wire.cu",
           "diral_host.cpp"]
HEADERS = [os.path.join(CSRC, "diral_dev.cuh"), os.path.join(CSRC, "diral_launch.h"), os.path.join(CSRC, "diral_host.h"),
           os.path.join(os.path.dirname(HERE), "include", "diral_env.h")]
HASH = LIB + ".srchash"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-Xptxas", "-v"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libdiral_env.so cannot be built here")
    return exe


def source_hash() -> str:
    """sha256 over the flags and every source / header: what the built library is checked against at load time
    (content, not mtimes -- the library travels to the GPU box next to a fresh copy of the sources)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def up_to_date() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(HASH)):
        return False
    with open(HASH) as f:
        return f.read().strip() == source_hash()


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compile the library.  `defines` (e.g. ["DIRAL_MIN_BLOCKS=16"]) and `out` build an experimental
    variant next to the default one (tuning runs select it with DIRAL_ENV_LIB=<path>)."""
    if defines or out:
        out = out or LIB.replace(".so", "_" + "_".join(d.replace("=", "") for d in defines) + ".so")
        cmd = [nvcc()] + NVCC_FLAGS + ["-shared", "-o", out] + ["-D" + d for d in defines] \
            + [os.path.join(CSRC, s) for s in SOURCES]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(out + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building %s" % out)
        return out
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + HEADERS):
            jobs.append((src, o, subprocess.Popen([nvcc()] + NVCC_FLAGS + ["-c", s, "-o", o], stdout=subprocess.PIPE,
                                                  stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, o, proc in jobs:            # the translation units compile side by side (the slot kernels take minutes)
        out, _ = proc.communicate()
        with open(o + ".log", "w") as f:
            f.write(out)
        if verbose or proc.returncode:
            sys.stderr.write(out)
        if proc.returncode:
            failed.append(src)
    if failed:
        raise RuntimeError("nvcc failed on %s" % ", ".join(failed))
    if force or _stale(LIB, objs) or not up_to_date():
        subprocess.check_call([nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread"])
        with open(HASH, "w") as f:
            f.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
