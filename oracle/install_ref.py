#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference environment, made available to the GPU box.

The reference path is pure Python (envs/vehicle.py, envs/network.py, envs/test_env.py -- SURVEY.md 8(a)); there is
nothing to compile.  "Installing" it means placing those three files, byte for byte, under oracle/_ref/envs/ -- a
directory that is git-ignored (no reference source ever enters the history) but is NOT gpurun-ignored, so it travels
to the GPU box with the working tree like our own built .so files.  bench.py then times the real
``TestEnv.my_step`` + ``obtain_state`` (reference envs/test_env.py:124,527) on the box's host cores next to the GPU
number (cpu_baseline.reference_python), which is what BASELINE.md section 3 asks for.

Run in the build container (needs /root/reference); __graft_entry__.build() calls it when the reference is present:

    python oracle/install_ref.py
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("DIRAL_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "envs")
FILES = ("vehicle.py", "network.py", "test_env.py")


def install(verbose: bool = False) -> bool:
    """Copy the three env files; returns False (and changes nothing) when the reference is not on this machine."""
    src_dir = os.path.join(REF, "envs")
    if not all(os.path.exists(os.path.join(src_dir, f)) for f in FILES):
        return False
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src_dir, f), os.path.join(DEST, f))
        with open(os.path.join(DEST, f), "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w") as fh:
        json.dump({"source": src_dir, "sha256": manifest}, fh, indent=1)
    if verbose:
        print("installed %s -> %s" % (", ".join(FILES), DEST))
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(DEST, f)) for f in FILES)


if __name__ == "__main__":
    sys.exit(0 if install(verbose=True) else 1)
