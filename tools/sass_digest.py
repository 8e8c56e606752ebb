#!/usr/bin/env python
"""md5 of the device code (SASS text) of every CUDA object under build/obj.

A host-only edit must leave these digests unchanged; an edit that is meant to be
compiled out by default (an #ifdef'd experiment) too.  Handy when no GPU is at
hand to re-run the parity tests:

    python tools/sass_digest.py > /tmp/before; <edit>; make; python tools/sass_digest.py | diff /tmp/before -
"""
import glob
import hashlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SKIP = ("Fatbin", "=====", "identifier", "compile_size", "producer", "host", "arch =", "code version")

for obj in sorted(glob.glob(os.path.join(ROOT, "build", "obj", "rt_*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    if "Function :" not in out:
        continue
    text = "\n".join(l for l in out.split("\n") if not l.startswith(SKIP) and "identifier" not in l)
    print(hashlib.md5(text.encode()).hexdigest(), os.path.basename(obj))
