#!/usr/bin/env python
"""Per-kernel digest of the device code, for review round over round: registers, stack
frame and spills (ptxas log of the build), SASS instruction count and opcode histogram
(cuobjdump) of the render kernels of the exact and the fast build.

    make && python tools/sass_kernel_digest.py > profiles/rNN_sass_digest.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ("render_queued_kernel", "render_persistent_kernel", "render_wavefront_kernel", "render_pixel_kernel")


def ptxas(log):
    out, cur = {}, None
    for line in open(log):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur and ("spill" in line or "Used" in line):
            out[cur].append(line.strip().replace("ptxas info    : ", ""))
    return out


def main():
    for build in ("exact", "fast"):
        obj = os.path.join(ROOT, "build", "obj", f"rt_render_{build}.o")
        info = ptxas(os.path.join(ROOT, "build", "obj", f"rt_render_{build}.ptxas.log"))
        names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        funcs = re.findall(r"Function : (\S+)", names)
        for f in funcs:
            dem = subprocess.run(["cu++filt", f], capture_output=True, text=True).stdout.strip() or f
            if not any(k in dem for k in KERNELS):
                continue
            sass = subprocess.run(["cuobjdump", "-sass", "-fun", f, obj], capture_output=True, text=True).stdout
            ops = collections.Counter()
            n = 0
            for line in sass.split("\n"):
                m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
                if m:
                    ops[m.group(1)] += 1
                    n += 1
            print(f"== {dem}  [{build}]")
            for l in info.get(f, []):
                print("   " + l)
            print(f"   {n} SASS instructions ({n * 16 / 1024:.1f} KB)")
            print("   " + "  ".join(f"{op} {c}" for op, c in ops.most_common(28)))
            print()


if __name__ == "__main__":
    main()
