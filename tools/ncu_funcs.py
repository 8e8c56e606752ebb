#!/usr/bin/env python
"""Roll an .ncu-rep source page up by the device FUNCTION each source line
belongs to (inlined code is attributed to the inlined function).

    python tools/ncu_funcs.py gpurun_out/prof.ncu-rep [rays]
"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def func_map(path):
    starts = []
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__).*?\b([A-Za-z_]\w*)\s*\(")
    lines = open(path, errors="replace").read().split("\n")
    for i, l in enumerate(lines, 1):
        m = pat.match(l)
        if m:
            starts.append((i, m.group(1)))
    return starts


def main():
    rep = sys.argv[1]
    rays = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    maps = {}
    fname, fpath, hdr = "", "", None
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]
            fname = os.path.basename(fpath)
            local = os.path.join(ROOT, "ray_tracing_b200", "csrc", fname)
            if fname not in maps:
                maps[fname] = func_map(local) if os.path.exists(local) else []
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or r[0] == "":
            continue
        try:
            w = float(r[hdr["Instructions Executed"]].replace(",", ""))
            t = float(r[hdr["Thread Instructions Executed"]].replace(",", ""))
        except Exception:
            continue
        ln = int(r[0])
        fn = fname
        for s, name in maps.get(fname, []):
            if s <= ln:
                fn = name
            else:
                break
        a = agg.setdefault(fn, [0.0, 0.0])
        a[0] += w
        a[1] += t
    tw = sum(a[0] for a in agg.values()) or 1
    print(f"{'function':28s} {'warp-inst%':>10s} {'thr/inst':>8s}" + (f" {'slots/ray':>10s} {'thread-inst/ray':>16s}" if rays else ""))
    for fn, (w, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if w * 500 < tw:
            continue
        line = f"{fn:28s} {100 * w / tw:10.2f} {t / max(w, 1):8.1f}"
        if rays:
            line += f" {w * 32 / rays:10.1f} {t / rays:16.1f}"
        print(line)
    if rays:
        print(f"total slots/ray {tw * 32 / rays:.0f}")


if __name__ == "__main__":
    main()
