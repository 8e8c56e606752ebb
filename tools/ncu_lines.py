#!/usr/bin/env python
"""Per-CUDA-source-line roll-up of an .ncu-rep (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [N]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname = ""
    hdr = None
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {n: i for i, n in enumerate(r)}
            continue
        if hdr is None or r[0] == "":
            continue
        def g(name):
            try:
                return float(r[hdr[name]].replace(",", ""))
            except Exception:
                return 0.0
        w = g("Instructions Executed")
        t = g("Thread Instructions Executed")
        s = g("# Samples")
        lines.append((w, t, s, fname, r[0], r[1].strip()[:110]))
    tw = sum(x[0] for x in lines) or 1
    ts = sum(x[2] for x in lines) or 1
    print(f"total warp-inst {tw:.3e}; lines with metrics: {sum(1 for x in lines if x[0])}")
    for w, t, s, f, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
        print(f"{100 * w / tw:6.2f}%  thr {t / max(w, 1):5.1f}  smp {100 * s / ts:5.1f}%  {f}:{ln:>4s}  {src}")


if __name__ == "__main__":
    main()
