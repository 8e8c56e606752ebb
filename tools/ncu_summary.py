#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline counters of the raw
page and a per-source-line roll-up of the source page.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--lines 25]
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    for v in rows[2:]:
        print("== kernel:", v[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
        for i, h in enumerate(hdr):
            if h in KEYS or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                try:
                    x = float(v[i].replace(",", ""))
                except ValueError:
                    continue
                if h.startswith("smsp__average_warps_issue_stalled") and x < 0.15:
                    continue
                print(f"  {h:90s} {x:16.3f} {units[i]}")
    src = ncu(rep, "source")
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    data = rows[2:]
    tot_w = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
    tot_t = sum(int(r[ix["Thread Instructions Executed"]] or 0) for r in data)
    tot_s = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print(f"== SASS: {len(data)} instructions ({len(data) * 16 / 1024:.1f} KB), warp-inst {tot_w:.3e}, thread-inst {tot_t:.3e}, avg threads/inst {tot_t / max(tot_w, 1):.2f}")
    ops = defaultdict(lambda: [0, 0])
    for r in data:
        m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        op = m.group(2).split(".")[0] if m else "?"
        ops[op][0] += int(r[ix["Instructions Executed"]] or 0)
        ops[op][1] += int(r[ix["Thread Instructions Executed"]] or 0)
    print("== opcode mix (warp-inst share, avg active threads)")
    for op, (w, t) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:nlines]:
        print(f"  {op:12s} {100 * w / tot_w:6.2f}%  {t / max(w, 1):5.1f}")
    print("== regions of 64 SASS instructions (share of warp-inst, avg threads, stall samples, no_inst samples)")
    for k in range(0, len(data), 64):
        seg = data[k:k + 64]
        w = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
        t = sum(int(r[ix["Thread Instructions Executed"]] or 0) for r in seg)
        s = sum(int(r[ix["# Samples"]] or 0) for r in seg)
        ni = sum(int(r[ix["stall_no_inst"]] or 0) for r in seg)
        if w * 200 > tot_w:
            print(f"  [{k:5d}] {100 * w / tot_w:6.2f}%  thr {t / max(w, 1):5.1f}  samples {100 * s / max(tot_s, 1):5.1f}%  no_inst {ni}")


if __name__ == "__main__":
    main()
