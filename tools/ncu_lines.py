#!/usr/bin/env python
"""Per-CUDA-source-line sample / instruction totals of an .ncu-rep captured with --import-source on (-lineinfo build).
usage: python tools/ncu_lines.py gpurun_out/x.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, hdr = None, None
    agg = defaultdict(lambda: [0, 0, "", defaultdict(int)])
    for r in rows:
        if len(r) == 2 and r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            isamp, iexec = hdr.index("# Samples"), hdr.index("Instructions Executed")
            stall_cols = [(i, h.replace("stall_", "")) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        line = r[0]
        if r[1]:
            agg[(fname, line)][2] = r[1].strip()
        try:
            s, e = int(r[isamp] or 0), int(r[iexec] or 0)
        except ValueError:
            continue
        a = agg[(fname, line)]
        a[0] += s
        a[1] += e
        for i, nm in stall_cols:
            try:
                a[3][nm] += int(r[i] or 0)
            except ValueError:
                pass
    tot_s = sum(a[0] for a in agg.values())
    tot_e = sum(a[1] for a in agg.values())
    print(f"total samples {tot_s}, executed warp instructions {tot_e:.4e}")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n_top]:
        top = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
        tops = " ".join(f"{k}:{100 * v / max(a[0], 1):.0f}%" for k, v in top)
        print(f"{100 * a[0] / tot_s:6.2f}% smp {100 * a[1] / tot_e:6.2f}% ins  {f}:{ln:>4s}  {a[2][:70]:70s} {tops}")


if __name__ == "__main__":
    main()
