#!/usr/bin/env python
"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) and duration of every kernel in the given
ncu --set full reports and merge them into profiles/r2_dram_traffic.json, which bench.py reads for `roofline.traffic`.
usage: python tools/ncu_traffic.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def short(name):
    m = re.match(r"(?:void )?(?:niq::)?(k_\w+)(<[^>]*>)?", name)
    if not m:
        return name
    targ = m.group(2) or ""
    targ = re.sub(r"niq::", "", targ).replace(", TileRay", "").replace("(int)", "")
    return m.group(1) + targ


def main():
    rec = {}
    if os.path.exists(OUT):
        rec = json.load(open(OUT))
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = short(r[col["Kernel Name"]])
            rd = float(r[col["dram__bytes_read.sum"]]) * UNIT[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]]) * UNIT[units[col["dram__bytes_write.sum"]]]
            dur = float(r[col["gpu__time_duration.sum"]]) * TIME[units[col["gpu__time_duration.sum"]]]
            if rd != rd or wr != wr:
                continue
            prev = rec.get(name)
            if prev is None or rd + wr > prev["bytes_per_launch"]:          # keep the largest launch of each kernel
                rec[name] = {"bytes_per_launch": rd + wr, "read": rd, "write": wr, "duration_s": dur,
                             "dram_GBps": (rd + wr) / dur / 1e9 if dur > 0 else None, "grid": int(r[col["launch__grid_size"]]),
                             "report": os.path.basename(rep)}
    json.dump(rec, open(OUT, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
