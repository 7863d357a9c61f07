#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): headline metrics + stall reasons + hottest instructions.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_top]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__sass_thread_inst_executed_op_ffma_pred_on.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    for k in keys:
        if k in m:
            print(f"{k:75s} {m[k][:90]} {u.get(k, '')}")
    print("-- stall reasons per issue-active (warp-level) --")
    st = {k: float(v) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")}
    for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):22s} {v:.3f}")
    src = page(rep, "source")
    h = src[1]
    col = {x: i for i, x in enumerate(h)}
    data = src[2:]
    tot = sum(int(r[col["# Samples"]]) for r in data)
    byop, ex = Counter(), Counter()
    for r in data:
        toks = r[col["Source"]].split()
        op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")).split(".")[0]
        byop[op] += int(r[col["# Samples"]])
        ex[op] += int(r[col["Instructions Executed"]])
    allex = sum(ex.values())
    print(f"-- opcode mix (executed warp-instr {allex:.3e}, samples {tot}) --")
    for op, s in byop.most_common(10):
        print(f"  {op:8s} executed {100 * ex[op] / allex:5.1f}%  samples {100 * s / tot:5.1f}%")
    stall_cols = [x for x in h if x.startswith("stall_") and "Not Issued" in x]
    agg = Counter()
    for r in data:
        for x in stall_cols:
            try:
                agg[x] += int(r[col[x]])
            except ValueError:
                pass
    print("-- not-issued samples by reason --")
    for k, v in agg.most_common(8):
        print(f"  {k:38s} {v:9d} ({100 * v / tot:.1f}% of samples)")
    print(f"-- top {n_top} instructions by samples --")
    for r in sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:n_top]:
        nz = {x.replace("stall_", ""): r[col[x]] for x in h if x.startswith("stall_") and "Not" not in x and r[col[x]] not in ("0", "")}
        top = sorted(nz.items(), key=lambda kv: -int(kv[1]))[:4]
        print(f"  {int(r[col['# Samples']]):8d}  {r[col['Source']].strip()[:58]:58s} {top}")


if __name__ == "__main__":
    main()
