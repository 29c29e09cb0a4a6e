#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full --import-source on) into text: per-launch headline metrics, warp-stall
breakdown, SASS opcode mix and the hottest source lines.   usage: summarize_ncu.py report.ncu-rep [units_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, unit_row = raw[0], raw[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("== launches ==")
    for r in raw[2:]:
        print(r[ix["Kernel Name"]][:70], "| grid", r[ix.get("Grid Size", 0)], "block", r[ix.get("Block Size", 0)])
        for k in KEYS:
            if k in ix:
                print(f"    {k:62s} {r[ix[k]]:>16s} {unit_row[ix[k]]}")
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    kern, cur, hdr2 = {}, None, None
    for r in src:
        if not r:
            continue
        if r[0] == "Function Name":
            cur = r[1]
            kern.setdefault(cur, {"lines": [], "sass": []})
        elif r[0] == "Line No":
            hdr2 = r
        elif cur and hdr2 and len(r) > 10:
            (kern[cur]["lines"] if r[0] != "" else kern[cur]["sass"]).append(r)
    for name, d in kern.items():
        h = {}
        for i, c in enumerate(hdr2):
            h.setdefault(c, i)
        ie, isamp = h["Instructions Executed"], h["# Samples"]
        tot = sum(num(r[ie]) for r in d["sass"]) or 1
        ts = sum(num(r[isamp]) for r in d["sass"]) or 1
        print(f"\n== {name[:90]} ==  warp-instructions {tot}" + (f"  = {tot / units:.1f} per unit" if units else ""))
        stall = collections.Counter()
        for i, c in enumerate(hdr2):
            if c.startswith("stall_") and "Not Issued" not in c and i < len(d["sass"][0]):
                stall[c] += sum(num(r[i]) for r in d["sass"] if i < len(r))
        ssum = sum(stall.values()) or 1
        print("  stalls: " + ", ".join(f"{k[6:]} {100 * v / ssum:.1f}%" for k, v in stall.most_common(8)))
        ops = collections.Counter()
        for r in d["sass"]:
            t = r[3].strip().split()
            if not t:
                continue
            o = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            ops[o] += num(r[ie])
        print("  opcode mix: " + ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in ops.most_common(14)))
        print("  hottest source lines (share of instructions | share of stall samples):")
        for r in sorted(d["lines"], key=lambda r: -num(r[ie]))[:22]:
            print(f"    {r[0]:>4s} {100 * num(r[ie]) / tot:5.1f}% {100 * num(r[isamp]) / ts:5.1f}%  {r[1].strip()[:100]}")


if __name__ == "__main__":
    main()
