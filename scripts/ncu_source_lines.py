#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` export (SASS rows with stall samples) per CUDA source line.

usage: ncu_source_lines.py <source.csv> <cubin> <kernel-substring> [top]
The cubin (cuobjdump -xelf <name> lib.so) must come from the build that was profiled and be compiled with -lineinfo: nvdisasm
--print-line-info gives the address -> file:line map that the CSV export lacks."""
import collections
import csv
import re
import subprocess
import sys


def main():
    src_csv, cubin, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    sass = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True, check=True).stdout.split("\n")
    off2line, cur, inside = {}, None, False
    for line in sass:
        t = line.strip()
        if t.startswith(".section"):
            inside = ".text." in t and kern in t
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", line)
        if m:
            off2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    base = int(rows[2][0], 16)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    samples, instr = collections.Counter(), collections.Counter()
    stalls = collections.defaultdict(collections.Counter)
    for r in rows[2:]:
        key = off2line.get(int(r[0], 16) - base)
        samples[key] += int(r[ix["# Samples"]])
        instr[key] += int(r[ix["Instructions Executed"]])
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            if v:
                stalls[key][c[6:]] += v
    tot, tot_i = sum(samples.values()), sum(instr.values())
    print(f"total warp-level instructions {tot_i}, stall samples {tot}")
    print("| source line | samples | % | warp instructions | % | top stall reasons |")
    print("|---|---:|---:|---:|---:|---|")
    for key, v in samples.most_common(top):
        why = ", ".join(f"{c} {n}" for c, n in stalls[key].most_common(3))
        name = f"{key[0]}:{key[1]}" if key else "?"
        print(f"| {name} | {v} | {100 * v / tot:.1f} | {instr[key]} | {100 * instr[key] / tot_i:.1f} | {why} |")


if __name__ == "__main__":
    main()
