#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by source line.

usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > k.csv
       python tools/ncu_lines.py k.csv [top]
Prints, per (file, line): stall samples, warp instructions executed, source text — heaviest first.
"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fpath, out = None, []
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] in ("Function Name", "Kernel Name") or hdr is None:
            continue
        if r[0].isdigit():
            try:
                out.append((fpath, int(r[0]), int(r[4] or 0), int(r[7] or 0), r[1].strip()))
            except ValueError:
                pass
    tot_s = sum(o[2] for o in out) or 1
    tot_i = sum(o[3] for o in out) or 1
    print(f"total samples {tot_s}, warp instructions {tot_i}")
    for f, ln, s, i, src in sorted(out, key=lambda o: -o[2])[:top]:
        print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {f}:{ln:<4d} {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
