#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and stall samples from an .ncu-rep captured with
`--import-source on` on a `-lineinfo` build:   python tools/ncu_source_hot.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    rows, hdr = [], None
    for row in csv.reader(io.StringIO(txt)):
        if len(row) > 10 and row[0] == "Line No":
            hdr = row
            continue
        if hdr and len(row) == len(hdr) and row[0].isdigit():
            try:
                rows.append((int(row[0]), row[1], int(row[hdr.index("Instructions Executed")]), int(row[hdr.index("# Samples")])))
            except ValueError:
                pass
    tot = sum(x[2] for x in rows) or 1
    ts = sum(x[3] for x in rows) or 1
    print("total warp instructions %d, stall samples %d" % (tot, ts))
    rows.sort(key=lambda x: -x[2])
    for ln, src, inst, samp in rows[:top]:
        print("%5d %6.2f%% inst %6.2f%% samp | %s" % (ln, 100.0 * inst / tot, 100.0 * samp / ts, src.strip()[:120]))


if __name__ == "__main__":
    main()
