#!/usr/bin/env python3
"""Aggregate an ncu report's source page per CUDA source line: share of executed warp instructions and
of stall samples. Usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    ci = hdr.index("Instructions Executed")
    si = hdr.index("# Samples")
    data, tot, stot = [], 0, 0
    for r in rows[hi + 1:]:
        if len(r) <= ci or not r[0].isdigit():
            continue
        try:
            n, s = int(r[ci]), int(r[si])
        except ValueError:
            continue
        tot += n
        stot += s
        data.append((n, s, int(r[0]), r[1]))
    print(f"total warp instructions {tot}, samples {stot}")
    data.sort(reverse=True)
    for n, s, line, src in data[:top]:
        print(f"{100 * n / max(tot, 1):5.1f}% inst {100 * s / max(stot, 1):5.1f}% samp  L{line}: {src.strip()[:120]}")


if __name__ == "__main__":
    main()
