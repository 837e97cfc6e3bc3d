#!/usr/bin/env python3
"""Per CUDA source line: stall-reason samples of an ncu report's source page, sorted by one reason.
Usage: tools/ncu_stalls.py report.ncu-rep [reason=stall_long_sb] [top_n=25]"""
import csv
import subprocess
import sys

REASONS = ["stall_long_sb", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_math", "stall_mio", "stall_lg",
           "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_dispatch"]


def main():
    rep = sys.argv[1]
    key = sys.argv[2] if len(sys.argv) > 2 else "stall_long_sb"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    idx = {r: hdr.index(r) for r in REASONS}
    si = hdr.index("# Samples")
    data, tot = [], {r: 0 for r in REASONS}
    stot = 0
    for r in rows[hi + 1:]:
        if not r or not r[0].isdigit():
            continue
        try:
            v = {k: int(r[i] or 0) for k, i in idx.items()}
            s = int(r[si] or 0)
        except (ValueError, IndexError):
            continue
        for k in v:
            tot[k] += v[k]
        stot += s
        data.append((v, s, int(r[0]), r[1]))
    print("samples", stot, {k: f"{100 * v / max(stot, 1):.1f}%" for k, v in tot.items()})
    data.sort(key=lambda d: -d[0][key])
    for v, s, line, src in data[:top]:
        print(f"{100 * v[key] / max(tot[key], 1):5.1f}% of {key} ({v[key]:6d}; all {s:6d})  L{line}: {src.strip()[:110]}")


if __name__ == "__main__":
    main()
