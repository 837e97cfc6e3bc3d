#!/usr/bin/env python3
"""Per-kernel SASS summary of the shipped library: instruction count and the memory / fp64 / sync mnemonics.
Usage: tools/sass_summary.py [liblfx.so] > profiles/<round>_sass_summary.md   (cuobjdump -sass, sm_100a only)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = re.compile(r"^(LDG|STG|LDS|STS|LDGSTS|UBLKCP|UTMALDG|UTMASTG|UTMAPF|SYNCS|BAR|ATOM|ATOMG|ATOMS|RED|SHFL|VOTE|MATCH|"
                   r"DADD|DMUL|DFMA|DSETP|MUFU|MEMBAR|FENCE|CCTL|ERRBAR|LDL|STL)")


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lidar_feature_extraction_b200", "liblfx.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = [0, collections.Counter()]
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur][0] += 1
            op = m.group(1)
            if WATCH.match(op):
                kernels[cur][1][op] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS of the shipped liblfx.so (sm_100a), per kernel: instruction count and the memory / fp64 / sync mnemonics\n")
    print("`cuobjdump -sass lidar_feature_extraction_b200/liblfx.so`, summarised by `tools/sass_summary.py`. The hot kernel's 32-byte point "
          "loads are `LDG.E.NA.ENL2.256.CONSTANT` (sm_100 only); the converter stages its tiles with the TMA unit's bulk copy (`UBLKCP.S.G` + "
          "`SYNCS.*TRANS64`). No shipped kernel uses tensor-map TMA: the `UTMALDG` landing variant and the `UTMAPF` prefetch variant of the sector "
          "kernel were built, measured and dropped (profiles/r02a_tma_per_warp.md, tools/experiments/). Spills (`LDL` / `STL`) are listed where present.\n")
    for (name, (n, ops)), nice in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*", "", nice)
        print(f"## `{short}` — {n} instructions")
        top = ops.most_common(14)
        special = [(k, v) for k, v in ops.items() if re.match(r"^(UBLKCP|UTMA|SYNCS|LDL|STL|ATOM|RED)", k) and (k, v) not in top]
        print(", ".join(f"{k} x{v}" for k, v in top + special) or "(none of the watched mnemonics)")
        print()


if __name__ == "__main__":
    main()
