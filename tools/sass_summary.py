#!/usr/bin/env python3
"""tools/sass_summary.py -- per-kernel SASS mnemonic counts of lib/libte_pool.so (cuobjdump -sass): which kernels carry TMA bulk
copies (UBLKCP), how much FP64 work (DFMA / DMUL / DADD), local-memory spills (LDL / STL), global and shared accesses, shuffles.
    python tools/sass_summary.py [regex] > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "target_estimation_b200", "lib", "libte_pool.so")
WATCH = ["UBLKCP", "UBLKPF", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "LDG", "STG", "LDS", "STS", "LDL", "STL", "SHFL", "BAR", "LDC", "ACQBULK"]


def main():
    pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else re.compile(r"kf_step|isolver_kernel|gather_estimates|rebuild_kernel|mb_apply")
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    name = None
    counts = collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[name]["total"] += 1
            if op in WATCH:
                counts[name][op] += 1
    print("# SASS mnemonic counts per kernel of lib/libte_pool.so (cuobjdump -sass, sm_100a); only the mnemonics listed; total = all instructions")
    print("# UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier ops, LDL / STL = local memory (spills / stack)")
    for k, c in counts.items():
        if not pat.search(k):
            continue
        short = re.sub(r"\(te::StepArgs\)|void |te::", "", k)
        print("%-78s total %6d  %s" % (short[:78], c["total"], "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])))


if __name__ == "__main__":
    main()
