#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass), e.g.
    python tools/sass_stats.py ntt_pass_cluster2 msm_accumulate
prints, for every kernel whose name contains one of the patterns, the instruction count per opcode (top 16) and the
counts of the opcodes that prove the Blackwell/Hopper async paths (UBLKCP = cp.async.bulk, UTMALDG = TMA tensor load,
SYNCS = mbarrier, UCGABAR = cluster barrier, CCTL = cache control)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "halo2_gpu_specific_b200", "libb2pcs.so")


def kernels(lib=LIB):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, res = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            res[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", ln)
        if m and cur:
            op = m.group(1)
            if op == "IMAD" and ".WIDE" in m.group(2):
                op = "IMAD.WIDE"
            elif op == "IMAD" and ".HI" in m.group(2):
                op = "IMAD.HI"
            elif op == "IMAD" and (".MOV" in m.group(2) or ".SHL" in m.group(2) or ".IADD" in m.group(2)):
                op = "IMAD.MOV/IADD/SHL"
            res[cur][op] += 1
    return res


if __name__ == "__main__":
    pats = sys.argv[1:] or [""]
    for name, c in kernels().items():
        if not any(p in name for p in pats):
            continue
        total = sum(c.values())
        top = ", ".join(f"{k} {v}" for k, v in c.most_common(16))
        tells = {k: c[k] for k in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "CCTL", "STL", "LDL") if c[k]}
        print(f"{name}\n  total {total}: {top}\n  tells: {tells}")
