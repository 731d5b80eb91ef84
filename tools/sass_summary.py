#!/usr/bin/env python
"""SASS instruction counts per kernel of the built library (cuobjdump -sass; runs here, no GPU needed):

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dune_eigensolver_b200", "csrc", "libdune_eigensolver_b200.so")
OPS = ["DMMA", "DFMA", "UBLKCP", "SYNCS", "LDGSTS", "UTCMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    counts, order, cur, k = collections.defaultdict(collections.Counter), [], None, 0
    for line in sass.splitlines():
        mfn = re.search(r"Function : (\S+)", line)
        if mfn:
            cur = re.sub(r"\(.*", "", names[k])
            k += 1
            if cur not in order:
                order.append(cur)
            continue
        if cur is None:
            continue
        mop = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if mop:
            op = mop.group(1)
            for o in OPS:
                if op.startswith(o):
                    counts[cur][o] += 1
    print("# SASS instruction counts per kernel of libdune_eigensolver_b200.so (cuobjdump -sass, sm_100a), round 2; tools/sass_summary.py")
    print("# DMMA = FP64 tensor-core mma (mma.sync.m8n8k4.f64); UBLKCP = cp.async.bulk (TMA-engine bulk copy); SYNCS = mbarrier ops;")
    print("# LDGSTS = cp.async (global -> shared, no register staging). tcgen05 (UTCMMA) has no FP64 kind: 0 everywhere by design.")
    print("%-110s %5s %5s %6s %5s %6s %6s" % ("kernel", *OPS))
    for name in order:
        c = counts[name]
        print("%-110s %5d %5d %6d %5d %6d %6d" % (name[:110], *[c[o] for o in OPS]))


if __name__ == "__main__":
    main()
