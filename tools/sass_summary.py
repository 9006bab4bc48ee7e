#!/usr/bin/env python
"""Per-kernel SASS opcode summary of libhvit_sm100.so (cuobjdump -sass): which kernels use the 5th-generation tensor
cores (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit), tensor memory (LDTM / STTM = tcgen05.ld / st), TMA
(UTMALDG load, UTMASTG store, UTMAREDG reduce-add, UTMAPF prefetch), mbarriers (SYNCS) and the SFU (MUFU).

  python tools/sass_summary.py [lib.so] > profiles/r2_sass_opcodes.csv
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    ROOT, "speech-enhancement-via-hybrid-vision-transformer-project_b200", "csrc", "libhvit_sm100.so")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "MUFU", "FFMA2", "HMMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    out = [n.replace("hvit::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "") for n in out]
    return [re.sub(r"\((?!bool\)).*$", "", n) for n in out]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    names = demangle(list(kernels))
    print("kernel,instructions," + ",".join(COLS))
    for name, (_, c) in sorted(zip(names, kernels.items())):
        print(f"\"{name}\",{c['_total']}," + ",".join(str(c[k]) for k in COLS))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"\"TOTAL ({len(kernels)} kernels)\",{tot['_total']}," + ",".join(str(tot[k]) for k in COLS))


if __name__ == "__main__":
    main()
