#!/usr/bin/env python
"""Per-kernel SASS evidence for the Blackwell-native instructions (profiles/*_sass_summary.md): counts of UTCHMMA
(tcgen05.mma; `.2CTA` = cta_group::2), LDTM (tcgen05.ld), UTMALDG / UTMASTG (TMA loads / stores, `.IM2COL`), UTCBAR
(tcgen05.commit), HMMA (mma.sync), FFMA2 (packed fp32) in the shipped library, via `cuobjdump -sass`.
usage: python tools/sass_summary.py [lib.so] > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "fce_yolo_b200", "libfce_yolo_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"),
       ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.IM2COL", r"\bUTMALDG\.\w*IM2COL|UTMALDG.*IM2COL"), ("UTMASTG", r"\bUTMASTG"),
       ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"), ("FFMA2", r"\bFFMA2"), ("MUFU.TANH", r"MUFU\.TANH"),
       ("UCGABAR", r"\bUCGABAR"), ("SYNCS", r"\bSYNCS")]
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    for name, pat in PAT:
        if re.search(pat, line):
            counts[kern][name] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS instruction summary of `{os.path.basename(lib)}` (cuobjdump -sass, sm_100a)\n")
cols = [n for n, _ in PAT]
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---|" * len(cols))
tot = collections.Counter()
for (k, c), d in zip(counts.items(), demangled):
    if not any(c.values()):
        continue
    name = d.replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)\s*$", "", name).replace("fce::", "")
    print(f"| `{name}` | " + " | ".join(str(c.get(n, 0) or "") for n in cols) + " |")
    tot.update(c)
print("| **total** | " + " | ".join(str(tot.get(n, 0)) for n in cols) + " |")
