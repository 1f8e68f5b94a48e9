"""SASS opcode histogram of the kernels in libsnesgpu.so (cuobjdump -sass): the Blackwell-specific mnemonics that prove the
tile machinery (UTMALDG = TMA tensor loads, SYNCS = mbarrier, FFMA2 / FMUL2 / FADD2 = packed f32x2) and the top opcodes per kernel.

usage: python scripts/sass_histogram.py [kernel-substring ...]  > profiles/<tag>_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

sys.path.insert(0, ".")
from snesimage_b200 import _build  # noqa: E402

want = sys.argv[1:] or ["k_score_v3", "k_assign_pyr", "k_assign_dither", "k_assign_prepare", "k_kmeans", "k_pool_argmin"]
txt = subprocess.run(["cuobjdump", "-sass", _build.SO], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kern = name if any(w in name for w in want) else None
        if kern:
            hist.setdefault(kern, collections.Counter())
        continue
    if kern:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            hist[kern][m.group(1)] += 1
print(f"# cuobjdump -sass {_build.SO.split('/')[-1]} (sm_100a), static instruction counts; scorer sources sha256 {_build.scorer_source_hash()[:16]}")
KEY = ["UTMALDG", "UTMAPF", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "MUFU", "ATOMG", "LDGSTS"]
for k, c in hist.items():
    tot = sum(c.values())
    print(f"\n{k}: {tot} instructions")
    print("  Blackwell / packed / memory: " + ", ".join(f"{op} {c[op]}" for op in KEY if c[op]))
    print("  top: " + ", ".join(f"{op} {n}" for op, n in c.most_common(14)))
