"""cuobjdump -sass of the built library -> per kernel: instruction count and the memory / synchronisation / warp-collective
mnemonics with their counts and first occurrence.  usage: python profiles/tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "voxelis_b200", "libvoxelis_b200.so")
KEEP = ("LDG", "STG", "LDS", "STS", "LDL", "STL", "ATOM", "RED", "MEMBAR", "ERRBAR", "CCTL", "MATCH", "VOTE", "SHFL", "REDUX",
        "BAR", "UBLKCP", "SYNCS", "UTC", "HMMA", "IMMA", "ACQBULK", "NANOSLEEP", "FENCE", "LDGSTS", "WARPSYNC")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), {"n": 0, "ops": collections.OrderedDict()})
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if not m or cur is None:
        continue
    cur["n"] += 1
    text = m.group(2).strip()
    toks = text.split()
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    if op.startswith(KEEP):
        e = cur["ops"].setdefault(op, [0, "/*%s*/ %s" % (m.group(1), text)])
        e[0] += 1
print("SASS evidence (cuobjdump -sass voxelis_b200/libvoxelis_b200.so, sm_100a cubin; built by voxelis_b200/build.py; this file is")
print("written by profiles/tools/sass_evidence.py).  Per kernel: instruction count, the memory / synchronisation / warp-collective")
print("mnemonics with their counts, and the first occurrence of each (address + operands).  Round 2, final build.  No tensor-core")
print("(UTC*MMA / HMMA) and no TMA (UBLKCP) instructions by design: nothing on this path is a contraction, and the bulk-copy feed of")
print("the plan kernel measured slower.")
for name, k in kernels.items():
    if "cub" in name:
        continue
    print("\n== %s   (%d instructions)" % (name, k["n"]))
    for op, (cnt, first) in sorted(k["ops"].items(), key=lambda kv: -kv[1][0]):
        print("   %5d  %-34s first: %s" % (cnt, op, first[:110]))
