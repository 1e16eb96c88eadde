"""SASS opcode census of the shipped library: which kernels use tcgen05 (UTCHMMA), TMEM loads (LDTM), bulk
copies (UBLKCP), tensor-map TMA (UTMALDG), legacy tensor-core MMA (HMMA), 128-bit global loads.

    python tools/sass_census.py > profiles/r02_sass_census.md
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "raglite_b200" / "lib" / "libraglite_b200.so"
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "UBLKCP", "UTMALDG", "UTMAPF", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDG.E.128", "LDGSTS", "ATOMG", "REDG", "RED.E"]

sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()  # noqa: E731
per = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        cur = cur.replace("(anonymous namespace)::", "").replace("rl::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        per[cur] = Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        per[cur]["_total"] += 1
        for o in OPS:
            if op.startswith(o):
                per[cur][o] += 1
print("# r02 SASS opcode census of `raglite_b200/lib/libraglite_b200.so` (`cuobjdump -sass`, sm_100a)\n")
print("Counts are static instruction counts per kernel (template instances listed separately).\n")
print("| kernel | instr | " + " | ".join(OPS) + " |")
print("|---|---|" + "---|" * len(OPS))
tot = Counter()
for name, c in per.items():
    print(f"| `{name}` | {c['_total']} | " + " | ".join(str(c[o]) if c[o] else "" for o in OPS) + " |")
    tot.update(c)
print(f"| **all** | {tot['_total']} | " + " | ".join(str(tot[o]) for o in OPS) + " |")
print("\nUTCHMMA = tcgen05.mma (fp16 kind), LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (1-D, TMA engine), "
      "UTMALDG = cp.async.bulk.tensor (tensor-map TMA), UTMAPF = cp.async.bulk.prefetch.tensor, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, "
      "HMMA = legacy mma.sync, LDGSTS = cp.async.")
