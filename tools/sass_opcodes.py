"""Per-kernel counts of the Blackwell-specific SASS opcodes in the built library (evidence that the contraction kernels
are tcgen05 / TMEM / TMA code, not recompiled mma.sync): UTCHMMA (tcgen05.mma, `.2CTA` = cta_group::2), UTMALDG / UTMASTG
(TMA load / store), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), and the legacy HMMA for contrast.
    python tools/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "edgestyle_b200", "libedgestyle_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "MUFU.EX2", "HMMA", "BRA.U.ANY"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(CUtensorMap_st.*|\(es::.*|\(int.*|\(float.*|\(.*", "", name)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for key in OPS:
            if key == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    counts[cur][key] += 1
            elif key == "UTCHMMA":
                if op.startswith("UTCHMMA") and ".2CTA" not in op:
                    counts[cur][key] += 1
            elif op.startswith(key):
                counts[cur][key] += 1
print(f"# {os.path.relpath(LIB, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass; only kernels with at least one listed opcode)")
print("| kernel | " + " | ".join(OPS) + " |")
print("|---|" + "---|" * len(OPS))
tot = collections.Counter()
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    tot.update(c)
    print(f"| {k} | " + " | ".join(str(c.get(o, 0)) for o in OPS) + " |")
print("| **total** | " + " | ".join(str(tot.get(o, 0)) for o in OPS) + " |")
