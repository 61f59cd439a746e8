"""Per-kernel SASS opcode counts of libcmrl_b200.so (what proves a Blackwell-native kernel: B200_PROFILING.md):
UTCHMMA (tcgen05.mma; .2CTA = cta_group::2), LDTM / STTM (tcgen05.ld / st), UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), MUFU, HMMA (legacy mma.sync: must be 0).  Usage: python profiles/tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = os.path.join(ROOT, "causal-mbrl_b200", "libcmrl_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
ops = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "MUFU", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS", "ATOM"]
rows, cur, counts, n_instr = [], None, None, 0
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, counts, n_instr))
        cur, counts, n_instr = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        n_instr += 1
        op = m.group(1)
        for o in ops:
            if o == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    counts[o] += 1
            elif op.split(".")[0] == o:
                counts[o] += 1
if cur:
    rows.append((cur, counts, n_instr))
demangle = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass causal-mbrl_b200/libcmrl_b200.so (sm_100a): opcode counts per kernel")
print("%-100s %7s " % ("kernel", "instrs") + " ".join("%7s" % o.replace("UTCHMMA.2CTA", ".2CTA") for o in ops))
for (name, c, n), dm in sorted(zip(rows, demangle), key=lambda x: -x[0][1]["UTCHMMA"]):
    short = re.sub(r"\(.*", "", dm)[:100]
    print("%-100s %7d " % (short, n) + " ".join("%7d" % c[o] for o in ops))
tot = collections.Counter()
for _, c, _ in rows:
    tot.update(c)
print("%-100s %7s " % ("TOTAL", "") + " ".join("%7d" % tot[o] for o in ops))
