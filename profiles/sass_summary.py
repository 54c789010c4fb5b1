"""Counts the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA / bulk copies -> UTMALDG/UTMASTG/UBLKCP, mbarrier -> SYNCS, legacy mma.sync -> HMMA) per kernel of the in-tree
objects.      python profiles/sass_summary.py > profiles/r02_sass_summary.txt      (needs only cuobjdump, no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deep-turbulence_b200", "tmglow_b200", "lib")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FFMA", "LDG", "STG", "LDS", "STS"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
        return [o if o else n for o, n in zip(out, names)]
    except OSError:
        return names


rows = []
for obj in sorted(f for f in os.listdir(LIB) if f.endswith(".o")):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(LIB, obj)], capture_output=True, text=True).stdout
    cur, cnt = None, None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                rows.append((obj, cur, cnt))
            cur, cnt = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            cnt["total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    cnt[k] += 1
                    break
    if cur:
        rows.append((obj, cur, cnt))
names = demangle([r[1] for r in rows])
print("# SASS mnemonic counts per kernel (cuobjdump -sass of deep-turbulence_b200/tmglow_b200/lib/*.o, sm_100a)")
print("# %-18s %-92s %7s " % ("object", "kernel", "instr") + " ".join("%7s" % k for k in KEYS[:11]))
for (obj, _, cnt), name in zip(rows, names):
    name = re.sub(r"\(.*$", "", name.replace("void ", "").replace("tmg::", ""))
    if not any(cnt[k] for k in KEYS[:11]) and cnt["total"] < 800:
        continue
    print("%-20s %-92s %7d " % (obj, name[:92], cnt["total"]) + " ".join("%7d" % cnt[k] for k in KEYS[:11]))
tot = collections.Counter()
for _, _, cnt in rows:
    tot.update(cnt)
print("# totals: " + ", ".join("%s %d" % (k, tot[k]) for k in KEYS[:11]))
print("# UTMALDG/UTMASTG (cp.async.bulk.tensor) = %d: operands that need the fp32 -> fp16 hi/lo conversion cannot be produced by TMA; "
      "weights stream with UBLKCP (cp.async.bulk), and in flow_level_kernel the activation operands never leave the SM." % (tot["UTMALDG"] + tot["UTMASTG"]))
