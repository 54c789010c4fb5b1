"""Aggregate an ncu source page (SASS level) by CUDA source line.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all file.o && nvdisasm -g -c file.sm_100a.cubin > file.sass
    python tools/ncu_by_line.py src.csv file.sass <kernel substring> [top]

Instruction k of the kernel in the ncu page is instruction k of the function in the nvdisasm listing (same cubin)."""
import csv
import re
import sys
from collections import defaultdict


def line_map(sass_path, kernel):
    lines = open(sass_path).read().splitlines()
    out, cur, inside = [], None, False
    for l in lines:
        if l.startswith("\t.section\t.text."):
            inside = kernel in l
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            out.append(cur)
    return out


def main():
    src_csv, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = rows[hi + 1:]
    lm = line_map(sass, kernel)
    if len(lm) != len(body):
        print("warning: %d SASS instructions in the listing, %d in the ncu page" % (len(lm), len(body)))
    col = {n: i for i, n in enumerate(hdr)}
    want = ["# Samples", "Instructions Executed", "L1 Tag Requests Global", "L1 Wavefronts Shared", "L2 Theoretical Sectors Global",
            "stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_barrier", "stall_not_selected", "stall_selected"]
    want = [w for w in want if w in col]
    agg = defaultdict(lambda: [0.0] * len(want))
    for k, r in enumerate(body):
        key = lm[k] if k < len(lm) else None
        for j, w in enumerate(want):
            try:
                agg[key][j] += float(r[col[w]] or 0)
            except ValueError:
                pass
    tot = [sum(v[j] for v in agg.values()) for j in range(len(want))]
    print("%-28s" % "line", " ".join("%12s" % w[:12] for w in want))
    print("%-28s" % "TOTAL", " ".join("%12.0f" % t for t in tot))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        name = "%s:%d" % key if key else "?"
        print("%-28s" % name, " ".join("%12.0f" % x for x in v))


if __name__ == "__main__":
    main()
