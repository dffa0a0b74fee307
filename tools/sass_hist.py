"""Opcode histogram (warp instructions executed) from `ncu --page source --csv` output."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[h]
si, ci, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
cnt, smp = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    try:
        n = int(r[ci])
    except Exception:
        continue
    op = r[si].split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") else op[0]
    parts = o.split(".")
    key = parts[0]
    if key.startswith(("LD", "ST", "ATOM", "RED")):
        key = ".".join(parts[:1] + [p for p in parts[1:] if p in ("64", "128", "U8", "S8", "U16", "CONSTANT", "LU", "EF")])
    cnt[key] += n
    try:
        smp[key] += int(r[st])
    except Exception:
        pass
tot, stot = sum(cnt.values()), max(1, sum(smp.values()))
for k, v in cnt.most_common(40):
    print("%-22s %12d %5.1f%%  %9.1f/unit   stall-samples %5.1f%%" % (k, v, 100 * v / tot, v / div, 100 * smp[k] / stot))
print("total warp-instr %d, per unit %.1f" % (tot, tot / div))
