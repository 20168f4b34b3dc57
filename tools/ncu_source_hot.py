"""Hot spots of one kernel from `ncu -i rep --page source --csv` output: stall samples per opcode and the
top instructions. usage: ncu_source_hot.py <source.csv> [top]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hdr]
si, ie = H.index("Warp Stall Sampling (All Samples)"), H.index("Instructions Executed")
data = [r for r in rows[hdr + 1:] if len(r) > ie and r[si].isdigit()]
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "instructions", len(data))
op, cnt = collections.Counter(), collections.Counter()
for r in data:
    t = r[1].split()
    o = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
    op[o] += int(r[si]); cnt[o] += int(r[ie] or 0)
for o, v in op.most_common(14):
    print(f"{o:18s} samples {v:8d} {v / max(tot, 1):.3f}  executed {cnt[o]}")
for r in sorted(data, key=lambda r: -int(r[si]))[:top_n]:
    print(r[0][-5:], r[1].strip()[:100], r[si], r[ie])
