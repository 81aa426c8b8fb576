"""Per-warp-role stall summary of the warp-specialised TC block kernel from an `ncu --page source --csv` dump.
Roles are delimited by the USETMAXREG instructions (setup | T-mix | A-mix | MMA+loader | epilogue).
usage: python profiles/role_summary.py src.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hdr = rows[hi[0]]; ci = {h: i for i, h in enumerate(hdr)}
end = hi[1] - 1 if len(hi) > 1 else len(rows)
data = [r for r in rows[hi[0] + 1:end] if len(r) >= len(hdr)]
def n(x):
    try: return int(x)
    except ValueError: return 0
marks = [i for i, r in enumerate(data) if 'USETMAXREG' in r[ci['Source']]]
bounds = [0] + marks + [len(data)]
names = ['setup', 'T-mix warps', 'A-mix warps', 'MMA+loader', 'epilogue']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print(rows[0][1][:150])
for k in range(len(bounds) - 1):
    a, b = bounds[k], bounds[k + 1]
    tot = sum(n(r[ci['# Samples']]) for r in data[a:b])
    c, ops = collections.Counter(), collections.Counter()
    for r in data[a:b]:
        for s in stalls: c[s] += n(r[ci[s]])
        t = r[ci['Source']].split(); ops[t[1] if t[0].startswith('@') else t[0]] += n(r[ci['# Samples']])
    print(f"{names[k] if k < len(names) else k:12s} samples {tot:6d}  stalls {dict(c.most_common(4))}  ops {dict(ops.most_common(5))}")
