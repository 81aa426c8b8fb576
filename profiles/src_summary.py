"""Summarise an `ncu --page source --csv` dump: stall reasons, per-opcode samples, shared-memory
wavefronts (ideal vs excessive = bank conflicts).  usage: python profiles/src_summary.py file.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[hi[0]]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for i, r in enumerate(rows) if i > hi[0] and len(r) >= len(hdr) and r[0] != "Address"]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot, op, exc, wf, n = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter(), 0

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    for r in data:
        s = num(r[ci["# Samples"]])
        n += s
        for k in stalls:
            tot[k] += num(r[ci[k]])
        toks = r[ci["Source"]].split()
        o = toks[1] if toks[0].startswith("@") else toks[0]
        op[o] += s
        exc[o] += num(r[ci["L1 Wavefronts Shared Excessive"]])
        wf[o] += num(r[ci["L1 Wavefronts Shared"]])
    print(rows[0][1][:140])
    print("samples", n)
    for k, v in tot.most_common(8):
        print(f"  {k:26s} {v:8d} {v / max(n, 1) * 100:5.1f}%")
    for o, v in op.most_common(10):
        print(f"  {o:18s} samples {v:7d} ({v / max(n, 1) * 100:4.1f}%)  smem wavefronts {wf[o]:10d}  excessive {exc[o]:10d}")


if __name__ == "__main__":
    main(sys.argv[1])
