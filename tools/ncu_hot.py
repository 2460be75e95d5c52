#!/usr/bin/env python
"""tools/ncu_hot.py -- the SASS lines with the most warp-stall samples of an .ncu-rep (source page), with their top stall reason.

    python tools/ncu_hot.py gpurun_out/x.ncu-rep [N] [context]
"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, body = rows[h], [r for r in rows[h + 1:] if len(r) > 5]
    ix = {k: i for i, k in enumerate(hdr)}
    S, src, ex = ix["# Samples"], ix["Source"], ix["Instructions Executed"]
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    f = lambda v: float(v or 0)
    tot = sum(f(r[S]) for r in body)
    top = sorted(range(len(body)), key=lambda i: -f(body[i][S]))[:n]
    shown = set()
    for i in top:
        for j in range(max(0, i - ctx), min(len(body), i + ctx + 1)):
            if j in shown:
                continue
            shown.add(j)
            r = body[j]
            why = max(stalls, key=lambda k: f(r[ix[k]]))
            print(f"{'*' if j == i else ' '} {j:5d} {100 * f(r[S]) / tot:5.1f}% exec {r[ex]:>9} {why[6:]:<16} {r[src][:120]}")
        if ctx:
            print()


if __name__ == "__main__":
    main()
