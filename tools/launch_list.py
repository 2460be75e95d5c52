#!/usr/bin/env python
"""tools/launch_list.py -- per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_list.py gpurun_out/x_launches.csv "<command that was profiled>" > profiles/x_launch_list.txt
"""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0] != "ID"]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        ns = float(r[14].replace(",", ""))
        tot[r[4]][0] += 1
        tot[r[4]][1] += ns / 1000.0
    total = sum(v[1] for v in tot.values())
    cmd = sys.argv[2] if len(sys.argv) > 2 else "?"
    print(f"# ncu launch list of `{cmd}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)")
    print("# kernel | launches | total us | share")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k} | {n} | {us:.1f} | {100 * us / total:.1f}%")
    print(f"# total {len(rows)} launches, {total:.1f} us")


if __name__ == "__main__":
    main()
