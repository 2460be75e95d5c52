#!/usr/bin/env python
"""tools/sass_excerpt.py -- static SASS evidence for profiles/: per kernel, the opcode histogram and the lines around the copy-engine
instructions (UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk, SYNCS = mbarrier), from `cuobjdump -sass` of the built library.

    python tools/sass_excerpt.py <kernel-name-regex> [...] > profiles/r02_sass_excerpts.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "openmp-parallel-computing_b200" / "lib" / "libompimg_b200.so"
KEY = re.compile(r"\b(UTMALDG|UBLKCP|SYNCS|ELECT|UTMAPF|UTMASTG)\b")


def main():
    pats = [re.compile(p) for p in sys.argv[1:]]
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    names = {}
    dem = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    for m, d in zip(re.findall(r"Function : (\S+)", sass), dem):
        names[m] = d
    blocks = re.split(r"\n\s*Function : ", sass)[1:]
    print(f"# cuobjdump -sass {LIB.relative_to(ROOT)} (sm_100a), kernels matching {sys.argv[1:]}")
    for blk in blocks:
        mangled = blk.split("\n", 1)[0].strip()
        name = names.get(mangled, mangled)
        if not any(p.search(name) for p in pats):
            continue
        lines = [l for l in blk.splitlines() if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
        ops = collections.Counter()
        for l in lines:
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops[m.group(1).split(".")[0]] += 1
        print(f"\n== {name}\n   {len(lines)} instructions; static opcode counts: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
        shown = set()
        for i, l in enumerate(lines):
            if KEY.search(l) and ("UTMALDG" in l or "UBLKCP" in l or "UTMASTG" in l):
                for j in range(max(0, i - 3), min(len(lines), i + 3)):
                    if j not in shown:
                        shown.add(j)
                        print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", lines[j]).strip()[:150])
                print("   ...")
            if len(shown) > 60:
                print("   (further copy-engine sites omitted)")
                break


if __name__ == "__main__":
    main()
