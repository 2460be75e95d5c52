#!/usr/bin/env python
"""tools/make_bands_checksum.py -- derive, ON THE CPU ORACLE ONLY, the known answers bench.py and the
tests use for BASELINE config 5 (fused gray->Sobel of one synthetic 32768x32768 RGB image):

  * the sum of all interior edge bytes of the whole image (a single 64-bit number), and
  * per-row sums of the edge rows either side of every 1/2/4/8-way band seam,

written to tests/golden/bands_cfg5.json.  The image is walked in slabs of rows with a one-row overlap, so
it needs ~400 MB, not 4 GB.  No GPU, no product code: synth bytes from oracle_synth_fill, pixel math from
oracle_rgb_to_edge (monolithic/src/main_with_sobel.c:51-74 restated).  ~2 minutes on 8 cores.

    python tools/make_bands_checksum.py [--width 32768 --height 32768 --slab 1024]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as O  # noqa: E402

SEED = 0xB2000005


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=32768)
    ap.add_argument("--height", type=int, default=32768)
    ap.add_argument("--slab", type=int, default=1024)
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "bands_cfg5.json"))
    a = ap.parse_args()
    W, H = a.width, a.height
    row = W * 3
    seams = sorted({H * r // n for n in (2, 4, 8) for r in range(1, n)})
    watch = sorted({y for s in seams for y in (s - 2, s - 1, s, s + 1)} | {1, 2, H - 3, H - 2})
    total = 0
    row_sums = {}
    y = 1                                   # first interior row
    while y < H - 1:
        y1 = min(y + a.slab, H - 1)         # output rows [y, y1)
        src = O.synth((y1 - y + 2) * row, SEED, (y - 1) * row).reshape(y1 - y + 2, W, 3)
        e = O.rgb_to_edge(src)[1:-1, 1:-1].astype(np.int64)
        total += int(e.sum())
        for yy in watch:
            if y <= yy < y1:
                row_sums[str(yy)] = int(e[yy - y].sum())
        y = y1
        print(f"\r{y}/{H}", end="", file=sys.stderr)
    print(file=sys.stderr)
    out = {"workload": f"cfg5: fused gray->Sobel (strict luma) of synth({W}x{H}x3, seed 0x{SEED:08X})", "width": W, "height": H, "seed": SEED,
           "interior_sum": total, "row_sums": row_sums, "source": "tools/make_bands_checksum.py (CPU oracle only)"}
    Path(a.out).write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps({"interior_sum": total}))


if __name__ == "__main__":
    main()
