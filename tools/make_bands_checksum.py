#!/usr/bin/env python
"""tools/make_bands_checksum.py -- derive, ON THE CPU ORACLE ONLY, the known answers bench.py and the
tests use for BASELINE config 5 (fused gray->Sobel of one synthetic 32768x32768 RGB image):

  * the sum of all interior edge bytes of the whole image (a single 64-bit number), and
  * per-row sums of the edge rows either side of every 1/2/4/8-way band seam,
  * (--extras) for the row-band forms of the old/ kernels on the same image: the three channel sums
    (old/parallel_avg_pixel.c:17-30) and the sum of all output bytes of the 5x5 box blur, the image taken as the
    pre-padded input (old/parallel_convolution.c:12-23; oracle_convolution_u8),

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
    ap.add_argument("--extras", action="store_true", help="only add channel_sums and conv5_box_sum to the existing file")
    a = ap.parse_args()
    W, H = a.width, a.height
    row = W * 3
    if a.extras:
        K = 5
        kern = np.ones((K, K), np.int64)
        sums = np.zeros(3, np.int64)
        conv_sum = 0
        out_h = H - K + 1
        y = 0
        while y < H:
            y1 = min(y + a.slab, H)
            src = O.synth((y1 - y) * row, SEED, y * row).reshape(y1 - y, W, 3)
            sums += src.reshape(-1, 3).astype(np.int64).sum(axis=0)
            y = y1
        y = 0
        while y < out_h:
            y1 = min(y + a.slab, out_h)     # output rows [y, y1) need input rows [y, y1+K-1)
            src = O.synth((y1 - y + K - 1) * row, SEED, y * row).reshape(y1 - y + K - 1, W, 3)
            conv_sum += int(O.convolution_u8(src, kern, K * K).astype(np.uint8).astype(np.int64).sum())
            y = y1
            print(f"\r{y}/{out_h}", end="", file=sys.stderr)
        print(file=sys.stderr)
        d = json.loads(Path(a.out).read_text())
        assert d["width"] == W and d["height"] == H and d["seed"] == SEED
        d["channel_sums"] = [int(v) for v in sums]
        d["conv5_box_sum"] = conv_sum
        Path(a.out).write_text(json.dumps(d, indent=1) + "\n")
        print(json.dumps({"channel_sums": d["channel_sums"], "conv5_box_sum": conv_sum}))
        return
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
