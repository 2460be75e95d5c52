#!/usr/bin/env python
"""tools/make_conv_checksums.py -- derive, ON THE CPU ORACLE ONLY, the known answers bench.py uses for BASELINE config 4
(5x5 and 9x9 box / binomial blur of a synthetic 8192x8192 RGB image, old/parallel_convolution.c:8-24 semantics):
the sum of all output bytes, for the two device layouts the bench times (rows 128-byte aligned; the dense (cols+K-1)*3
pitch).  The input is the bench's: the whole pitched buffer filled by synth(seed 4), so the two layouts hold different
pixels.  Walks the image in slabs of rows; no GPU, no product code.  ~10 minutes on 8 cores.

    python tools/make_conv_checksums.py  ->  tests/golden/conv_cfg4.json
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as O  # noqa: E402

SEED, ROWS, COLS, SLAB = 4, 8192, 8192, 1024


def kernels(K):
    b = np.array([1.0])
    for _ in range(K - 1):
        b = np.convolve(b, [1, 1])
    return {"box": np.ones((K, K), np.int64), "binomial": np.outer(b, b).astype(np.int64)}


def main():
    O.build()
    out = {"workload": f"cfg4: {ROWS}x{COLS} RGB, input = synth(seed {SEED}) over the whole pitched buffer", "seed": SEED, "rows": ROWS, "cols": COLS,
           "sums": {}, "source": "tools/make_conv_checksums.py (CPU oracle only)"}
    for K in (5, 9):
        pc = COLS + K - 1
        for layout, pitch in (("aligned128", (pc * 3 + 127) // 128 * 128), ("dense", pc * 3)):
            for name, kern in kernels(K).items():
                norm, total, y = int(kern.sum()), 0, 0
                while y < ROWS:
                    y1 = min(y + SLAB, ROWS)            # output rows [y, y1) need input rows [y, y1+K-1)
                    buf = O.synth((y1 - y + K - 1) * pitch, SEED, y * pitch).reshape(y1 - y + K - 1, pitch)
                    padded = np.ascontiguousarray(buf[:, :pc * 3]).reshape(y1 - y + K - 1, pc, 3)
                    total += int(O.convolution_u8(padded, kern, norm).astype(np.uint8).astype(np.int64).sum())
                    y = y1
                out["sums"][f"{name}{K}x{K}_{layout}"] = {"in_pitch": pitch, "out_byte_sum": total}
                print(name, K, layout, total, file=sys.stderr)
    (ROOT / "tests" / "golden" / "conv_cfg4.json").write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
