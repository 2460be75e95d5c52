#!/usr/bin/env python
"""tools/prof.py -- launch one kernel of the path a few times on a working set larger than L2, for ncu.

    python tools/prof.py fused|sobel|gray|pass|conv5|conv9|conv5g|conv9g|sums [--frames N] [--iters K]

Prints the CUDA-event time per launch (never quote a number taken under a profiler).
"""
import argparse
import importlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel")
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--rows", type=int, default=8192, help="rows of the square-ish cfg4 image (sums)")
    ap.add_argument("--pitch-align", type=int, default=128, help="conv*: input pitch rounded up to this many bytes (1 = the dense (cols+K-1)*3 pitch)")
    ap.add_argument("--natural", action="store_true", help="fused/sobel: frames tiled from the reference's photo instead of uniform noise")
    a = ap.parse_args()
    pkg = importlib.import_module("openmp-parallel-computing_b200")
    pkg.init(0)
    L = pkg.lib()
    W, H, n = a.width, a.height, a.frames
    px = W * H
    def natural(channels):
        """one frame tiled from oracle/_ref/fixtures/more_than_one_mega_photo.jpg (a natural image: few saturated edges)"""
        from PIL import Image
        img = np.asarray(Image.open(ROOT / "oracle" / "_ref" / "fixtures" / "more_than_one_mega_photo.jpg").convert("RGB"))
        reps = (-(-H // img.shape[0]), -(-W // img.shape[1]), 1)
        frame = np.ascontiguousarray(np.tile(img, reps)[:H, :W])
        if channels == 1:
            frame = np.ascontiguousarray((frame.astype(np.float32) @ np.array([0.299, 0.587, 0.114], np.float32)).astype(np.uint8))
        return frame

    if a.kernel == "fused":
        src, dst = pkg.DeviceBuffer(n * px * 3), pkg.DeviceBuffer(n * px)
        pkg.synth_dev(src, src.nbytes, 3); dst.fill(0)
        if a.natural:
            f = natural(3)
            for i in range(n):
                src.upload(f, offset=i * px * 3)
        fn = lambda: L.ompimg_rgb_to_edge_dev(src.ptr, W * 3, 3, dst.ptr, W, W, H, n, px * 3, px, 0)
        bpp = 4
    elif a.kernel == "sobel":
        src, dst = pkg.DeviceBuffer(n * px), pkg.DeviceBuffer(n * px)
        pkg.synth_dev(src, src.nbytes, 2); dst.fill(0)
        if a.natural:
            f = natural(1)
            for i in range(n):
                src.upload(f, offset=i * px)
        fn = lambda: L.ompimg_sobel_dev(src.ptr, W, dst.ptr, W, W, H, n, px, px, 0)
        bpp = 2
    elif a.kernel == "gray":
        src = pkg.DeviceBuffer(n * px * 3)
        pkg.synth_dev(src, src.nbytes, 1)
        fn = lambda: L.ompimg_gray_inplace_dev(src.ptr, W * 3, W, H, 3, n, px * 3, 0)
        bpp = 6
    elif a.kernel == "pass":
        src, dst = pkg.DeviceBuffer(n * px * 3), pkg.DeviceBuffer(n * px)
        pkg.synth_dev(src, src.nbytes, 3); dst.fill(0)
        fn = lambda: L.ompimg_gray_sobel_pass_dev(src.ptr, W * 3, 3, dst.ptr, W, W, H, n, px * 3, px, 0)
        bpp = 6
    elif a.kernel == "passes":                           # the driver's loop: 8 passes per call, one preparation kernel
        src, dst = pkg.DeviceBuffer(n * px * 3), pkg.DeviceBuffer(n * px)
        pkg.synth_dev(src, src.nbytes, 3); dst.fill(0)
        fn = lambda: L.ompimg_gray_sobel_passes_dev(src.ptr, W * 3, 3, dst.ptr, W, W, H, n, px * 3, px, 8, 0)
        bpp = 6 * 8
    elif a.kernel in ("conv5", "conv9", "conv5g", "conv9g"):
        K = int(a.kernel[4])
        rows = cols = 8192
        n, px = 1, rows * cols
        pr, pc = rows + K - 1, cols + K - 1
        ip = -(-pc * 3 // a.pitch_align) * a.pitch_align
        src, dst = pkg.DeviceBuffer(pr * ip), pkg.DeviceBuffer(rows * cols * 3)
        pkg.synth_dev(src, src.nbytes, 4)
        taps = np.ones(K * K, np.int32)
        if a.kernel.endswith("g"):                      # binomial taps
            b = np.array([1.0])
            for _ in range(K - 1):
                b = np.convolve(b, [1, 1])
            taps = np.ascontiguousarray(np.outer(b, b).reshape(-1), np.int32)
        norm = int(taps.sum())
        fn = lambda: L.ompimg_conv_u8_dev(src.ptr, ip, rows, cols, K, taps.ctypes.data, norm, dst.ptr, cols * 3, 0)
        bpp = 6
    elif a.kernel == "sums":
        rows, cols = a.rows, 8192
        n, px = 1, rows * cols
        src, dst = pkg.DeviceBuffer(px * 3), pkg.DeviceBuffer(64)
        pkg.synth_dev(src, src.nbytes, 5); dst.fill(0)
        fn = lambda: L.ompimg_channel_sums_dev(src.ptr, cols * 3, rows, cols, dst.ptr, 0)
        bpp = 3
    else:
        raise SystemExit("unknown kernel " + a.kernel)
    for _ in range(3):
        assert fn() == 0, L.ompimg_last_error()
    pkg.sync()
    t = pkg.Timer()
    t.start(0)
    for _ in range(a.iters):
        fn()
    t.stop(0)
    ms = t.ms() / a.iters
    print(f"{a.kernel}: {ms:.4f} ms/launch, {n * px / ms / 1e6:.1f} Gpx/s, {bpp * n * px / ms / 1e6:.1f} GB/s "
          f"(sqrt_exact={L.ompimg_sqrt_is_exact()})")


if __name__ == "__main__":
    main()
