#!/usr/bin/env python
"""tests/golden/make_jpeg_golden.py -- small JPEG fixtures for the decode step and their known answers.

The JPEGs are encoded here with Pillow from synthetic pixels (sizes around the 8/16-pixel block and MCU boundaries, the
sampling factors stb handles with dedicated resamplers, baseline and progressive, optimised Huffman tables, restart
intervals, grayscale); the known answer of each is the sha256 of what the REFERENCE's own decoder makes of it --
stbi_load(path, &w, &h, &ch, 0) from the vendored stb_image.h compiled as it lies (oracle/_ref/bin/stb_decode,
`make -C oracle ref`; monolithic/src/main.c:21).  Run in the build container only (needs /root/reference); commits
tests/golden/jpeg/*.jpg and tests/golden/jpeg/expected.json.
"""
import hashlib
import io
import json
import sys
from pathlib import Path

import numpy as np
from PIL import Image, ImageFile

ImageFile.MAXBLOCK = 1 << 22          # tiny progressive / optimised images otherwise overflow Pillow's encoder buffer

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as O  # noqa: E402

OUT = ROOT / "tests" / "golden" / "jpeg"


def picture(w, h, seed, gray=False):
    """smooth gradients + edges + noise: exercises DC prediction, long zero runs and dense blocks alike"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x + y) * 255 // max(w + h - 2, 1))], -1).astype(np.int32)
    base[(x // 5 + y // 7) % 2 == 0] += 60
    base += rng.integers(-40, 41, base.shape)
    img = np.clip(base, 0, 255).astype(np.uint8)
    return img[:, :, 0] if gray else img


def main():
    O.build()
    OUT.mkdir(parents=True, exist_ok=True)
    cases = []
    sizes = [(1, 1), (7, 5), (8, 8), (16, 16), (17, 9), (33, 31), (37, 29), (64, 48), (130, 70), (255, 257)]
    seed = 0
    for (w, h) in sizes:
        for sub, subname in ((0, "444"), (1, "422"), (2, "420")):
            for prog in (False, True):
                if (w, h) in ((8, 8), (16, 16), (64, 48)) and sub == 1:
                    continue                                     # keep the set small
                seed += 1
                name = f"s{w}x{h}_{subname}_{'prog' if prog else 'base'}"
                kw = dict(quality=[35, 75, 92, 100][seed % 4], subsampling=sub, progressive=prog, optimize=bool(seed % 3 == 0))
                cases.append((name, picture(w, h, seed), kw))
    cases.append(("gray_40x30_base", picture(40, 30, 101, gray=True), dict(quality=80)))
    cases.append(("gray_41x33_prog", picture(41, 33, 102, gray=True), dict(quality=60, progressive=True)))
    cases.append(("rst_100x60_420_base", picture(100, 60, 103), dict(quality=70, subsampling=2, restart_marker_blocks=3)))
    cases.append(("rst_100x60_444_prog", picture(100, 60, 104), dict(quality=85, subsampling=0, progressive=True, restart_marker_rows=1)))
    cases.append(("q10_120x90_420_prog", picture(120, 90, 105), dict(quality=10, subsampling=2, progressive=True)))
    expected = {}
    for name, img, kw in cases:
        buf = io.BytesIO()
        try:
            Image.fromarray(img).save(buf, "JPEG", **kw)
        except TypeError:
            kw = {k: v for k, v in kw.items() if not k.startswith("restart")}
            Image.fromarray(img).save(buf, "JPEG", **kw)
        path = OUT / f"{name}.jpg"
        path.write_bytes(buf.getvalue())
        dec = O.stb_decode(path)
        assert dec is not None, "oracle/_ref/bin/stb_decode is missing: make -C oracle ref"
        expected[name] = {"shape": list(dec.shape), "sha256": hashlib.sha256(dec.tobytes()).hexdigest(), "bytes": len(buf.getvalue()), "encode": {k: (v if not isinstance(v, bool) else int(v)) for k, v in kw.items()}}
    # the reference's own three images (not committed here: they travel in oracle/_ref/fixtures)
    for f in ("test.jpg", "more_than_one_mega_photo.jpg", "half_of_a_mega_photo.jpg"):
        dec = O.stb_decode(Path("/root/reference/images") / f)
        expected["ref:" + f] = {"shape": list(dec.shape), "sha256": hashlib.sha256(dec.tobytes()).hexdigest()}
    (OUT / "expected.json").write_text(json.dumps(expected, indent=1, sort_keys=True) + "\n")
    print(len(cases), "fixtures,", sum(v.get("bytes", 0) for v in expected.values()), "bytes")


if __name__ == "__main__":
    main()
