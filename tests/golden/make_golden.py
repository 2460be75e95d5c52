#!/usr/bin/env python
"""Regenerates tests/golden/* from the UNMODIFIED reference (run in the build container only,
where /root/reference exists and `make -C oracle ref` has compiled it into oracle/_ref/).

Outputs
  tests/golden/golden.json          digests + known answers (committed)
  tests/golden/test_crop_rgb.png    512x256 crop of images/test.jpg as decoded by the reference's
                                    stb_image (lossless PNG container; committed)
  tests/golden/test_crop_gray.png   the same crop of the reference's golden images/test_gray.png
  tests/golden/photo_crop_rgb.png   384x320 crop of images/more_than_one_mega_photo.jpg (stb-decoded)
  tests/golden/photo_crop_edge_{strict,fma}.png   reference gray+Sobel pass on that crop
  oracle/_ref/fixtures/*.raw        full-size decoded inputs / reference outputs (NOT committed;
                                    they travel to the GPU box with the snapshot)
"""
import hashlib
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as O  # noqa: E402

REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"
FIX = ROOT / "oracle" / "_ref" / "fixtures"
BIN = ROOT / "oracle" / "_ref" / "bin"


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def md5_file(p) -> str:
    return hashlib.md5(Path(p).read_bytes()).hexdigest()


def stb_decode(path: Path, name: str) -> np.ndarray:
    raw = FIX / f"{name}.raw"
    out = subprocess.run([str(BIN / "stb_decode"), str(path), str(raw)], check=True, capture_output=True, text=True)
    w, h, ch = map(int, out.stdout.split())
    return np.fromfile(raw, np.uint8).reshape(h, w, ch)


def main():
    FIX.mkdir(parents=True, exist_ok=True)
    subprocess.run(["make", "-C", str(ROOT / "oracle"), "-j4", "all", "ref"], check=True, capture_output=True)
    G = {"generated_by": "tests/golden/make_golden.py", "reference": "PedemonteGiacomo/OpenMp-Parallel-Computing"}

    # ---- grayscale LUTs from the compiled reference (all 2^24 triples) ------------------
    G["lut_sha256"] = {v: sha(O.Ref(v).lut()) for v in ("strict", "mono", "native")}
    ls, lf = O.Ref("strict").lut(), O.Ref("mono").lut()
    d = np.nonzero(ls != lf)[0]
    G["lut_strict_vs_fma"] = {"n_diff": int(d.size), "first_triples": [int(x) for x in d[:16]],
                              "max_abs": int(np.abs(ls.astype(int) - lf.astype(int)).max())}

    # ---- config 1: images/test.jpg vs the reference's golden images/test_gray.png ---------
    rgb = stb_decode(REF / "images" / "test.jpg", "test_rgb")
    gold = np.asarray(Image.open(REF / "images" / "test_gray.png"))
    assert gold.shape == rgb.shape
    G["test_jpg"] = {"shape": list(rgb.shape), "rgb_sha256": sha(rgb), "golden_gray_sha256": sha(gold),
                     "golden_png_md5": md5_file(REF / "images" / "test_gray.png")}
    gold.tofile(FIX / "test_gray_golden.raw")
    for variant in ("strict", "native"):
        for passes in (1, 2, 3):
            png = FIX / f"test_gray_{variant}_p{passes}.png"
            subprocess.run([str(BIN / f"grayscale_{variant}"), str(REF / "images" / "test.jpg"), str(png), str(passes)],
                           check=True, capture_output=True)
            G["test_jpg"][f"{variant}_p{passes}_png_md5"] = md5_file(png)
            G["test_jpg"][f"{variant}_p{passes}_sha256"] = sha(np.asarray(Image.open(png)))
            png.unlink()
    y0, x0, ch_, cw = 400, 700, 256, 512
    Image.fromarray(rgb[y0:y0 + ch_, x0:x0 + cw]).save(OUT / "test_crop_rgb.png", optimize=True)
    Image.fromarray(gold[y0:y0 + ch_, x0:x0 + cw]).save(OUT / "test_crop_gray.png", optimize=True)
    G["test_jpg"]["crop"] = [y0, x0, ch_, cw]

    # ---- config 2: gray+Sobel driver on images/more_than_one_mega_photo.jpg ---------------
    photo = stb_decode(REF / "images" / "more_than_one_mega_photo.jpg", "photo_rgb")
    G["photo_jpg"] = {"shape": list(photo.shape), "rgb_sha256": sha(photo)}
    for variant in ("strict", "native"):
        for passes in (1, 2):
            png = FIX / f"photo_edge_{variant}_p{passes}.png"
            subprocess.run([str(BIN / f"sobel_{variant}"), str(REF / "images" / "more_than_one_mega_photo.jpg"),
                            str(png), str(passes)], check=True, capture_output=True)
            out = np.asarray(Image.open(png))
            G["photo_jpg"][f"{variant}_p{passes}_png_md5"] = md5_file(png)
            G["photo_jpg"][f"{variant}_p{passes}_sha256"] = sha(out)
            if passes == 1:
                out.tofile(FIX / f"photo_edge_{variant}_p1.raw")
            png.unlink()
    half = stb_decode(REF / "images" / "half_of_a_mega_photo.jpg", "half_rgb")   # 2037x1362: odd pitch
    G["half_jpg"] = {"shape": list(half.shape), "rgb_sha256": sha(half)}
    for variant, lib in (("strict", "strict"), ("fma", "mono")):
        img, _ = O.Ref(lib).gray_sobel_pass(half.copy())
        G["half_jpg"][f"{variant}_p1_sha256"] = sha(img)
    py0, px0, pch, pcw = 1200, 800, 320, 384
    crop = np.ascontiguousarray(photo[py0:py0 + pch, px0:px0 + pcw])
    Image.fromarray(crop).save(OUT / "photo_crop_rgb.png", optimize=True)
    for variant, lib in (("strict", "strict"), ("fma", "mono")):
        img, _ = O.Ref(lib).gray_sobel_pass(crop.copy())
        Image.fromarray(img).save(OUT / f"photo_crop_edge_{variant}.png", optimize=True)
    G["photo_jpg"]["crop"] = [py0, px0, pch, pcw]

    # ---- synthetic inputs through the compiled reference --------------------------------------
    syn = {}
    for (h, w, seed) in [(1, 1, 1), (2, 7, 2), (3, 3, 3), (3, 64, 4), (17, 3, 5), (64, 64, 6), (257, 333, 7),
                         (480, 2037, 8), (1080, 1920, 9), (2160, 3840, 10)]:
        rgb_s = O.synth(h * w * 3, 0xB2000000 + seed).reshape(h, w, 3)
        plane = O.synth(h * w, 0xB2100000 + seed).reshape(h, w)
        e = {"seed": seed}
        for variant, lib in (("strict", "strict"), ("fma", "mono")):
            e[f"gray_{variant}"] = sha(O.Ref(lib).gray(rgb_s.copy()))
            e[f"pass_{variant}"] = sha(O.Ref(lib).gray_sobel_pass(rgb_s.copy())[0])
        e["sobel_prefill_ab"] = sha(O.Ref("strict").sobel(plane, np.full((h, w), 0xAB, np.uint8)))
        rgba = O.synth(h * w * 4, 0xB2200000 + seed).reshape(h, w, 4)
        e["gray_rgba_strict"] = sha(O.Ref("strict").gray(rgba.copy()))
        syn[f"{h}x{w}"] = e
    G["synthetic"] = syn

    # ---- old/ kernels through the stub-utils.h build -----------------------------------------
    old = {}
    for (rows, cols, K, norm) in [(37, 53, 5, 273), (64, 96, 9, 81), (24, 31, 3, 16), (48, 40, 5, 25)]:
        R = O.RefOld(rows, cols, K, norm)
        seed = rows * 1000 + cols
        padded = O.synth((rows + K - 1) * (cols + K - 1) * 3, 0xB2300000 + seed).reshape(rows + K - 1, cols + K - 1, 3)
        rng = np.random.default_rng(seed)
        kern = rng.integers(-3, 8, (K, K)).astype(np.int64) if norm not in (25, 81) else np.ones((K, K), np.int64)
        out = R.convolution(padded.astype(np.int64), kern, np.zeros((rows, cols, 3), np.int64))
        img = np.ascontiguousarray(padded[:rows, :cols])
        avg = R.avg_pixel(img.astype(np.int64))
        g3, mm = R.gray_avg3(img.astype(np.int64))
        old[f"{rows}_{cols}_{K}_{norm}"] = {
            "seed": seed, "kernel": kern.flatten().tolist(), "conv_sha256_i64": sha(out),
            "conv_head": out.flatten()[:12].tolist(), "avg": avg.tolist(), "gray3_sha256_i64": sha(g3), "minmax": mm.tolist()}
    G["old"] = old

    (OUT / "golden.json").write_text(json.dumps(G, indent=1, sort_keys=True) + "\n")
    print(json.dumps({k: G[k] for k in ("lut_sha256", "test_jpg")}, indent=1))


if __name__ == "__main__":
    main()
