"""CPU tests of the oracle (oracle/oracle.c): against the golden vectors generated from the
UNMODIFIED reference (tests/golden/, made by tests/golden/make_golden.py) and, where oracle/_ref
exists (the build container, and the GPU box via the snapshot), against the compiled reference."""
import numpy as np
import pytest

from common import SYNTH_SHAPES, fixture_raw, png, sha


def test_lut_matches_reference_digests(O, golden):
    assert sha(O.gray_lut(False)) == golden["lut_sha256"]["strict"]
    assert sha(O.gray_lut(True)) == golden["lut_sha256"]["mono"] == golden["lut_sha256"]["native"]


def test_lut_numpy_restatements_agree(O, golden):
    idx = np.arange(1 << 24, dtype=np.uint32)
    r, g, b = (idx >> 16).astype(np.uint8), ((idx >> 8) & 255).astype(np.uint8), (idx & 255).astype(np.uint8)
    ls, lf = O.np_luma_strict(r, g, b), O.np_luma_fma(r, g, b)
    assert np.array_equal(ls, O.gray_lut(False))
    assert np.array_equal(lf, O.gray_lut(True))
    d = np.nonzero(ls != lf)[0]
    assert d.size == golden["lut_strict_vs_fma"]["n_diff"] == 1001
    assert [int(x) for x in d[:16]] == golden["lut_strict_vs_fma"]["first_triples"]


def test_gray_golden_crop_is_the_reference_golden_png(O, golden):
    """images/test_gray.png (the reference's only golden vector) pins the strict mode."""
    rgb, gold = png("test_crop_rgb.png"), png("test_crop_gray.png")
    assert np.array_equal(O.gray_inplace(rgb.copy(), 3, fma=False), gold)
    assert not np.array_equal(O.gray_inplace(rgb.copy(), 3, fma=True), gold)


def test_gray_full_image_when_fixture_present(O, golden):
    rgb = fixture_raw("test_rgb", golden["test_jpg"]["shape"])
    if rgb is None:
        pytest.skip("oracle/_ref/fixtures not built")
    assert sha(rgb) == golden["test_jpg"]["rgb_sha256"]
    one = O.gray_inplace(rgb.copy(), 3)
    assert sha(one) == golden["test_jpg"]["golden_gray_sha256"] == golden["test_jpg"]["strict_p1_sha256"]
    two = O.gray_inplace(one.copy(), 3)
    assert sha(two) == golden["test_jpg"]["strict_p2_sha256"]           # not idempotent: pass 2 differs
    assert sha(O.gray_inplace(two.copy(), 3)) == golden["test_jpg"]["strict_p3_sha256"]
    f1 = O.gray_inplace(rgb.copy(), 3, fma=True)
    assert sha(f1) == golden["test_jpg"]["native_p1_sha256"]


@pytest.mark.parametrize("variant,fma", [("strict", False), ("fma", True)])
def test_gray_sobel_pass_crop(O, variant, fma):
    rgb, want = png("photo_crop_rgb.png"), png(f"photo_crop_edge_{variant}.png")
    got, _ = O.gray_sobel_pass(rgb.copy(), fma=fma)
    assert np.array_equal(got, want)


def test_gray_sobel_pass_full_photo(O, golden):
    rgb = fixture_raw("photo_rgb", golden["photo_jpg"]["shape"])
    if rgb is None:
        pytest.skip("oracle/_ref/fixtures not built")
    img, edge = O.gray_sobel_pass(rgb.copy())
    assert sha(img) == golden["photo_jpg"]["strict_p1_sha256"]
    img2, _ = O.gray_sobel_pass(img.copy(), edge)
    assert sha(img2) == golden["photo_jpg"]["strict_p2_sha256"]
    imgf, _ = O.gray_sobel_pass(rgb.copy(), fma=True)
    assert sha(imgf) == golden["photo_jpg"]["native_p1_sha256"]


@pytest.mark.parametrize("h,w,seed", SYNTH_SHAPES)
def test_synthetic_against_reference_digests(O, golden, h, w, seed):
    e = golden["synthetic"][f"{h}x{w}"]
    rgb = O.synth(h * w * 3, 0xB2000000 + seed).reshape(h, w, 3)
    assert np.array_equal(rgb.ravel(), O.np_synth(h * w * 3, 0xB2000000 + seed))
    plane = O.synth(h * w, 0xB2100000 + seed).reshape(h, w)
    rgba = O.synth(h * w * 4, 0xB2200000 + seed).reshape(h, w, 4)
    for variant, fma in (("strict", False), ("fma", True)):
        assert sha(O.gray_inplace(rgb.copy(), 3, fma)) == e[f"gray_{variant}"]
        assert sha(O.gray_sobel_pass(rgb.copy(), fma=fma)[0]) == e[f"pass_{variant}"]
    assert sha(O.gray_inplace(rgba.copy(), 4)) == e["gray_rgba_strict"]
    pre = np.full((h, w), 0xAB, np.uint8)
    assert sha(O.sobel(plane, pre.copy())) == e["sobel_prefill_ab"]
    assert sha(O.sobel(plane, pre.copy(), isqrt=True)) == e["sobel_prefill_ab"]
    assert sha(O.np_sobel(plane, pre.copy())) == e["sobel_prefill_ab"]
    # the fused statement equals the driver's four loops
    edge = O.rgb_to_edge(rgb)
    img, _ = O.gray_sobel_pass(rgb.copy())
    assert np.array_equal(img[:, :, 0], edge)


@pytest.mark.parametrize("key", ["37_53_5_273", "64_96_9_81", "24_31_3_16", "48_40_5_25"])
def test_old_kernels_against_reference_digests(O, golden, key):
    rows, cols, K, norm = map(int, key.split("_"))
    e = golden["old"][key]
    padded = O.synth((rows + K - 1) * (cols + K - 1) * 3, 0xB2300000 + e["seed"]).reshape(rows + K - 1, cols + K - 1, 3)
    kern = np.array(e["kernel"], np.int64).reshape(K, K)
    out = O.convolution_long(padded.astype(np.int64), kern, norm, np.zeros((rows, cols, 3), np.int64))
    assert sha(out) == e["conv_sha256_i64"] and out.ravel()[:12].tolist() == e["conv_head"]
    assert np.array_equal(O.convolution_u8(padded, kern, norm).astype(np.int64), out)
    img = np.ascontiguousarray(padded[:rows, :cols])
    assert O.avg_pixel_u8(img).tolist() == e["avg"] == O.avg_pixel_long(img.astype(np.int64)).tolist()
    g3, mm = O.gray_avg3_u8(img)
    assert sha(g3.astype(np.int64)) == e["gray3_sha256_i64"] and mm.tolist() == e["minmax"]
    # accumulate-onto-prior-contents semantics (parallel_convolution.c:17, parallel_avg_pixel.c:28-30)
    prior = np.full((rows, cols, 3), 7, np.int64)
    acc = O.convolution_long(padded.astype(np.int64), kern, norm, prior.copy())
    full = O.convolution_long(padded.astype(np.int64), kern, 1, np.zeros((rows, cols, 3), np.int64))
    assert np.array_equal(acc, np.trunc((full + 7) / norm).astype(np.int64))


def test_compiled_reference_when_present(O):
    """Direct comparison with the unmodified reference objects (oracle/_ref)."""
    if not O.have_ref("strict"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(5)
    for (h, w) in [(3, 3), (5, 17), (64, 64), (129, 257), (480, 641)]:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        plane = rng.integers(0, 256, (h, w), dtype=np.uint8)
        for variant, fma in (("strict", False), ("mono", True)):
            R = O.Ref(variant)
            assert np.array_equal(R.gray(rgb.copy()), O.gray_inplace(rgb.copy(), 3, fma))
            assert np.array_equal(R.gray_sobel_pass(rgb.copy())[0], O.gray_sobel_pass(rgb.copy(), fma=fma)[0])
            pre = np.full((h, w), 0x5A, np.uint8)
            assert np.array_equal(R.sobel(plane, pre.copy()), O.sobel(plane, pre.copy()))
    # extreme inputs: step edges give perfect squares, white gives 255
    step = np.zeros((32, 48), np.uint8); step[:, 24:] = 255
    assert np.array_equal(O.Ref("strict").sobel(step), O.sobel(step))
    assert np.array_equal(O.Ref("strict").lut(), O.gray_lut(False))
    assert np.array_equal(O.Ref("mono").lut(), O.gray_lut(True))


def test_degenerate_sizes_write_nothing(O):
    for (h, w) in [(1, 1), (2, 9), (9, 2), (1, 40)]:
        plane = np.arange(h * w, dtype=np.uint8).reshape(h, w)
        pre = np.full((h, w), 0xEE, np.uint8)
        assert np.array_equal(O.sobel(plane, pre.copy()), pre)
