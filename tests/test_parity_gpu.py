"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI of
libompimg_b200.so, against the CPU oracle on the same inputs, against the golden vectors generated
from the unmodified reference, and -- at BASELINE.json's full sizes -- through size-independent
properties.  Everything is bit-exact: these are uint8/int32 results."""
import ctypes as C

import numpy as np
import pytest

from common import SYNTH_SHAPES, fixture_raw, png, sha

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ helpers
def to_dev_pitched(gpu, a2d: np.ndarray, pitch: int, pad_value: int = 0xCD):
    """rows x row_bytes numpy -> device buffer with the given pitch (padding pre-filled)."""
    rows, rb = a2d.shape
    buf = gpu.DeviceBuffer(pitch * rows + 64)
    buf.fill(pad_value)
    L = gpu.lib()
    a2d = np.ascontiguousarray(a2d)
    assert L.ompimg_copy2d_h2d(buf.ptr, pitch, a2d.ctypes.data, rb, rb, rows, 0) == 0
    gpu.sync(0)
    return buf


def from_dev_pitched(gpu, buf, rows: int, rb: int, pitch: int) -> np.ndarray:
    out = np.empty((rows, rb), np.uint8)
    assert gpu.lib().ompimg_copy2d_d2h(out.ctypes.data, rb, buf.ptr, pitch, rb, rows, 0) == 0
    gpu.sync(0)
    return out


def ck(gpu, rc):
    assert rc == 0, gpu.lib().ompimg_last_error().decode()


# ------------------------------------------------------------------------------------------------ arithmetic
@pytest.mark.parametrize("mode,key", [("strict", "strict"), ("fma", "mono")])
def test_luma_all_2_24_triples(gpu, O, golden, mode, key):
    lut = gpu.luma_lut(mode)
    assert sha(lut) == golden["lut_sha256"][key]
    assert np.array_equal(lut, O.gray_lut(mode == "fma"))


def test_isqrt_path_exhaustive(gpu):
    """sat_u8(trunc(sqrt.approx(m+0.5))) == min(255, isqrt(m)) for every reachable m = gx^2+gy^2 <= 2*1020^2."""
    n = 2 * 1020 * 1020 + 1
    got = gpu.isqrt_table(n)
    m = np.arange(n, dtype=np.int64)
    r = np.floor(np.sqrt(m.astype(np.float64))).astype(np.int64)
    r = np.where(r * r > m, r - 1, r)
    r = np.where((r + 1) * (r + 1) <= m, r + 1, r)
    want = np.minimum(r, 255).astype(np.uint8)
    assert np.array_equal(got, want)
    # without the +0.5 guard the result depends on the MUFU.SQRT unit; the library checks this same table at
    # init and only then drops the guard
    exact = gpu.lib().ompimg_sqrt_is_exact()
    assert exact in (0, 1)
    assert bool(exact) == bool(np.array_equal(gpu.isqrt_table(n, 0.0), want))


def test_sobel_with_the_sqrt_guard_forced(O):
    """OMPIMG_SQRT_GUARD=1 keeps the +0.5 guard kernels: same bits."""
    import os, subprocess, sys
    code = r'''
import importlib, sys, numpy as np
sys.path.insert(0, "oracle"); import oracle as O
pkg = importlib.import_module("openmp-parallel-computing_b200"); pkg.init(0)
assert pkg.lib().ompimg_sqrt_is_exact() == 0
rgb = O.synth(270 * 1000 * 3, 77).reshape(270, 1000, 3)
assert np.array_equal(pkg.gray_sobel_pass(rgb.copy())[0], O.gray_sobel_pass(rgb.copy())[0])
plane = O.synth(300 * 2000, 78).reshape(300, 2000)
step = np.zeros((64, 1024), np.uint8); step[:, 500:] = 200; step[30:, :] += 55
for p in (plane, step):
    assert np.array_equal(pkg.sobel_edge(p), O.sobel(p))
print("guard ok")
'''
    from common import ROOT
    r = subprocess.run([sys.executable, "-c", code], cwd=str(ROOT), env=dict(os.environ, OMPIMG_SQRT_GUARD="1"), capture_output=True, text=True)
    assert r.returncode == 0 and "guard ok" in r.stdout, r.stderr[-2000:]


def test_synth_matches_oracle(gpu, O):
    for n, seed, off in [(1, 1, 0), (17, 2, 3), (4096 + 5, 3, 0), (1 << 20, 0xB2000003, 12345), (999, 7, 1 << 33)]:
        buf = gpu.DeviceBuffer(n + 64)
        for dst_off in (0, 1, 7):
            gpu.synth_dev(buf, n, seed, off, dst_offset=dst_off)
            assert np.array_equal(buf.download(n, offset=dst_off), O.synth(n, seed, off))


# ------------------------------------------------------------------------------------------------ a1 grayscale
def test_gray_golden_crop(gpu):
    gpu.set_luma_mode("strict")
    rgb, gold = png("test_crop_rgb.png"), png("test_crop_gray.png")
    assert np.array_equal(gpu.convert_to_grayscale(rgb.copy()), gold)


def test_gray_full_test_jpg(gpu, golden):
    rgb = fixture_raw("test_rgb", golden["test_jpg"]["shape"])
    if rgb is None:
        pytest.skip("oracle/_ref/fixtures not built")
    gpu.set_luma_mode("strict")
    one = gpu.convert_to_grayscale(rgb.copy())
    assert sha(one) == golden["test_jpg"]["golden_gray_sha256"]
    two = gpu.convert_to_grayscale(one.copy())
    assert sha(two) == golden["test_jpg"]["strict_p2_sha256"]
    gpu.set_luma_mode("fma")
    try:
        assert sha(gpu.convert_to_grayscale(rgb.copy())) == golden["test_jpg"]["native_p1_sha256"]
    finally:
        gpu.set_luma_mode("strict")


@pytest.mark.parametrize("h,w,seed", SYNTH_SHAPES)
def test_gray_synthetic_shapes(gpu, O, golden, h, w, seed):
    e = golden["synthetic"][f"{h}x{w}"]
    rgb = O.synth(h * w * 3, 0xB2000000 + seed).reshape(h, w, 3)
    rgba = O.synth(h * w * 4, 0xB2200000 + seed).reshape(h, w, 4)
    for mode in ("strict", "fma"):
        gpu.set_luma_mode(mode)
        try:
            got = gpu.convert_to_grayscale(rgb.copy())
            assert sha(got) == e[f"gray_{mode}"]
            assert np.array_equal(got, O.gray_inplace(rgb.copy(), 3, mode == "fma"))
            got4 = gpu.convert_to_grayscale(rgba.copy())
            assert np.array_equal(got4, O.gray_inplace(rgba.copy(), 4, mode == "fma"))
            assert np.array_equal(got4[:, :, 3], rgba[:, :, 3])          # alpha untouched
            if mode == "strict":
                assert sha(got4) == e["gray_rgba_strict"]
        finally:
            gpu.set_luma_mode("strict")


def test_gray_rejects_undefined_channel_counts(gpu):
    for ch in (1, 2, 5):
        with pytest.raises(gpu.OmpimgError):
            gpu.convert_to_grayscale(np.zeros((4, 4, ch), np.uint8))


@pytest.mark.parametrize("ch", [3, 4])
@pytest.mark.parametrize("w,h,pitch_extra,nframes", [(640, 37, 0, 3), (333, 20, 9, 2), (1000, 8, 48, 4), (4096, 5, 128, 1), (16, 3, 0, 5)])
def test_gray_device_pitched_batches(gpu, O, ch, w, h, pitch_extra, nframes):
    """device-resident entry point: pitched rows, several frames, aligned (bulk) and unaligned (generic) layouts"""
    row = w * ch
    pitch = row + pitch_extra
    src = O.synth(nframes * h * row, 77 + w + ch).reshape(nframes * h, row)
    buf = to_dev_pitched(gpu, src, pitch)
    ck(gpu, gpu.lib().ompimg_gray_inplace_dev(buf.ptr, pitch, w, h, ch, nframes, pitch * h, 0))
    got = from_dev_pitched(gpu, buf, nframes * h, pitch, pitch)
    want = O.gray_inplace(src.copy().reshape(nframes, h, w, ch), ch).reshape(nframes * h, row)
    assert np.array_equal(got[:, :row], want)
    assert np.all(got[:, row:] == 0xCD)                                  # padding untouched


def test_rgb_to_luma_device(gpu, O):
    for (w, h, ch) in [(1920, 16, 3), (1001, 7, 3), (512, 9, 4), (37, 5, 4)]:
        rgb = O.synth(h * w * ch, 5 + w).reshape(h, w * ch)
        rp, lp = (w * ch + 127) // 128 * 128, (w + 127) // 128 * 128
        src = to_dev_pitched(gpu, rgb, rp)
        dst = gpu.DeviceBuffer(lp * h + 64)
        dst.fill(0xEE)
        ck(gpu, gpu.lib().ompimg_rgb_to_luma_dev(src.ptr, rp, ch, dst.ptr, lp, w, h, 1, 0, 0, 0))
        got = from_dev_pitched(gpu, dst, h, lp, lp)
        assert np.array_equal(got[:, :w], O.rgb_to_luma(rgb.reshape(h, w, ch)))
        assert np.all(got[:, w:] == 0xEE)


# ------------------------------------------------------------------------------------------------ a2 Sobel
@pytest.mark.parametrize("h,w,seed", SYNTH_SHAPES)
def test_sobel_synthetic_shapes_border_untouched(gpu, O, golden, h, w, seed):
    plane = O.synth(h * w, 0xB2100000 + seed).reshape(h, w)
    pre = np.full((h, w), 0xAB, np.uint8)
    got = gpu.sobel_edge(plane, pre.copy())
    assert sha(got) == golden["synthetic"][f"{h}x{w}"]["sobel_prefill_ab"]
    assert np.array_equal(got, O.sobel(plane, pre.copy()))


@pytest.mark.parametrize("h,w", [(1, 1), (2, 9), (9, 2), (1, 40), (3, 3), (3, 17), (17, 3), (4, 513), (5, 1025), (16, 4097), (7, 8200)])
def test_sobel_small_and_ragged(gpu, O, h, w):
    rng = np.random.default_rng(h * 10007 + w)
    plane = rng.integers(0, 256, (h, w), dtype=np.uint8)
    pre = rng.integers(0, 256, (h, w), dtype=np.uint8)
    assert np.array_equal(gpu.sobel_edge(plane, pre.copy()), O.sobel(plane, pre.copy()))


def test_sobel_extreme_patterns(gpu, O):
    h, w = 96, 1024 + 48
    yy, xx = np.mgrid[0:h, 0:w]
    pats = {
        "step_v": np.where(xx >= w // 2, 255, 0), "step_h": np.where(yy >= h // 2, 255, 0),
        "checker": ((xx + yy) & 1) * 255, "checker2": (((xx >> 1) + (yy >> 1)) & 1) * 255,
        "white": np.full((h, w), 255), "ramp": (xx + yy) & 255, "stripes": (xx % 3 == 0) * 255,
        "perfect_squares": np.where(xx >= w // 2, 60, 0) + np.where(yy >= h // 2, 45, 0),   # |g| = (3,4,5)-multiples
    }
    for name, p in pats.items():
        plane = np.ascontiguousarray(p.astype(np.uint8))
        assert np.array_equal(gpu.sobel_edge(plane), O.sobel(plane)), name


@pytest.mark.parametrize("w,h,nframes,pitch_extra", [(1920, 40, 3, 0), (2000, 33, 2, 48), (333, 50, 4, 50), (4608, 12, 2, 0), (9000, 9, 1, 16)])
def test_sobel_device_batches(gpu, O, w, h, nframes, pitch_extra):
    sp = (w + 15) // 16 * 16 + pitch_extra if pitch_extra % 16 == 0 else w + pitch_extra
    dp = sp
    src = O.synth(nframes * h * w, 31 + w).reshape(nframes * h, w)
    sbuf = to_dev_pitched(gpu, src, sp)
    dbuf = gpu.DeviceBuffer(dp * h * nframes + 64)
    dbuf.fill(0x77)
    ck(gpu, gpu.lib().ompimg_sobel_dev(sbuf.ptr, sp, dbuf.ptr, dp, w, h, nframes, sp * h, dp * h, 0))
    got = from_dev_pitched(gpu, dbuf, nframes * h, dp, dp)
    for f in range(nframes):
        want = O.sobel(np.ascontiguousarray(src[f * h:(f + 1) * h]), np.full((h, w), 0x77, np.uint8))
        assert np.array_equal(got[f * h:(f + 1) * h, :w], want), f
    assert np.all(got[:, w:] == 0x77)


# ------------------------------------------------------------------------------------------------ a3..a5 fused
@pytest.mark.parametrize("variant", ["strict", "fma"])
def test_gray_sobel_pass_golden_crop(gpu, variant):
    gpu.set_luma_mode(variant)
    try:
        rgb, want = png("photo_crop_rgb.png"), png(f"photo_crop_edge_{variant}.png")
        got, _ = gpu.gray_sobel_pass(rgb.copy())
        assert np.array_equal(got, want)
    finally:
        gpu.set_luma_mode("strict")


def test_gray_sobel_pass_full_photo_two_passes(gpu, golden):
    """BASELINE config 2: the Sobel driver on images/more_than_one_mega_photo.jpg, passes = 1 and 2."""
    rgb = fixture_raw("photo_rgb", golden["photo_jpg"]["shape"])
    if rgb is None:
        pytest.skip("oracle/_ref/fixtures not built")
    h, w, _ = rgb.shape
    edge = np.zeros((h, w), np.uint8)                 # the driver's malloc'ed scratch: zero pages in practice
    img, edge = gpu.gray_sobel_pass(rgb.copy(), edge)
    assert sha(img) == golden["photo_jpg"]["strict_p1_sha256"]
    img2, _ = gpu.gray_sobel_pass(img.copy(), edge)
    assert sha(img2) == golden["photo_jpg"]["strict_p2_sha256"]
    gpu.set_luma_mode("fma")
    try:
        assert sha(gpu.gray_sobel_pass(rgb.copy())[0]) == golden["photo_jpg"]["native_p1_sha256"]
    finally:
        gpu.set_luma_mode("strict")


def test_gray_sobel_pass_half_photo_odd_pitch(gpu, golden):
    rgb = fixture_raw("half_rgb", golden["half_jpg"]["shape"])      # 2037 wide: no 16-byte friendly pitch
    if rgb is None:
        pytest.skip("oracle/_ref/fixtures not built")
    assert sha(gpu.gray_sobel_pass(rgb.copy())[0]) == golden["half_jpg"]["strict_p1_sha256"]


@pytest.mark.parametrize("h,w,seed", SYNTH_SHAPES)
def test_gray_sobel_pass_synthetic_shapes(gpu, O, golden, h, w, seed):
    e = golden["synthetic"][f"{h}x{w}"]
    rgb = O.synth(h * w * 3, 0xB2000000 + seed).reshape(h, w, 3)
    for mode in ("strict", "fma"):
        gpu.set_luma_mode(mode)
        try:
            got, _ = gpu.gray_sobel_pass(rgb.copy())
            assert sha(got) == e[f"pass_{mode}"]
        finally:
            gpu.set_luma_mode("strict")
    # scratch plane border survives and is replicated (main_with_sobel.c:37, :65-73)
    edge = np.full((h, w), 0x3C, np.uint8)
    got, edge_out = gpu.gray_sobel_pass(rgb.copy(), edge)
    want_img, want_edge = O.gray_sobel_pass(rgb.copy(), np.full((h, w), 0x3C, np.uint8))
    assert np.array_equal(got, want_img) and np.array_equal(edge_out, want_edge)


@pytest.mark.parametrize("ch", [3, 4])
@pytest.mark.parametrize("w,h,nframes", [(3840, 24, 2), (1920, 30, 3), (2000, 19, 2), (4200, 10, 2), (8200, 7, 1), (100, 40, 3), (16, 5, 2)])
def test_rgb_to_edge_device_batches(gpu, O, ch, w, h, nframes):
    row = w * ch
    rp, ep = (row + 127) // 128 * 128, (w + 127) // 128 * 128
    src = O.synth(nframes * h * row, 900 + w + ch).reshape(nframes * h, row)
    sbuf = to_dev_pitched(gpu, src, rp)
    dbuf = gpu.DeviceBuffer(ep * h * nframes + 64)
    dbuf.fill(0x11)
    ck(gpu, gpu.lib().ompimg_rgb_to_edge_dev(sbuf.ptr, rp, ch, dbuf.ptr, ep, w, h, nframes, rp * h, ep * h, 0))
    got = from_dev_pitched(gpu, dbuf, nframes * h, ep, ep)
    for f in range(nframes):
        frame = np.ascontiguousarray(src[f * h:(f + 1) * h]).reshape(h, w, ch)
        want = O.rgb_to_edge(frame, np.full((h, w), 0x11, np.uint8))
        assert np.array_equal(got[f * h:(f + 1) * h, :w], want), f
    assert np.all(got[:, w:] == 0x11)


def test_rgb_to_edge_unaligned_generic_path(gpu, O):
    w, h, ch = 333, 21, 3
    src = O.synth(h * w * ch, 4242).reshape(h, w * ch)
    sbuf = to_dev_pitched(gpu, src, w * ch)                 # dense, pitch 999: not a multiple of 16
    dbuf = gpu.DeviceBuffer(w * h + 64)
    dbuf.fill(0)
    ck(gpu, gpu.lib().ompimg_rgb_to_edge_dev(sbuf.ptr, w * ch, ch, dbuf.ptr, w, w, h, 1, 0, 0, 0))
    assert np.array_equal(from_dev_pitched(gpu, dbuf, h, w, w), O.rgb_to_edge(src.reshape(h, w, ch)))


def test_rgb_to_edge_batch_host(gpu, O):
    """BASELINE config 3 shape (3840x2160 RGB frames), a small batch through the pipelined host API."""
    n, h, w = 5, 2160, 3840
    frames = O.synth(n * h * w * 3, 0xB2000003).reshape(n, h, w, 3)
    got = gpu.rgb_to_edge_batch(frames)
    for f in (0, n - 1):
        assert np.array_equal(got[f], O.rgb_to_edge(frames[f]))
    small = O.synth(11 * 50 * 70 * 3, 99).reshape(11, 50, 70, 3)      # more frames than streams, odd pitch
    got = gpu.rgb_to_edge_batch(small)
    for f in range(11):
        assert np.array_equal(got[f], O.rgb_to_edge(small[f]))


def test_row_bands_equal_single_pass(gpu, O):
    """Row-band decomposition (BASELINE config 5) is bit-identical to the whole-image result; the bands
    here live on one device, so the halo pointers are exercised without needing a second GPU."""
    w, h = 1500, 257
    rgb = O.synth(h * w * 3, 0xB2000005).reshape(h, w, 3)
    want = O.rgb_to_edge(rgb)
    for ndev in (1, 2, 3, 8):
        b = gpu.Bands(w, h, 3, [0] * ndev)
        b.upload(rgb)
        b.rgb_to_edge()
        got = b.download_edge()
        assert np.array_equal(got, want), ndev
        assert b.checksum() == int(want[1:-1, 1:-1].astype(np.int64).sum())
        b.close()
    b = gpu.Bands(w, h, 3, [0, 0])
    b.synth(0xB2000005)
    b.rgb_to_edge()
    assert np.array_equal(b.download_edge(), want)
    b.close()


@pytest.mark.parametrize("h,w,nframes", [(1080, 1920, 2), (257, 1500, 1), (64, 2037, 3), (131, 512, 1), (3, 3, 1), (5, 17, 2), (4, 16, 1),
                                         (262, 4112, 1), (133, 513, 1), (2160, 3840, 1), (40, 8191, 1)])
def test_gray_sobel_pass_dev_in_place_fused(gpu, O, h, w, nframes):
    """a5 on the device, in place (main_with_sobel.c:51-74): 3-channel images take the preparation kernel + ONE fused kernel that
    writes RGB; two passes (the second runs on the first one's output), zero / absent / non-zero scratch border, both luma
    modes, pitched rows whose padding must survive, widths that end inside a 16-pixel group."""
    L = gpu.lib()
    pitch, epitch = (w * 3 + 127) // 128 * 128, (w + 127) // 128 * 128
    rng = np.random.default_rng(h * 10007 + w)
    for mode, border in (("strict", None), ("strict", 0xAB), ("fma", 0)):
        gpu.set_luma_mode(mode)
        try:
            frames = rng.integers(0, 256, (nframes, h, w, 3), dtype=np.uint8)
            edge_host = None
            if border is not None:
                edge_host = np.zeros((nframes, h, w), np.uint8)
                if border:
                    edge_host[:, 0, :] = rng.integers(0, 256, (nframes, w)); edge_host[:, -1, :] = rng.integers(0, 256, (nframes, w))
                    edge_host[:, :, 0] = rng.integers(0, 256, (nframes, h)); edge_host[:, :, -1] = rng.integers(0, 256, (nframes, h))
                edge_host[:, 1:-1, 1:-1] = 0x5C                                # interior: must come back untouched on the fused path
            img = to_dev_pitched(gpu, frames.reshape(nframes * h, w * 3), pitch)
            ebuf = to_dev_pitched(gpu, edge_host.reshape(nframes * h, w), epitch) if edge_host is not None else None
            want = [f.copy() for f in frames]
            for p in range(2):
                n0 = gpu.launch_count()
                ck(gpu, L.ompimg_gray_sobel_pass_dev(img.ptr, pitch, 3, ebuf.ptr if ebuf else None, epitch, w, h, nframes, pitch * h, epitch * h, 0))
                gpu.sync(0)
                assert gpu.launch_count() - n0 == 2                            # preparation + the fused kernel
                raw = from_dev_pitched(gpu, img, nframes * h, pitch, pitch)
                for f in range(nframes):
                    e0 = edge_host[f].copy() if edge_host is not None else None
                    want[f], _ = O.gray_sobel_pass(want[f], e0, fma=(mode == "fma"))
                    got = raw[f * h:(f + 1) * h, :w * 3].reshape(h, w, 3)
                    assert np.array_equal(got, want[f]), (mode, border, p, f)
                assert (raw[:, w * 3:] == 0xCD).all(), "row padding was written"
            if ebuf is not None:
                assert np.array_equal(from_dev_pitched(gpu, ebuf, nframes * h, epitch, epitch)[:, :w].reshape(nframes, h, w), edge_host)
                ebuf.free()
            img.free()
        finally:
            gpu.set_luma_mode("strict")


def test_gray_sobel_pass_dev_other_layouts_take_two_kernels(gpu, O):
    """4 channels (alpha must survive) and rows that are not 16-byte friendly: fused gray->Sobel into the scratch plane + replicate"""
    L = gpu.lib()
    rng = np.random.default_rng(5)
    for (h, w, ch, pitch) in ((70, 333, 4, 333 * 4), (70, 333, 3, 333 * 3), (33, 640, 4, 640 * 4 + 128)):
        img_h = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
        edge_h = np.zeros((h, w), np.uint8); edge_h[0] = 7; edge_h[:, 0] = 9
        img = to_dev_pitched(gpu, img_h.reshape(h, w * ch), pitch)
        ebuf = to_dev_pitched(gpu, edge_h, w)
        ck(gpu, L.ompimg_gray_sobel_pass_dev(img.ptr, pitch, ch, ebuf.ptr, w, w, h, 1, 0, 0, 0))
        gpu.sync(0)
        want_img, want_edge = O.gray_sobel_pass(img_h.copy(), edge_h.copy())
        assert np.array_equal(from_dev_pitched(gpu, img, h, w * ch, pitch).reshape(h, w, ch), want_img)
        assert np.array_equal(from_dev_pitched(gpu, ebuf, h, w, w), want_edge)
        assert L.ompimg_gray_sobel_pass_dev(img.ptr, pitch, ch, None, w, w, h, 1, 0, 0, 0) == -2      # these layouts need the scratch plane
        img.free(); ebuf.free()


@pytest.mark.parametrize("h,w,nframes", [(1080, 1920, 2), (257, 1500, 1), (64, 2037, 2), (131, 512, 1), (40, 1024, 1), (3, 3, 1), (5, 17, 2), (4, 16, 1),
                                         (262, 4112, 1), (133, 513, 1), (700, 3840, 1), (19, 8191, 1)])
def test_gray_sobel_passes_dev_multi_pass(gpu, O, h, w, nframes):
    """the driver's loop on the device (main_with_sobel.c:51-74 x passes) through ompimg_gray_sobel_passes_dev: every pass but the last
    writes the next pass's seam data itself (one preparation kernel for the whole run).  1, 2, 3 and 6 passes against the oracle's
    repeated pass, absent / zero / non-zero scratch border, both luma modes, widths on and off the 512-pixel strip grid."""
    L = gpu.lib()
    pitch, epitch = (w * 3 + 127) // 128 * 128, (w + 127) // 128 * 128
    rng = np.random.default_rng(h * 7919 + w)
    for mode, border in (("strict", 0xAB), ("fma", None), ("strict", 0)):
        gpu.set_luma_mode(mode)
        try:
            frames = rng.integers(0, 256, (nframes, h, w, 3), dtype=np.uint8)
            edge_host = None
            if border is not None:
                edge_host = np.zeros((nframes, h, w), np.uint8)
                if border:
                    edge_host[:, 0, :] = rng.integers(0, 256, (nframes, w)); edge_host[:, -1, :] = rng.integers(0, 256, (nframes, w))
                    edge_host[:, :, 0] = rng.integers(0, 256, (nframes, h)); edge_host[:, :, -1] = rng.integers(0, 256, (nframes, h))
            for passes in (1, 2, 3, 6):
                want = frames.copy()
                for f in range(nframes):
                    e = edge_host[f].copy() if edge_host is not None else np.zeros((h, w), np.uint8)
                    for _ in range(passes):
                        want[f], e = O.gray_sobel_pass(want[f].copy(), e, fma=(mode == "fma"))
                img = to_dev_pitched(gpu, frames.reshape(nframes * h, w * 3), pitch)
                ebuf = to_dev_pitched(gpu, edge_host.reshape(nframes * h, w), epitch) if edge_host is not None else None
                ck(gpu, L.ompimg_gray_sobel_passes_dev(img.ptr, pitch, 3, ebuf.ptr if ebuf else None, epitch, w, h, nframes, pitch * h, epitch * h, passes, 0))
                gpu.sync(0)
                raw = from_dev_pitched(gpu, img, nframes * h, pitch, pitch)
                assert np.array_equal(raw[:, :w * 3].reshape(nframes, h, w, 3), want), (mode, border, passes)
                assert (raw[:, w * 3:] == 0xCD).all(), "row padding was written"
                img.free()
                if ebuf:
                    ebuf.free()
        finally:
            gpu.set_luma_mode("strict")


def test_row_bands_on_distinct_devices_equal_oracle(gpu, O):
    """the real peer path: every band on its own GPU, halo rows fetched by the kernel's bulk copies over NVLink
    (skips on a one-GPU box; `gpurun --gpus 2` and the 8-GPU scaling run exercise it)"""
    ndev = gpu.lib().ompimg_device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    for w, h in ((1500, 257), (8192, 1040)):
        rgb = O.synth(h * w * 3, 0xB2000005).reshape(h, w, 3)
        want = O.rgb_to_edge(rgb)
        for n in sorted({2, min(ndev, 4), min(ndev, 8)}):
            b = gpu.Bands(w, h, 3, list(range(n)))
            b.upload(rgb)
            b.rgb_to_edge()
            b.rgb_to_edge()
            assert np.array_equal(b.download_edge(), want), (w, h, n)
            assert b.checksum() == int(want[1:-1, 1:-1].astype(np.int64).sum())
            b.close()


def _check_band_conv_and_sums(gpu, O, w, h, devices):
    rgb = O.synth(h * w * 3, 0xB2000004 + len(devices)).reshape(h, w, 3)
    b = gpu.Bands(w, h, 3, devices)
    b.upload(rgb)
    try:
        sums, _ = b.channel_sums()
        assert sums == [int(v) for v in rgb.reshape(-1, 3).astype(np.int64).sum(axis=0)], devices
        # the reference's result from those sums: (avgs_in + 3*sum) / (rows*cols*3), old/parallel_avg_pixel.c:19-23,40
        assert [(3 * v) // (h * w * 3) for v in sums] == O.avg_pixel_u8(rgb).tolist(), devices
        for K in (5, 9, 4):
            bn = np.array([1.0])
            for _ in range(K - 1):
                bn = np.convolve(bn, [1, 1])
            for kern in (np.ones((K, K), np.int64), np.outer(bn, bn).astype(np.int64)):
                norm = int(kern.sum())
                b.convolution(kern, norm)
                b.convolution(kern, norm)                   # the halo rows copied by the first call must not disturb the second
                want = O.convolution_u8(rgb, kern, norm).astype(np.uint8)
                assert np.array_equal(b.download_convolution(), want), (w, h, devices, K)
                assert b.convolution_checksum() == int(want.astype(np.int64).sum())
        b.rgb_to_edge()                                     # the Sobel bands still see their own rows only
        assert np.array_equal(b.download_edge(), O.rgb_to_edge(rgb))
    finally:
        b.close()


def test_row_band_convolution_and_channel_sums(gpu, O):
    """Row-band forms of a6 and a7 (old/parallel_convolution.c:12-23, old/parallel_avg_pixel.c:17-30): K-1 halo rows copied from
    the next band(s), partial channel sums added on the host -- equal to the oracle on the whole image.  Bands here live on one
    device (including bands shorter than K-1 rows); the next test puts every band on its own GPU."""
    for (w, h, ndev) in ((700, 131, 3), (517, 64, 8), (1030, 40, 16), (300, 33, 1)):
        _check_band_conv_and_sums(gpu, O, w, h, [0] * ndev)


def test_row_band_convolution_on_distinct_devices(gpu, O):
    """the real P2P path: cudaMemcpyPeerAsync of the K-1 rows below a band from the neighbour GPU (skips on a one-GPU box)"""
    ndev = gpu.lib().ompimg_device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    for n in sorted({2, min(ndev, 4), min(ndev, 8)}):
        _check_band_conv_and_sums(gpu, O, 2100, 515, list(range(n)))


@pytest.mark.parametrize("w", [17, 511, 512, 513, 2047, 4096, 4097, 4112, 8191])
@pytest.mark.parametrize("h", [3, 4, 5, 6, 7, 17, 130, 131, 132, 133, 134, 262, 263, 264])
def test_sobel_and_fused_tricky_shapes(gpu, O, w, h):
    """shapes around the kernel's internal boundaries: 512-column warps, 4096-column strips, 16-pixel groups,
    three-row trips and the 130/64/34/16-row bands (first band one row taller)"""
    if w > 4200 and h > 140:
        pytest.skip("covered by the smaller combinations")
    L = gpu.lib()
    seed = w * 1000 + h
    for ch in (1, 3):
        src = O.synth(h * w * ch, seed + ch).reshape(h, w * ch)
        want = O.sobel(src, np.full((h, w), 0x5A, np.uint8)) if ch == 1 else O.rgb_to_edge(src.reshape(h, w, 3), np.full((h, w), 0x5A, np.uint8))
        sp, ep = (w * ch + 127) // 128 * 128, (w + 127) // 128 * 128
        sbuf = to_dev_pitched(gpu, src, sp)
        ebuf = gpu.DeviceBuffer(ep * h + 64)
        ebuf.fill(0x5A)
        if ch == 1:
            ck(gpu, L.ompimg_sobel_dev(sbuf.ptr, sp, ebuf.ptr, ep, w, h, 1, 0, 0, 0))
        else:
            ck(gpu, L.ompimg_rgb_to_edge_dev(sbuf.ptr, sp, 3, ebuf.ptr, ep, w, h, 1, 0, 0, 0))
        got = from_dev_pitched(gpu, ebuf, h, ep, ep)
        assert np.array_equal(got[:, :w], want), (w, h, ch)
        assert np.all(got[:, w:] == 0x5A)
        sbuf.free(); ebuf.free()


@pytest.mark.parametrize("ch", [1, 3])
def test_band_entry_point_with_explicit_halo_rows(gpu, O, ch):
    """ompimg_rgb_to_edge_band_dev: what a rank of the one-process-per-GPU row-band mode calls after it has
    received its neighbours' rows (bands_mp.py); three bands with halo buffers reproduce the whole image."""
    L = gpu.lib()
    w, h = 1100, 151
    img = O.synth(h * w * ch, 0xB2000055 + ch).reshape(h, w * ch)
    whole = O.rgb_to_edge(img.reshape(h, w, 3)) if ch == 3 else O.sobel(img)
    pitch, ep = (w * ch + 127) // 128 * 128, (w + 127) // 128 * 128
    got = np.zeros((h, w), np.uint8)
    bounds = [0, 47, 48, h]                                            # a one-row band in the middle as well
    for r0, r1 in zip(bounds, bounds[1:]):
        band = to_dev_pitched(gpu, img[r0:r1], pitch)
        above = to_dev_pitched(gpu, img[r0 - 1:r0], pitch) if r0 > 0 else None
        below = to_dev_pitched(gpu, img[r1:r1 + 1], pitch) if r1 < h else None
        edge = gpu.DeviceBuffer(ep * (r1 - r0) + 64)
        edge.fill(0)
        ck(gpu, L.ompimg_rgb_to_edge_band_dev(band.ptr, pitch, ch, edge.ptr, ep, w, r1 - r0,
                                             above.ptr if above else None, below.ptr if below else None, 0))
        gpu.sync(0)
        got[r0:r1] = from_dev_pitched(gpu, edge, r1 - r0, ep, ep)[:, :w]
    assert np.array_equal(got[1:-1, 1:-1], whole[1:-1, 1:-1])
    assert not got[0].any() and not got[-1].any() and not got[:, 0].any() and not got[:, -1].any()


# ------------------------------------------------------------------------------------------------ old/ kernels
@pytest.mark.parametrize("key", ["37_53_5_273", "64_96_9_81", "24_31_3_16", "48_40_5_25"])
def test_old_kernels_golden(gpu, O, golden, key):
    rows, cols, K, norm = map(int, key.split("_"))
    e = golden["old"][key]
    padded = O.synth((rows + K - 1) * (cols + K - 1) * 3, 0xB2300000 + e["seed"]).reshape(rows + K - 1, cols + K - 1, 3)
    kern = np.array(e["kernel"], np.int64).reshape(K, K)
    out = gpu.parallel_convolution(padded, kern, norm)
    assert sha(out.astype(np.int64)) == e["conv_sha256_i64"]
    img = np.ascontiguousarray(padded[:rows, :cols])
    assert gpu.parallel_avg_pixel(img).tolist() == e["avg"]
    g3, mm = gpu.parallel_to_grayscale(img)
    assert sha(g3.astype(np.int64)) == e["gray3_sha256_i64"] and mm.tolist() == e["minmax"]


@pytest.mark.parametrize("key", ["37_53_5_273", "24_31_3_16"])
def test_old_signatures_through_long_array_adapters(gpu, O, key, tmp_path):
    """The reference's old/ functions and tests/adapters/old_long_adapters.c (same `long`-array signatures, on the C ABI)
    called with the SAME arguments: identical `long` arrays -- including the accumulate-onto-the-caller's-array statement
    (old/parallel_convolution.c:17) with a non-zero array, and avgs that do not start at zero (old/parallel_avg_pixel.c:28-30)."""
    import ctypes as C
    import subprocess
    from common import ROOT
    rows, cols, K, norm = map(int, key.split("_"))
    so = tmp_path / f"adapters_{key}.so"
    lib_dir = gpu.LIB_PATH.parent
    r = subprocess.run(["gcc", "-O2", "-fPIC", "-shared", f"-I{ROOT / 'oracle' / 'stub'}", f"-I{ROOT / 'include'}",
                        f"-DDIM_ROW={rows}", f"-DDIM_COL={cols}", f"-DDIM_KERNEL={K}", f"-DGBLUR_NORM={norm}",
                        str(ROOT / "tests" / "adapters" / "old_long_adapters.c"), "-o", str(so), f"-L{lib_dir}", "-lompimg_b200", f"-Wl,-rpath,{lib_dir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ours = O.RefOld.__new__(O.RefOld)
    ours.rows, ours.cols, ours.K, ours.norm, ours.L = rows, cols, K, norm, C.CDLL(str(so))
    rng = np.random.default_rng(rows * 7 + K)
    padded = O.synth((rows + K - 1) * (cols + K - 1) * 3, 0xAD00 + K).reshape(rows + K - 1, cols + K - 1, 3).astype(np.int64)
    kern = rng.integers(-5, 9, (K, K)).astype(np.int64)
    prior = rng.integers(-100000, 100000, (rows, cols, 3)).astype(np.int64)
    want = O.convolution_long(padded, kern, norm, prior.copy())
    if O.have_ref_old(rows, cols, K, norm):
        assert np.array_equal(O.RefOld(rows, cols, K, norm).convolution(padded, kern, prior.copy()), want)
    assert np.array_equal(ours.convolution(padded, kern, prior.copy()), want)
    assert np.array_equal(ours.convolution(padded, kern, np.zeros_like(prior)), O.convolution_long(padded, kern, norm, np.zeros_like(prior)))
    img = np.ascontiguousarray(padded[:rows, :cols])
    assert ours.avg_pixel(img, (5, 1000003, -77)).tolist() == O.avg_pixel_long(img, (5, 1000003, -77)).tolist()
    g3, mm = ours.gray_avg3(img)
    w3, wmm = O.gray_avg3_long(img)
    assert np.array_equal(g3, w3) and mm.tolist() == wmm.tolist()


def test_convolution_accumulate_and_argument_checks(gpu, O):
    rows, cols, K = 70, 333, 5
    padded = O.synth((rows + K - 1) * (cols + K - 1) * 3, 77).reshape(rows + K - 1, cols + K - 1, 3)
    kern = np.outer([1, 4, 6, 4, 1], [1, 4, 6, 4, 1]).astype(np.int64)
    prior = np.random.default_rng(3).integers(-(1 << 20), 1 << 20, (rows, cols, 3)).astype(np.int32)
    want = O.convolution_long(padded.astype(np.int64), kern, 256, prior.astype(np.int64))
    assert np.array_equal(gpu.parallel_convolution(padded, kern, 256, out=prior.copy()), want)
    L = gpu.lib()
    src = to_dev_pitched(gpu, padded.reshape(rows + K - 1, -1), (cols + K - 1) * 3 + 1)
    dst = gpu.DeviceBuffer(rows * cols * 3 * 4)
    taps = np.ascontiguousarray(kern, np.int32)
    # uint8 output never clamps (neither does the reference): it is the low byte of the exact result, documented in the header
    assert L.ompimg_conv_u8_dev(src.ptr, (cols + K - 1) * 3 + 1, rows, cols, K, taps.ctypes.data, 100, dst.ptr, cols * 3, 0) == 0
    gpu.sync(0)
    assert np.array_equal(dst.download(rows * cols * 3).reshape(rows, cols, 3), O.convolution_u8(padded, kern, 100).astype(np.uint8))
    # pitches that cannot hold a row: refused instead of read/written out of bounds
    assert L.ompimg_conv_i32_dev(src.ptr, (cols + K - 1) * 3 - 1, rows, cols, K, taps.ctypes.data, 256, dst.ptr, cols * 12, 0) == -2
    assert L.ompimg_conv_i32_dev(src.ptr, (cols + K - 1) * 3 + 1, rows, cols, K, taps.ctypes.data, 256, dst.ptr, cols * 12 - 4, 0) == -2
    assert L.ompimg_channel_sums_dev(src.ptr, cols * 3 - 1, rows, cols, dst.ptr, 0) == -2
    for bad in (np.zeros((8, 8), np.uint8), np.zeros((8, 8, 4), np.uint8)):
        with pytest.raises(ValueError):
            gpu.parallel_avg_pixel(bad)
        with pytest.raises(ValueError):
            gpu.parallel_to_grayscale(bad)
        with pytest.raises(ValueError):
            gpu.parallel_convolution(bad, np.ones((3, 3)), 9)


def test_frozen_grayscale_leaves_one_and_two_channel_images_alone(gpu):
    """channels 1 / 2 are undefined behaviour in the reference (parallel_to_grayscale.c:10-14); the drop-in neither aborts
    (the services would answer 500) nor invents pixels: the buffer is left as it is."""
    import ctypes as C
    for ch in (1, 2):
        a = np.random.default_rng(ch).integers(0, 256, (16, 24, ch), dtype=np.uint8)
        b = a.copy()
        gpu.lib().convert_to_grayscale(C.c_void_p(b.ctypes.data), 24, 16, ch)
        assert np.array_equal(a, b)
        with pytest.raises(gpu.OmpimgError):
            gpu.convert_to_grayscale(b)                      # the non-frozen API still reports it


def test_host_pipelines_many_bands(gpu, O):
    """the row-band host pipeline with many bands, pageable and pinned buffers, odd widths (ring reuse, halo events)"""
    h, w = 1500, 2037
    rng = np.random.default_rng(11)
    plane = rng.integers(0, 256, (h, w), dtype=np.uint8)
    pre = np.full((h, w), 0x3C, np.uint8)
    assert np.array_equal(gpu.sobel_edge(plane, pre.copy()), O.sobel(plane, pre.copy()))
    rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    edge = np.zeros((h, w), np.uint8); edge[0, :] = 9; edge[:, -1] = 200
    want_img, want_edge = O.gray_sobel_pass(rgb.copy(), edge.copy())
    got_img, got_edge = gpu.gray_sobel_pass(rgb.copy(), edge.copy())
    assert np.array_equal(got_img, want_img) and np.array_equal(got_edge, want_edge)
    rgba = rng.integers(0, 256, (700, 1001, 4), dtype=np.uint8)
    assert np.array_equal(gpu.gray_sobel_pass(rgba.copy())[0], O.gray_sobel_pass(rgba.copy())[0])
    pin = gpu.PinnedBuffer(h * w * 3)
    pimg = pin.array.reshape(h, w, 3)
    pimg[:] = rgb
    assert np.array_equal(gpu.gray_sobel_pass(pimg)[0], O.gray_sobel_pass(rgb.copy())[0])
    pin.free()
    K = 9
    padded = rng.integers(0, 256, (900 + K - 1, 1111 + K - 1, 3), dtype=np.uint8)
    kern = rng.integers(-3, 4, (K, K)).astype(np.int64)
    assert np.array_equal(gpu.parallel_convolution(padded, kern, 7), O.convolution_u8(padded, kern, 7))
    big = rng.integers(0, 256, (3000, 1999, 3), dtype=np.uint8)
    assert gpu.parallel_avg_pixel(big, (3, 2, 1)).tolist() == O.avg_pixel_u8(big, (3, 2, 1)).tolist()
    g3, mm = gpu.parallel_to_grayscale(big)
    w3, wmm = O.gray_avg3_u8(big)
    assert np.array_equal(g3, w3) and mm.tolist() == wmm.tolist()


@pytest.mark.parametrize("K,kind", [(3, "rand"), (5, "box"), (5, "gauss"), (5, "rand"), (7, "rand"), (9, "box"), (9, "gauss"), (9, "rand"),
                                    (4, "rand"), (11, "box"), (2, "rand"), (1, "rand")])
def test_convolution_vs_oracle(gpu, O, K, kind):
    rows, cols = 131, 517
    rng = np.random.default_rng(K * 31 + len(kind))
    padded = O.synth((rows + K - 1) * (cols + K - 1) * 3, 0xC0 + K).reshape(rows + K - 1, cols + K - 1, 3)
    if kind == "box":
        kern, norm = np.ones((K, K), np.int64), K * K
    elif kind == "gauss":
        b = np.array([1.0]);
        for _ in range(K - 1):
            b = np.convolve(b, [1, 1])
        kern = np.outer(b, b).astype(np.int64); norm = int(kern.sum())
    else:
        kern, norm = rng.integers(-9, 10, (K, K)).astype(np.int64), int(rng.integers(1, 50)) * (1 if K % 2 else -1)
    assert np.array_equal(gpu.parallel_convolution(padded, kern, norm), O.convolution_u8(padded, kern, norm))


def _sep_kernel(K, kind):
    b = np.array([1.0])
    for _ in range(K - 1):
        b = np.convolve(b, [1, 1])
    b = b.astype(np.int64)
    rng = np.random.default_rng(K * 7 + len(kind))
    if kind == "box":
        return np.ones((K, K), np.int64), K * K
    if kind == "gauss":                               # norm 256 (K=5) / 65536 (K=9): the byte-select division
        k = np.outer(b, b); return k, int(k.sum())
    if kind == "gauss_half":                          # power-of-two norm that is not 2^8 / 2^16: plain shift
        k = np.outer(b, b); return k, int(k.sum()) * 2
    if kind == "pos_oddnorm":                         # positive rank-1 taps, norm not a power of two: multiply-high
        u, v = rng.integers(0, 12, K), rng.integers(0, 9, K)
        k = np.outer(u, v); return k, int(k.sum()) + 3
    if kind == "signed":                              # derivative-like rank-1 taps: signed dp4a/dp2a and true division
        u, v = rng.integers(-9, 10, K), rng.integers(-7, 8, K)
        u[0], v[-1] = -3, 5
        return np.outer(u, v), -37
    if kind == "vbox_pow2":                           # constant vertical taps with a power-of-two norm: window sums + shift
        k = np.outer(np.full(K, 2, np.int64), b if K <= 5 else np.ones(K, np.int64)); return k, 64
    if kind == "wide_v":                              # horizontal sums overflow 16 bits: falls back to conv.cu's kernels
        u, v = np.ones(K, np.int64), np.full(K, 120, np.int64)
        return np.outer(u, v), 1000
    raise ValueError(kind)


@pytest.mark.parametrize("K", [3, 5, 7, 9])
@pytest.mark.parametrize("kind", ["box", "gauss", "gauss_half", "pos_oddnorm", "signed", "vbox_pow2", "wide_v"])
def test_separable_convolution_device_layouts(gpu, O, K, kind):
    """conv_sep.cu through the device entry points: dense and pitched rows at every 4-byte misalignment, widths
    that end inside a warp / a lane, odd and tiny row counts, uint8 and int32 outputs with aligned and unaligned
    output rows -- all against the oracle (old/parallel_convolution.c:8-24)."""
    L = gpu.lib()
    kern, norm = _sep_kernel(K, kind)
    taps = np.ascontiguousarray(kern.reshape(-1), np.int32)
    #        rows cols  in_pad in_off out_pad
    cases = [(131, 517, 0, 0, 0), (64, 128, 4, 0, 0), (37, 1100, 20, 4, 16), (2, 3, 0, 8, 4), (1, 1, 0, 0, 0),
             (200, 1024, 12, 12, 0), (75, 2100, 8, 0, 3), (33, 1028, 0, 0, 0)]
    for (rows, cols, in_pad, in_off, out_pad) in cases:
        pr, pc = rows + K - 1, cols + K - 1
        padded = O.synth(pr * pc * 3, 0xC5E0 + K + rows).reshape(pr, pc, 3)
        want = O.convolution_u8(padded, kern, norm)
        in_row = pc * 3
        in_pitch = (in_row + 3) // 4 * 4 + in_pad if in_pad else in_row
        src = gpu.DeviceBuffer(in_off + in_pitch * pr + 64)
        src.fill(0xEE)
        assert L.ompimg_copy2d_h2d(src.ptr + in_off, in_pitch, padded.ctypes.data, in_row, in_row, pr, 0) == 0
        for u8 in (True, False):
            ob = 1 if u8 else 4
            out_row = cols * 3 * ob
            out_pitch = out_row + out_pad * ob
            dst = gpu.DeviceBuffer(out_pitch * rows + 64)
            dst.fill(0x5A)
            fn = L.ompimg_conv_u8_dev if u8 else L.ompimg_conv_i32_dev
            ck(gpu, fn(src.ptr + in_off, in_pitch, rows, cols, K, taps.ctypes.data, norm, dst.ptr, out_pitch, 0))
            gpu.sync(0)
            raw = dst.download(out_pitch * rows).reshape(rows, out_pitch)
            got = raw[:, :out_row]
            exp = want.astype(np.uint8).reshape(rows, out_row) if u8 else want.reshape(rows, cols * 3).view(np.uint8).reshape(rows, out_row)
            assert np.array_equal(got, exp), (K, kind, rows, cols, in_pad, in_off, out_pad, u8)
            assert (raw[:, out_row:] == 0x5A).all(), "bytes between output rows were written"
            assert (dst.download(64, offset=out_pitch * rows) == 0x5A).all(), "bytes past the output were written"
            dst.free()
        src.free()



@pytest.mark.parametrize("K", [3, 5, 7, 9])
@pytest.mark.parametrize("kind", ["box", "gauss", "gauss_half", "pos_oddnorm", "vbox_pow2"])
def test_tma_convolution_device_layouts(gpu, O, K, kind):
    """conv_tma.cuh (16-byte aligned base and pitch: tensor-map copies, one warp per band x 128-pixel strip) through the
    device entry points: widths that end inside a lane / a warp / exactly on a strip, one to several strips, tiny, odd
    and band-crossing row counts, uint8 and int32 outputs with aligned and unaligned output rows, padding bytes that must
    survive -- all against the oracle (old/parallel_convolution.c:8-24)."""
    L = gpu.lib()
    kern, norm = _sep_kernel(K, kind)
    taps = np.ascontiguousarray(kern.reshape(-1), np.int32)
    #        rows cols  in_pad in_off out_pad
    cases = [(131, 517, 0, 0, 0), (64, 128, 16, 16, 0), (37, 1100, 128, 0, 16), (2, 3, 0, 32, 4), (1, 1, 0, 0, 0), (3, 129, 0, 0, 0),
             (1063, 384, 48, 0, 0), (75, 2100, 0, 0, 3), (33, 1028, 0, 0, 0), (19, 127, 0, 0, 1), (600, 256, 0, 0, 0)]
    for (rows, cols, in_pad, in_off, out_pad) in cases:
        pr, pc = rows + K - 1, cols + K - 1
        padded = O.synth(pr * pc * 3, 0x7A50 + K + rows).reshape(pr, pc, 3)
        want = O.convolution_u8(padded, kern, norm)
        in_row = pc * 3
        in_pitch = (in_row + 15) // 16 * 16 + in_pad
        src = gpu.DeviceBuffer(in_off + in_pitch * pr + 64)
        src.fill(0xEE)
        assert L.ompimg_copy2d_h2d(src.ptr + in_off, in_pitch, padded.ctypes.data, in_row, in_row, pr, 0) == 0
        for u8 in (True, False):
            ob = 1 if u8 else 4
            out_row = cols * 3 * ob
            out_pitch = out_row + out_pad * ob
            dst = gpu.DeviceBuffer(out_pitch * rows + 64)
            dst.fill(0x5A)
            fn = L.ompimg_conv_u8_dev if u8 else L.ompimg_conv_i32_dev
            ck(gpu, fn(src.ptr + in_off, in_pitch, rows, cols, K, taps.ctypes.data, norm, dst.ptr, out_pitch, 0))
            gpu.sync(0)
            raw = dst.download(out_pitch * rows).reshape(rows, out_pitch)
            got = raw[:, :out_row]
            exp = want.astype(np.uint8).reshape(rows, out_row) if u8 else want.reshape(rows, cols * 3).view(np.uint8).reshape(rows, out_row)
            assert np.array_equal(got, exp), (K, kind, rows, cols, in_pad, in_off, out_pad, u8)
            assert (raw[:, out_row:] == 0x5A).all(), "bytes between output rows were written"
            assert (dst.download(64, offset=out_pitch * rows) == 0x5A).all(), "bytes past the output were written"
            dst.free()
        src.free()


def test_tma_convolution_refill_race_stress(gpu, O):
    """Regression: the warp-autonomous kernels refill a ring stage (an async-proxy write) right after reading it.  On an
    L2-resident image with int32 per-lane output stores (unaligned output rows: the LSU queue fills with stores) the copy used
    to land before queued loads of the same stage had read it -- about two runs in three gave a few wrong values in a band's
    last rows.  Fixed by a proxy fence before the refill (conv_tma.cuh, stencil.cu); repeated here so that a regression shows."""
    L = gpu.lib()
    K, rows, cols = 9, 75, 2100
    kern, norm = _sep_kernel(K, "pos_oddnorm")
    taps = np.ascontiguousarray(kern.reshape(-1), np.int32)
    pr, pc = rows + K - 1, cols + K - 1
    padded = O.synth(pr * pc * 3, 0x7A50 + K + rows).reshape(pr, pc, 3)
    want = O.convolution_u8(padded, kern, norm).reshape(rows, cols * 3).astype(np.int32)
    in_pitch = (pc * 3 + 15) // 16 * 16
    src = to_dev_pitched(gpu, padded.reshape(pr, pc * 3), in_pitch)
    out_pitch = (cols * 3 + 3) * 4
    dst = gpu.DeviceBuffer(out_pitch * rows + 64)
    for rep in range(40):
        dst.fill(0x5A)
        ck(gpu, L.ompimg_conv_i32_dev(src.ptr, in_pitch, rows, cols, K, taps.ctypes.data, norm, dst.ptr, out_pitch, 0))
        gpu.sync(0)
        got = dst.download(out_pitch * rows).reshape(rows, out_pitch)[:, :cols * 12].copy().view(np.int32)
        assert np.array_equal(got, want), rep
    src.free(); dst.free()


def test_channel_sums_device_layouts(gpu, O):
    """a7 through the device entry point: dense and pitched rows, aligned and unaligned bases, rows whose byte
    count is not a multiple of 16 or of the 1536-byte tile (old/parallel_avg_pixel.c:17-30 sums)."""
    L = gpu.lib()
    #        rows cols  pitch_pad base_off
    cases = [(1, 1, 0, 0), (3, 5, 0, 0), (17, 512, 0, 0), (9, 515, 13, 0), (40, 1001, 0, 0), (25, 2048, 16, 0), (31, 700, 4, 0),
             (12, 333, 3, 16), (12, 333, 0, 5), (64, 4096, 128, 0), (2, 70000, 0, 0)]
    for (rows, cols, pad, off) in cases:
        img = O.synth(rows * cols * 3, 0xA7 + rows * cols).reshape(rows, cols * 3)
        pitch = cols * 3 + pad
        if pad and pad % 16 == 0:
            pitch = (cols * 3 + 15) // 16 * 16 + pad
        buf = gpu.DeviceBuffer(off + pitch * rows + 64)
        buf.fill(0xFF)
        assert L.ompimg_copy2d_h2d(buf.ptr + off, pitch, img.ctypes.data, cols * 3, cols * 3, rows, 0) == 0
        sums = gpu.DeviceBuffer(64)
        sums.fill(0)
        ck(gpu, L.ompimg_channel_sums_dev(buf.ptr + off, pitch, rows, cols, sums.ptr, 0))
        gpu.sync(0)
        got = sums.download(24).view(np.uint64)
        want = img.reshape(rows, cols, 3).astype(np.uint64).sum(axis=(0, 1))
        assert got.tolist() == want.tolist(), (rows, cols, pad, off)
        buf.free(); sums.free()


def test_avg_and_gray3_accumulate_and_ragged(gpu, O):
    for (rows, cols) in [(1, 1), (7, 5), (33, 100), (240, 1001), (1080, 1920)]:
        img = O.synth(rows * cols * 3, rows + cols).reshape(rows, cols, 3)
        assert gpu.parallel_avg_pixel(img, (5, 900, 12345)).tolist() == O.avg_pixel_u8(img, (5, 900, 12345)).tolist()
        g, mm = gpu.parallel_to_grayscale(img)
        wg, wmm = O.gray_avg3_u8(img)
        assert np.array_equal(g, wg) and mm.tolist() == wmm.tolist()
    const = np.full((9, 9, 3), 200, np.uint8)
    assert gpu.parallel_to_grayscale(const)[1].tolist() == [200, 200]


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_batch_properties(gpu, O):
    """cfg3 at full frame size on the device-resident path: a batch made of copies of two distinct frames
    must give copies of the two oracle results (checksum of checksums), wherever the frame sits in the batch."""
    L = gpu.lib()
    w, h, n = 3840, 2160, 24
    fb, eb = w * h * 3, w * h
    a = O.synth(fb, 0xB2000003)
    b = O.synth(fb, 0xB2000013)
    src = gpu.DeviceBuffer(n * fb)
    for f in range(n):
        src.upload(a if f % 3 else b, offset=f * fb)
    dst = gpu.DeviceBuffer(n * eb)
    dst.fill(0)
    ck(gpu, L.ompimg_rgb_to_edge_dev(src.ptr, w * 3, 3, dst.ptr, w, w, h, n, fb, eb, 0))
    gpu.sync(0)
    ea, eb_ = O.rgb_to_edge(a.reshape(h, w, 3)), O.rgb_to_edge(b.reshape(h, w, 3))
    for f in range(n):
        got = dst.download(eb, offset=f * eb).reshape(h, w)
        assert np.array_equal(got, ea if f % 3 else eb_), f
    # in-place grayscale of the same batch, twice: pass 2 is the reference's fixed point behaviour
    ck(gpu, L.ompimg_gray_inplace_dev(src.ptr, w * 3, w, h, 3, n, fb, 0))
    gpu.sync(0)
    ga = O.gray_inplace(a.copy().reshape(h, w, 3), 3)
    assert np.array_equal(src.download(fb, offset=1 * fb).reshape(h, w, 3), ga)
    ck(gpu, L.ompimg_gray_inplace_dev(src.ptr, w * 3, w, h, 3, n, fb, 0))
    gpu.sync(0)
    assert np.array_equal(src.download(fb, offset=(n - 1) * fb).reshape(h, w, 3), O.gray_inplace(ga.copy(), 3))


def test_full_size_convolution_box_property(gpu, O):
    """cfg4 (8192x8192 RGB): box and binomial blurs on synthetic data match the oracle on sampled windows (the oracle on
    the full image takes too long for CI) -- for the dense (cols+K-1)*3 pitch (conv_sep.cuh) and for 128-byte aligned
    rows (conv_tma.cuh); the windows sit on band seams, strip seams and the image's last rows / columns."""
    L = gpu.lib()
    rows = cols = 8192
    for K in (5, 9):
        pr, pc = rows + K - 1, cols + K - 1
        b = np.array([1.0])
        for _ in range(K - 1):
            b = np.convolve(b, [1, 1])
        gauss = np.outer(b, b).astype(np.int64)
        for pitch in (pc * 3, (pc * 3 + 127) // 128 * 128):
            src = gpu.DeviceBuffer(pr * pitch)
            gpu.synth_dev(src, pr * pitch, 0xB2000004)
            out = gpu.DeviceBuffer(rows * cols * 3)
            for kern in (np.ones((K, K), np.int64), gauss):
                norm = int(kern.sum())
                taps = np.ascontiguousarray(kern.reshape(-1), np.int32)
                out.fill(0)
                ck(gpu, L.ompimg_conv_u8_dev(src.ptr, pitch, rows, cols, K, taps.ctypes.data, norm, out.ptr, cols * 3, 0))
                gpu.sync(0)
                for (y0, x0, hh, ww) in ((4001, 3001, 40, 64), (0, 0, 24, 300), (rows - 30, cols - 200, 30, 200), (150, 8192 - 131, 400, 131)):
                    win = np.stack([src.download((ww + K - 1) * 3, offset=(y0 + r) * pitch + x0 * 3) for r in range(hh + K - 1)]).reshape(hh + K - 1, ww + K - 1, 3)
                    want = O.convolution_u8(win, kern, norm).astype(np.uint8)
                    got = np.stack([out.download(ww * 3, offset=((y0 + r) * cols + x0) * 3) for r in range(hh)]).reshape(hh, ww, 3)
                    assert np.array_equal(got, want), (K, pitch, y0, x0)
            src.free(); out.free()


# ------------------------------------------------------------------------------------------------ drop-in CLI
def _run_cli(binary, image, passes, tmp_path, env_extra=None):
    import hashlib, os, subprocess
    from common import ROOT
    exe = ROOT / "oracle" / "_ref" / "bin" / binary
    img = ROOT / "oracle" / "_ref" / "fixtures" / image
    if not exe.exists() or not img.exists():
        pytest.skip("drop-in binaries not built (oracle/Makefile `dropin`, needs /root/reference at build time)")
    out = tmp_path / f"{binary}_{passes}.png"
    r = subprocess.run([str(exe), str(img), str(out), str(passes)], capture_output=True, text=True, env=dict(os.environ, **(env_extra or {})))
    assert r.returncode == 0, r.stderr
    return hashlib.md5(out.read_bytes()).hexdigest(), r.stdout


def test_reference_main_c_linked_against_the_library(gpu, golden, tmp_path):
    """BASELINE config 1: the reference's unmodified main.c, linked with libompimg_b200 instead of its own
    parallel_to_grayscale.c, writes a PNG byte-identical to the reference's golden images/test_gray.png."""
    md5, out = _run_cli("grayscale_b200", "test.jpg", 1, tmp_path)
    assert md5 == golden["test_jpg"]["golden_png_md5"] == "044e265e3fd33938ff9a655156297af6"
    assert "Compute kernel" in out                                  # main.c:39, the line the workers' scripts show
    assert _run_cli("grayscale_b200", "test.jpg", 3, tmp_path)[0] == golden["test_jpg"]["strict_p3_png_md5"]
    assert _run_cli("grayscale_b200", "test.jpg", 1, tmp_path, {"OMPIMG_LUMA": "fma"})[0] == golden["test_jpg"]["native_p1_png_md5"]
    # OMP_NUM_THREADS is what the reference's callers vary (1/2/4/6); the GPU library ignores it: same bytes
    for t in ("1", "6"):
        assert _run_cli("grayscale_b200", "test.jpg", 1, tmp_path, {"OMP_NUM_THREADS": t})[0] == golden["test_jpg"]["golden_png_md5"]


def test_reference_main_with_sobel_c_linked_against_the_library(gpu, golden, tmp_path):
    """BASELINE config 2: main_with_sobel.c (gray -> extract -> Sobel -> replicate) with both kernels on the GPU."""
    assert _run_cli("sobel_b200", "more_than_one_mega_photo.jpg", 1, tmp_path)[0] == golden["photo_jpg"]["strict_p1_png_md5"]
    assert _run_cli("sobel_b200", "more_than_one_mega_photo.jpg", 2, tmp_path)[0] == golden["photo_jpg"]["strict_p2_png_md5"]
    assert _run_cli("sobel_b200", "more_than_one_mega_photo.jpg", 1, tmp_path, {"OMPIMG_LUMA": "fma"})[0] == golden["photo_jpg"]["native_p1_png_md5"]
