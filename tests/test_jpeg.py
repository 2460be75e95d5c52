"""The JPEG decode step in front of the kernels (monolithic/src/main.c:21 = stbi_load with the
reference's vendored stb_image v2.30).

Known answers: tests/golden/jpeg/expected.json holds, for 59 small JPEGs (tests/golden/make_jpeg_golden.py) and for the
reference's three images, the sha256 of what the REFERENCE's own stb_image.h decodes (compiled as it lies).

CPU tests: the product's host-side entropy decoder (ompimg_jpeg_parse needs no GPU) followed by the oracle's restatement
of stb's back end must reproduce those hashes -- which pins the oracle restatement and checks the product's serial half.
GPU tests: the whole product path (host entropy decoding + device dequantise/IDCT/upsample/colour) must reproduce them,
and decode + grayscale in one call must equal the reference's golden PNG pixels."""
import hashlib
import json

import numpy as np
import pytest

from common import GOLD, ROOT, fixture_raw, sha

EXPECTED = json.loads((GOLD / "jpeg" / "expected.json").read_text())
SMALL = sorted(k for k in EXPECTED if not k.startswith("ref:"))
REF_IMAGES = ["test.jpg", "more_than_one_mega_photo.jpg", "half_of_a_mega_photo.jpg"]


def _bytes(name):
    if name.startswith("ref:"):
        p = ROOT / "oracle" / "_ref" / "fixtures" / name[4:]
        if not p.exists():
            pytest.skip(f"{p.name} did not travel (oracle/_ref/fixtures)")
        return p.read_bytes()
    return (GOLD / "jpeg" / f"{name}.jpg").read_bytes()


def _check(name, img):
    e = EXPECTED[name]
    assert list(img.shape) == e["shape"], name
    assert hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest() == e["sha256"], name


@pytest.mark.parametrize("name", SMALL + ["ref:" + f for f in REF_IMAGES])
def test_host_entropy_decoder_plus_oracle_backend_equals_stb(pkg, O, name):
    j = pkg.Jpeg(_bytes(name))
    comps = [j.component(c) for c in range(3 if j.channels == 3 else 1)]
    _check(name, O.jpeg_backend(j.width, j.height, comps, j.ycbcr))
    j.close()


def test_parser_rejects_what_it_does_not_handle(pkg):
    good = _bytes("s37x29_420_prog")
    with pytest.raises(pkg.OmpimgError):
        pkg.Jpeg(b"not a jpeg at all")
    with pytest.raises(pkg.OmpimgError):
        pkg.Jpeg(good[:2] + bytes(200))              # SOI, then nothing a decoder could use: no frame header
    bad = bytearray(good)
    i = bad.index(b"\xff\xc2")
    bad[i + 4] = 12                                   # 12-bit precision
    with pytest.raises(pkg.OmpimgError):
        pkg.Jpeg(bytes(bad))
    bad = bytearray(good)
    bad[i + 1] = 0xC9                                 # arithmetic coding
    with pytest.raises(pkg.OmpimgError):
        pkg.Jpeg(bytes(bad))
    # a truncated scan still parses (like stb: what was decoded is kept), and never reads out of bounds
    j = pkg.Jpeg(good[:len(good) // 2])
    assert (j.width, j.height) == (37, 29)
    j.close()


def test_oracle_backend_agrees_with_compiled_stb_when_present(pkg, O):
    """the restated back end against the compiled reference directly (build container / GPU box with oracle/_ref)"""
    p = GOLD / "jpeg" / "s130x70_420_prog.jpg"
    want = O.stb_decode(p)
    if want is None:
        pytest.skip("oracle/_ref/bin/stb_decode not built")
    j = pkg.Jpeg(p.read_bytes())
    got = O.jpeg_backend(j.width, j.height, [j.component(c) for c in range(3)], j.ycbcr)
    assert np.array_equal(got, want)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL + ["ref:" + f for f in REF_IMAGES])
def test_gpu_decode_equals_stb(gpu, name):
    j = gpu.Jpeg(_bytes(name))
    n0 = gpu.launch_count()
    img = j.decode()
    assert gpu.launch_count() - n0 == (4 if j.channels == 3 else 2)      # one IDCT launch per component + upsample/colour
    _check(name, img)
    j.close()


@pytest.mark.gpu
def test_gpu_decode_to_device_feeds_the_kernels(gpu, O, golden):
    """decode into device memory, grayscale in place there, download: the reference's CLI (main.c:21-41) without the RGB upload"""
    data = _bytes("ref:test.jpg")
    j = gpu.Jpeg(data)
    L = gpu.lib()
    pitch = (j.width * 3 + 127) // 128 * 128
    buf = gpu.DeviceBuffer(pitch * j.height)
    assert L.ompimg_jpeg_to_rgb_dev(j.h_, buf.ptr, pitch, 0) == 0
    assert L.ompimg_gray_inplace_dev(buf.ptr, pitch, j.width, j.height, 3, 1, 0, 0) == 0
    gpu.sync(0)
    out = np.empty((j.height, j.width * 3), np.uint8)
    assert L.ompimg_copy2d_d2h(out.ctypes.data, j.width * 3, buf.ptr, pitch, j.width * 3, j.height, 0) == 0
    gpu.sync(0)
    assert sha(out) == golden["test_jpg"]["golden_gray_sha256"]         # = the decoded pixels of images/test_gray.png
    one = gpu.jpeg_gray(data, 1)
    assert sha(one) == golden["test_jpg"]["golden_gray_sha256"]
    assert sha(gpu.jpeg_gray(data, 2)) == golden["test_jpg"]["strict_p2_sha256"]
    gray = gpu.jpeg_gray(_bytes("gray_40x30_base"), 3)                   # 1 component: returned as decoded
    _check("gray_40x30_base", gray)
    buf.free(); j.close()
