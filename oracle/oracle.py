"""oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

ctypes bindings for
  * liboracle.so      our CPU restatement (oracle.c), and
  * _ref/libref_*.so  the UNMODIFIED reference sources compiled by oracle/Makefile
                      (present only where `make -C oracle ref` ran, i.e. where
                      /root/reference exists; the built files travel to the GPU box),
plus small numpy restatements used as a third, independent statement of the
arithmetic.  Reference citations are relative to the reference repository root.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
_u8p = C.POINTER(C.c_ubyte)


def _ptr(a: np.ndarray, typ=_u8p):
    return a.ctypes.data_as(typ)


def build(force: bool = False) -> Path:
    """Compile oracle.c (and, when the reference tree is present, oracle/_ref)."""
    so = HERE / "liboracle.so"
    if force or not so.exists() or so.stat().st_mtime < (HERE / "oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "liboracle.so"], check=True, capture_output=True)
    if Path("/root/reference/monolithic/src/sobel.c").exists():
        subprocess.run(["make", "-C", str(HERE), "-j4", "ref"], check=True, capture_output=True)
    return so


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        so = build()
        L = C.CDLL(str(so))
        L.oracle_gray_inplace.argtypes = [_u8p, C.c_longlong, C.c_int, C.c_int]
        L.oracle_gray_lut.argtypes = [_u8p, C.c_int]
        L.oracle_rgb_to_luma.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_sobel.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_int, C.c_int]
        L.oracle_sobel_isqrt.argtypes = L.oracle_sobel.argtypes
        L.oracle_gray_sobel_pass.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p, _u8p, C.c_int]
        L.oracle_rgb_to_edge.argtypes = [_u8p, C.c_size_t, C.c_int, _u8p, C.c_size_t, C.c_int, C.c_int, _u8p, C.c_int]
        L.oracle_convolution_long.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long, C.c_void_p]
        L.oracle_convolution_u8.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_avg_pixel_long.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.oracle_avg_pixel_u8.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.oracle_gray_avg3_long.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_gray_avg3_u8.argtypes = [_u8p, C.c_size_t, _u8p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.oracle_extract_plane.argtypes = [_u8p, C.c_longlong, C.c_int, _u8p]
        L.oracle_replicate_plane.argtypes = [_u8p, C.c_longlong, C.c_int, _u8p]
        L.oracle_synth_fill.argtypes = [_u8p, C.c_ulonglong, C.c_uint, C.c_ulonglong]
        L.oracle_omp_max_threads.restype = C.c_int
        L.oracle_table3_alloc.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.oracle_table3_alloc.restype = C.c_void_p
        L.oracle_table3_free.argtypes = [C.c_void_p]
        L.oracle_table3_free.restype = None
        for f in ("oracle_gray_inplace oracle_gray_lut oracle_rgb_to_luma oracle_sobel oracle_sobel_isqrt "
                  "oracle_gray_sobel_pass oracle_rgb_to_edge oracle_convolution_long oracle_convolution_u8 "
                  "oracle_avg_pixel_long oracle_avg_pixel_u8 oracle_gray_avg3_long oracle_gray_avg3_u8 "
                  "oracle_synth_fill oracle_extract_plane oracle_replicate_plane").split():
            getattr(L, f).restype = None
        _lib = L
    return _lib


# ----------------------------------------------------------------------------- restatement (C)
def gray_inplace(img: np.ndarray, channels: int, fma: bool = False) -> np.ndarray:
    """monolithic/src/parallel_to_grayscale.c:5-17 on a flat/HxWxC uint8 array, in place."""
    assert img.dtype == np.uint8 and img.flags.c_contiguous
    lib().oracle_gray_inplace(_ptr(img), img.size // channels, channels, int(fma))
    return img


def gray_lut(fma: bool = False) -> np.ndarray:
    lut = np.empty(1 << 24, np.uint8)
    lib().oracle_gray_lut(_ptr(lut), int(fma))
    return lut


def rgb_to_luma(rgb: np.ndarray, fma: bool = False) -> np.ndarray:
    h, w, ch = rgb.shape
    out = np.empty((h, w), np.uint8)
    lib().oracle_rgb_to_luma(_ptr(rgb), w * ch, _ptr(out), w, w, h, ch, int(fma))
    return out


def sobel(src: np.ndarray, dst: np.ndarray | None = None, isqrt: bool = False) -> np.ndarray:
    """monolithic/src/sobel.c:6-25; dst border bytes are left as they were."""
    h, w = src.shape
    if dst is None:
        dst = np.zeros((h, w), np.uint8)
    fn = lib().oracle_sobel_isqrt if isqrt else lib().oracle_sobel
    fn(_ptr(src), src.strides[0], _ptr(dst), dst.strides[0], w, h)
    return dst


def gray_sobel_pass(img: np.ndarray, edge: np.ndarray | None = None, fma: bool = False):
    """One iteration of monolithic/src/main_with_sobel.c:51-74, in place on img (HxWxC)."""
    h, w, ch = img.shape
    gray = np.empty((h, w), np.uint8)
    if edge is None:
        edge = np.zeros((h, w), np.uint8)   # malloc'ed, zero pages in practice
    lib().oracle_gray_sobel_pass(_ptr(img), w, h, ch, _ptr(gray), _ptr(edge), int(fma))
    return img, edge


def rgb_to_edge(rgb: np.ndarray, edge: np.ndarray | None = None, fma: bool = False) -> np.ndarray:
    h, w, ch = rgb.shape
    if edge is None:
        edge = np.zeros((h, w), np.uint8)
    scratch = np.empty((h, w), np.uint8)
    lib().oracle_rgb_to_edge(_ptr(rgb), rgb.strides[0], ch, _ptr(edge), edge.strides[0], w, h, _ptr(scratch), int(fma))
    return edge


def convolution_u8(padded: np.ndarray, kernel: np.ndarray, norm: int) -> np.ndarray:
    """old/parallel_convolution.c:8-24 on uint8 storage, zeroed output; returns int32 HxWx3."""
    K = kernel.shape[0]
    rows, cols = padded.shape[0] - K + 1, padded.shape[1] - K + 1
    k = np.ascontiguousarray(kernel, np.int32)
    out = np.empty((rows, cols, 3), np.int32)
    lib().oracle_convolution_u8(_ptr(padded), padded.strides[0], rows, cols, K, k.ctypes.data, int(norm), out.ctypes.data)
    return out


def convolution_long(padded: np.ndarray, kernel: np.ndarray, norm: int, out: np.ndarray) -> np.ndarray:
    K = kernel.shape[0]
    rows, cols = padded.shape[0] - K + 1, padded.shape[1] - K + 1
    assert padded.dtype == np.int64 and out.dtype == np.int64 and kernel.dtype == np.int64
    lib().oracle_convolution_long(padded.ctypes.data, rows, cols, K, kernel.ctypes.data, int(norm), out.ctypes.data)
    return out


def avg_pixel_u8(img: np.ndarray, avgs_in=(0, 0, 0)) -> np.ndarray:
    avgs = np.array(avgs_in, np.int64)
    lib().oracle_avg_pixel_u8(_ptr(img), img.strides[0], img.shape[0], img.shape[1], avgs.ctypes.data)
    return avgs


def avg_pixel_long(img: np.ndarray, avgs_in=(0, 0, 0)) -> np.ndarray:
    avgs = np.array(avgs_in, np.int64)
    lib().oracle_avg_pixel_long(img.ctypes.data, img.shape[0], img.shape[1], avgs.ctypes.data)
    return avgs


def gray_avg3_u8(img: np.ndarray):
    out = np.empty_like(img)
    mm = np.zeros(2, np.int32)
    lib().oracle_gray_avg3_u8(_ptr(img), img.strides[0], _ptr(out), out.strides[0], img.shape[0], img.shape[1], mm.ctypes.data)
    return out, mm


def gray_avg3_long(img: np.ndarray):
    out = np.empty_like(img)
    mm = np.zeros(2, np.int64)
    lib().oracle_gray_avg3_long(img.ctypes.data, img.shape[0], img.shape[1], out.ctypes.data, mm.ctypes.data)
    return out, mm


def synth(nbytes: int, seed: int, offset: int = 0) -> np.ndarray:
    out = np.empty(nbytes, np.uint8)
    lib().oracle_synth_fill(_ptr(out), nbytes, seed & 0xFFFFFFFF, offset)
    return out


# ----------------------------------------------------------------------------- numpy restatements
CR, CG, CB = np.float32(0.299), np.float32(0.587), np.float32(0.114)


def np_luma_strict(r, g, b):
    """(c_r*r + c_g*g) + c_b*b with every product and sum rounded to fp32, truncated."""
    r, g, b = (np.asarray(v, np.uint8).astype(np.float32) for v in (r, g, b))
    return ((CR * r + CG * g) + CB * b).astype(np.uint8)


def np_luma_fma(r, g, b):
    """fma(c_b,b, fma(c_r,r, RN(c_g*g))) emulated exactly in float64 + one rounding per fma
    (products of fp32 by 8-bit ints and their sums with an fp32 fit in float64 exactly)."""
    r, g, b = (np.asarray(v, np.uint8).astype(np.float64) for v in (r, g, b))
    pg = (CG * np.asarray(g, np.float32)).astype(np.float32).astype(np.float64)
    t = (np.float64(CR) * r + pg).astype(np.float32).astype(np.float64)
    v = (np.float64(CB) * b + t).astype(np.float32)
    return v.astype(np.uint8)


def np_synth(nbytes: int, seed: int, offset: int = 0) -> np.ndarray:
    gi = np.arange(offset, offset + nbytes, dtype=np.uint64)
    x = (np.uint32(seed & 0xFFFFFFFF) + (gi >> np.uint64(2)).astype(np.uint32) * np.uint32(0x9E3779B9)).astype(np.uint32)
    x ^= x >> np.uint32(16); x *= np.uint32(0x7FEB352D); x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B); x ^= x >> np.uint32(16)
    return ((x >> ((gi & np.uint64(3)) * np.uint64(8)).astype(np.uint32)) & np.uint32(0xFF)).astype(np.uint8)


def np_sobel(src: np.ndarray, dst: np.ndarray | None = None) -> np.ndarray:
    """min(255, isqrt(gx^2+gy^2)) on the interior, integer arithmetic only."""
    h, w = src.shape
    if dst is None:
        dst = np.zeros((h, w), np.uint8)
    if h < 3 or w < 3:
        return dst
    s = src.astype(np.int64)
    a, b, c = s[:-2], s[1:-1], s[2:]
    gx = (a[:, 2:] + 2 * b[:, 2:] + c[:, 2:]) - (a[:, :-2] + 2 * b[:, :-2] + c[:, :-2])
    gy = (a[:, :-2] + 2 * a[:, 1:-1] + a[:, 2:]) - (c[:, :-2] + 2 * c[:, 1:-1] + c[:, 2:])
    m = gx * gx + gy * gy
    r = np.floor(np.sqrt(m.astype(np.float64))).astype(np.int64)
    r = np.where(r * r > m, r - 1, r)
    r = np.where((r + 1) * (r + 1) <= m, r + 1, r)
    dst[1:-1, 1:-1] = np.minimum(r, 255).astype(np.uint8)
    return dst


# ----------------------------------------------------------------------------- compiled reference
class Ref:
    """The unmodified reference objects (oracle/_ref/libref_<variant>.so)."""

    def __init__(self, variant: str = "strict"):
        p = REF_DIR / f"libref_{variant}.so"
        if not p.exists():
            raise FileNotFoundError(p)
        self.variant = variant
        self.L = C.CDLL(str(p))
        # monolithic/include/parallel_to_grayscale.h:4, monolithic/include/sobel.h:3-5
        self.L.convert_to_grayscale.argtypes = [_u8p, C.c_int, C.c_int, C.c_int]
        self.L.convert_to_grayscale.restype = None
        self.L.sobel_edge.argtypes = [_u8p, _u8p, C.c_int, C.c_int]
        self.L.sobel_edge.restype = None

    def gray(self, img: np.ndarray) -> np.ndarray:
        h, w, ch = img.shape
        self.L.convert_to_grayscale(_ptr(img), w, h, ch)
        return img

    def sobel(self, src: np.ndarray, dst: np.ndarray | None = None) -> np.ndarray:
        h, w = src.shape
        assert src.flags.c_contiguous
        if dst is None:
            dst = np.zeros((h, w), np.uint8)
        self.L.sobel_edge(_ptr(src), _ptr(dst), w, h)
        return dst

    def gray_sobel_pass(self, img: np.ndarray, edge: np.ndarray | None = None):
        """main_with_sobel.c:51-74 composed from the reference's two functions; the two
        glue loops (:57-59, :65-73) are plain strided copies done here in numpy."""
        h, w, ch = img.shape
        self.gray(img)
        gray = np.ascontiguousarray(img[:, :, 0])
        if edge is None:
            edge = np.zeros((h, w), np.uint8)
        self.sobel(gray, edge)
        img[:, :, 0] = edge; img[:, :, 1] = edge; img[:, :, 2] = edge
        return img, edge

    def lut(self) -> np.ndarray:
        idx = np.arange(1 << 24, dtype=np.uint32)
        rgb = np.empty((1 << 12, 1 << 12, 3), np.uint8)
        flat = rgb.reshape(-1, 3)
        flat[:, 0] = idx >> 16; flat[:, 1] = (idx >> 8) & 255; flat[:, 2] = idx & 255
        self.gray(rgb)
        return flat[:, 0].copy()


class RefOld:
    """old/parallel_*.c compiled with oracle/stub/utils.h at fixed (rows, cols, K, norm)."""

    def __init__(self, rows: int, cols: int, K: int, norm: int):
        p = REF_DIR / f"libref_old_{rows}_{cols}_{K}_{norm}.so"
        if not p.exists():
            raise FileNotFoundError(p)
        self.rows, self.cols, self.K, self.norm = rows, cols, K, norm
        self.L = C.CDLL(str(p))

    @staticmethod
    def _table3(arr: np.ndarray):
        """The long*** pointer table the reference's output parameters expect (built in C: oracle_table3_alloc)."""
        rows, cols, _ = arr.shape
        assert arr.dtype == np.int64 and arr.flags.c_contiguous
        t = lib().oracle_table3_alloc(arr.ctypes.data, rows, cols)
        if not t:
            raise MemoryError("oracle_table3_alloc")

        class _Keep:
            def __del__(self_inner, t=t):
                lib().oracle_table3_free(t)
        return C.c_void_p(t), _Keep()

    def convolution(self, padded: np.ndarray, kernel: np.ndarray, out: np.ndarray) -> np.ndarray:
        assert padded.shape == (self.rows + self.K - 1, self.cols + self.K - 1, 3) and padded.dtype == np.int64
        top, keep = self._table3(out)
        self.L.parallel_convolution(C.c_void_p(padded.ctypes.data), C.c_void_p(kernel.ctypes.data), top)
        return out

    def avg_pixel(self, img: np.ndarray, avgs_in=(0, 0, 0)) -> np.ndarray:
        avgs = np.array(avgs_in, np.int64)
        self.L.parallel_avg_pixel(C.c_void_p(img.ctypes.data), C.c_void_p(avgs.ctypes.data))
        return avgs

    def gray_avg3(self, img: np.ndarray):
        out = np.zeros_like(img)
        mm = np.zeros(2, np.int64)
        top, keep = self._table3(out)
        self.L.parallel_to_grayscale(C.c_void_p(img.ctypes.data), top, C.c_void_p(mm.ctypes.data))
        return out, mm


def jpeg_backend(width: int, height: int, comps: list, ycbcr: bool) -> np.ndarray:
    """stb_image's JPEG back end (dequantise, IDCT, resample, colour) restated, from per-component coefficient blocks:
    comps[k] = {"h", "v", "bw", "bh", "y", "coeff" int16 [bh][bw][64], "dequant" uint16 [64]} (monolithic/include/stb_image.h)."""
    n = len(comps)
    L = lib()
    L.oracle_jpeg_backend.restype = C.c_int
    hmax, vmax = max(c["h"] for c in comps), max(c["v"] for c in comps)
    coeff = [np.ascontiguousarray(c["coeff"], np.int16) for c in comps]
    dq = [np.ascontiguousarray(c["dequant"], np.uint16) for c in comps]
    PP = C.c_void_p * n
    IA = C.c_int * n
    out = np.empty((height, width, 3) if n == 3 else (height, width), np.uint8)
    rc = L.oracle_jpeg_backend(C.c_int(width), C.c_int(height), C.c_int(n), C.c_int(hmax), C.c_int(vmax), C.c_int(int(ycbcr)),
                               PP(*[a.ctypes.data for a in coeff]), PP(*[a.ctypes.data for a in dq]),
                               IA(*[c["h"] for c in comps]), IA(*[c["v"] for c in comps]), IA(*[c["bw"] for c in comps]), IA(*[c["bh"] for c in comps]),
                               IA(*[c["y"] for c in comps]), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def stb_decode(path) -> np.ndarray | None:
    """the reference's own stb_image.h (oracle/_ref/bin/stb_decode, built by `make -C oracle ref`): stbi_load(path, ..., 0)"""
    exe = REF_DIR / "bin" / "stb_decode"
    if not exe.exists():
        return None
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".raw") as t:
        r = subprocess.run([str(exe), str(path), t.name], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("stb_decode failed: " + r.stderr)
        w, h, ch = map(int, r.stdout.split())
        a = np.fromfile(t.name, np.uint8)
    return a.reshape(h, w, ch) if ch > 1 else a.reshape(h, w)


def omp_max_threads() -> int:
    return int(lib().oracle_omp_max_threads())


def have_ref_old(rows: int, cols: int, K: int, norm: int) -> bool:
    return (REF_DIR / f"libref_old_{rows}_{cols}_{K}_{norm}.so").exists()


def have_ref(variant: str = "strict") -> bool:
    return (REF_DIR / f"libref_{variant}.so").exists()
