"""openmp-parallel-computing_b200 -- Python host side of libompimg_b200.so.

A thin ctypes mirror of the C ABI in include/ompimg_b200.h, plus numpy-facing functions that carry
the reference's own names and argument meaning so that tests read like calls into the reference:

    convert_to_grayscale(img)            monolithic/src/parallel_to_grayscale.c:5-17  (in place)
    sobel_edge(src, dst)                 monolithic/src/sobel.c:6-25                  (interior only)
    gray_sobel_pass(img, edge)           monolithic/src/main_with_sobel.c:51-74       (one pass)
    parallel_convolution(padded, k, n)   old/parallel_convolution.c:8-24
    parallel_avg_pixel(img, avgs)        old/parallel_avg_pixel.c:5-42
    parallel_to_grayscale(img)           old/parallel_to_grayscale.c:7-38

There is no CPU implementation in this package.  If the shared library has not been built, or no
B200 is visible, every call raises: nothing silently falls back.  (The name of this directory is
not a valid Python identifier; import it with importlib.import_module("openmp-parallel-computing_b200").)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
LIB_PATH = Path(os.environ.get("OMPIMG_LIB", HERE / "lib" / "libompimg_b200.so"))   # OMPIMG_LIB: A/B builds while tuning
NUM_STREAMS = 8
LUMA_STRICT, LUMA_FMA = 0, 1

_u8p = C.POINTER(C.c_ubyte)
_vp = C.c_void_p
_sz = C.c_size_t


class OmpimgError(RuntimeError):
    pass


def build(verbose: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into lib/libompimg_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(HERE / "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise OmpimgError("building libompimg_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout)
    return LIB_PATH


# name -> (restype, argtypes); the single source the symbol-export test checks against include/ompimg_b200.h
_SIGS = {
    "convert_to_grayscale": (None, [_vp, C.c_int, C.c_int, C.c_int]),
    "sobel_edge": (None, [_vp, _vp, C.c_int, C.c_int]),
    "ompimg_init": (C.c_int, [C.c_int]),
    "ompimg_device": (C.c_int, []),
    "ompimg_device_count": (C.c_int, []),
    "ompimg_sm_count": (C.c_int, []),
    "ompimg_last_error": (C.c_char_p, []),
    "ompimg_version": (C.c_char_p, []),
    "ompimg_set_luma_mode": (C.c_int, [C.c_int]),
    "ompimg_get_luma_mode": (C.c_int, []),
    "ompimg_sqrt_is_exact": (C.c_int, []),
    "ompimg_launch_count": (C.c_longlong, []),
    "ompimg_shutdown": (C.c_int, []),
    "ompimg_dev_alloc": (_vp, [_sz]),
    "ompimg_dev_free": (C.c_int, [_vp]),
    "ompimg_host_alloc": (_vp, [_sz]),
    "ompimg_host_free": (C.c_int, [_vp]),
    "ompimg_host_realloc_sized": (_vp, [_vp, _sz, _sz]),
    "ompimg_host_register": (C.c_int, [_vp, _sz]),
    "ompimg_host_unregister": (C.c_int, [_vp]),
    "ompimg_memset_dev": (C.c_int, [_vp, C.c_int, _sz, C.c_int]),
    "ompimg_copy_h2d": (C.c_int, [_vp, _vp, _sz, C.c_int]),
    "ompimg_copy_d2h": (C.c_int, [_vp, _vp, _sz, C.c_int]),
    "ompimg_copy_d2d": (C.c_int, [_vp, _vp, _sz, C.c_int]),
    "ompimg_copy2d_h2d": (C.c_int, [_vp, _sz, _vp, _sz, _sz, _sz, C.c_int]),
    "ompimg_copy2d_d2h": (C.c_int, [_vp, _sz, _vp, _sz, _sz, _sz, C.c_int]),
    "ompimg_stream_sync": (C.c_int, [C.c_int]),
    "ompimg_stream_wait_stream": (C.c_int, [C.c_int, C.c_int]),
    "ompimg_timer_create": (C.c_int, []),
    "ompimg_timer_start": (C.c_int, [C.c_int, C.c_int]),
    "ompimg_timer_stop": (C.c_int, [C.c_int, C.c_int]),
    "ompimg_timer_elapsed_ms": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    "ompimg_timer_destroy": (C.c_int, [C.c_int]),
    "ompimg_gray_inplace_dev": (C.c_int, [_vp, _sz, C.c_int, C.c_int, C.c_int, C.c_int, _sz, C.c_int]),
    "ompimg_rgb_to_luma_dev": (C.c_int, [_vp, _sz, C.c_int, _vp, _sz, C.c_int, C.c_int, C.c_int, _sz, _sz, C.c_int]),
    "ompimg_gray_inplace_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "ompimg_sobel_dev": (C.c_int, [_vp, _sz, _vp, _sz, C.c_int, C.c_int, C.c_int, _sz, _sz, C.c_int]),
    "ompimg_sobel_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "ompimg_rgb_to_edge_dev": (C.c_int, [_vp, _sz, C.c_int, _vp, _sz, C.c_int, C.c_int, C.c_int, _sz, _sz, C.c_int]),
    "ompimg_rgb_to_edge_band_dev": (C.c_int, [_vp, _sz, C.c_int, _vp, _sz, C.c_int, C.c_int, _vp, _vp, C.c_int]),
    "ompimg_replicate_dev": (C.c_int, [_vp, _sz, _vp, _sz, C.c_int, C.c_int, C.c_int, C.c_int, _sz, _sz, C.c_int]),
    "ompimg_gray_sobel_pass_dev": (C.c_int, [_vp, _sz, C.c_int, _vp, _sz, C.c_int, C.c_int, C.c_int, _sz, _sz, C.c_int]),
    "ompimg_gray_sobel_passes_dev": (C.c_int, [_vp, _sz, C.c_int, _vp, _sz, C.c_int, C.c_int, C.c_int, _sz, _sz, C.c_int, C.c_int]),
    "ompimg_gray_sobel_pass_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp]),
    "ompimg_rgb_to_edge_batch_host": (C.c_int, [_vp, _sz, C.c_int, _vp, _sz, C.c_int, C.c_int, C.c_int]),
    "ompimg_conv_i32_dev": (C.c_int, [_vp, _sz, C.c_int, C.c_int, C.c_int, _vp, C.c_int32, _vp, _sz, C.c_int]),
    "ompimg_conv_u8_dev": (C.c_int, [_vp, _sz, C.c_int, C.c_int, C.c_int, _vp, C.c_int32, _vp, _sz, C.c_int]),
    "ompimg_conv_i32_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int32, _vp]),
    "ompimg_conv_i32_accum_host": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int32, _vp]),
    "ompimg_channel_sums_dev": (C.c_int, [_vp, _sz, C.c_int, C.c_int, _vp, C.c_int]),
    "ompimg_avg_pixel_host": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "ompimg_gray_avg3_dev": (C.c_int, [_vp, _sz, _vp, _sz, C.c_int, C.c_int, _vp, C.c_int]),
    "ompimg_gray_avg3_host": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "ompimg_synth_fill_dev": (C.c_int, [_vp, C.c_ulonglong, C.c_uint, C.c_ulonglong, C.c_int]),
    "ompimg_byte_sum_dev": (C.c_int, [_vp, _sz, C.c_int, C.c_int, _vp, C.c_int]),
    "ompimg_selftest_luma_lut": (C.c_int, [_vp, C.c_int]),
    "ompimg_selftest_isqrt": (C.c_int, [_vp, C.c_uint, C.c_float]),
    "ompimg_bands_create": (_vp, [C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "ompimg_bands_destroy": (C.c_int, [_vp]),
    "ompimg_bands_rows": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ompimg_bands_synth": (C.c_int, [_vp, C.c_uint]),
    "ompimg_bands_upload": (C.c_int, [_vp, _vp]),
    "ompimg_bands_rgb_to_edge": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "ompimg_bands_download_edge": (C.c_int, [_vp, _vp]),
    "ompimg_bands_checksum": (C.c_int, [_vp, C.POINTER(C.c_ulonglong)]),
    "ompimg_bands_conv_u8": (C.c_int, [_vp, C.c_int, _vp, C.c_int32, C.POINTER(C.c_float)]),
    "ompimg_bands_download_conv": (C.c_int, [_vp, _vp]),
    "ompimg_bands_conv_checksum": (C.c_int, [_vp, C.POINTER(C.c_ulonglong)]),
    "ompimg_bands_channel_sums": (C.c_int, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_float)]),
    "ompimg_jpeg_probe": (C.c_int, [_vp, _sz, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ompimg_jpeg_parse": (_vp, [_vp, _sz]),
    "ompimg_jpeg_free": (None, [_vp]),
    "ompimg_jpeg_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ompimg_jpeg_to_rgb_dev": (C.c_int, [_vp, _vp, _sz, C.c_int]),
    "ompimg_jpeg_to_rgb_host": (C.c_int, [_vp, _vp]),
    "ompimg_jpeg_gray_host": (C.c_int, [_vp, _sz, C.c_int, _vp, _sz, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ompimg_jpeg_component": (C.c_int, [_vp, C.c_int] + [C.POINTER(C.c_int)] * 6 + [C.POINTER(_vp), C.POINTER(_vp)]),
    "ompimg_jpeg_color_transform": (C.c_int, [_vp]),
    "ompimg_ipc_export": (C.c_int, [_vp, _vp]),
    "ompimg_ipc_open": (_vp, [_vp]),
    "ompimg_ipc_close": (C.c_int, [_vp]),
    "ompimg_worker_serve": (C.c_int, [C.c_char_p, C.c_int]),
}

_lib = None


def lib() -> C.CDLL:
    """dlopen libompimg_b200.so and bind every entry point.  Raises if the library is not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise OmpimgError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            if "OMPIMG_LIB" in os.environ and not hasattr(L, name):
                continue                   # an older A/B build of the library (tools/ only): newer entry points are simply absent
            fn = getattr(L, name)          # AttributeError = a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _ck(rc: int, what: str):
    if rc != 0:
        raise OmpimgError(f"{what} failed ({rc}): {lib().ompimg_last_error().decode()}")


def init(device: int | None = None) -> int:
    """Bind the process to a GPU (default: LOCAL_RANK, else OMPIMG_DEVICE, else 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("OMPIMG_DEVICE", "0")))
    _ck(lib().ompimg_init(device), "ompimg_init")
    return device


def set_luma_mode(mode: int | str):
    m = {"strict": LUMA_STRICT, "fma": LUMA_FMA}.get(mode, mode)
    _ck(lib().ompimg_set_luma_mode(int(m)), "ompimg_set_luma_mode")


def launch_count() -> int:
    return int(lib().ompimg_launch_count())


def sync(stream: int = -1):
    _ck(lib().ompimg_stream_sync(stream), "ompimg_stream_sync")


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _u8c(a: np.ndarray, name: str) -> np.ndarray:
    if not (isinstance(a, np.ndarray) and a.dtype == np.uint8 and a.flags.c_contiguous):
        raise TypeError(f"{name} must be a C-contiguous uint8 numpy array")
    return a


# ------------------------------------------------------------------------------------------------
# device / pinned memory helpers (bench.py and the device-resident tests)
# ------------------------------------------------------------------------------------------------
class DeviceBuffer:
    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = lib().ompimg_dev_alloc(self.nbytes)
        if not self.ptr:
            raise OmpimgError(f"ompimg_dev_alloc({nbytes}): {lib().ompimg_last_error().decode()}")

    def upload(self, a: np.ndarray, stream: int = 0, offset: int = 0):
        a = np.ascontiguousarray(a)
        _ck(lib().ompimg_copy_h2d(self.ptr + offset, _ptr(a), a.nbytes, stream), "copy_h2d")
        sync(stream)

    def download(self, nbytes: int | None = None, stream: int = 0, offset: int = 0) -> np.ndarray:
        n = self.nbytes - offset if nbytes is None else nbytes
        out = np.empty(n, np.uint8)
        _ck(lib().ompimg_copy_d2h(_ptr(out), self.ptr + offset, n, stream), "copy_d2h")
        sync(stream)
        return out

    def fill(self, value: int, stream: int = 0):
        _ck(lib().ompimg_memset_dev(self.ptr, value, self.nbytes, stream), "memset")

    def free(self):
        if self.ptr:
            lib().ompimg_dev_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory exposed as a numpy uint8 array (`.array`)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        self.ptr = lib().ompimg_host_alloc(self.nbytes)
        if not self.ptr:
            raise OmpimgError(f"ompimg_host_alloc({nbytes}): {lib().ompimg_last_error().decode()}")
        self.array = np.ctypeslib.as_array((C.c_ubyte * self.nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            lib().ompimg_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Timer:
    def __init__(self):
        self.id = lib().ompimg_timer_create()
        if self.id < 0:
            raise OmpimgError("ompimg_timer_create: " + lib().ompimg_last_error().decode())

    def start(self, stream: int = 0):
        _ck(lib().ompimg_timer_start(self.id, stream), "timer_start")

    def stop(self, stream: int = 0):
        _ck(lib().ompimg_timer_stop(self.id, stream), "timer_stop")

    def ms(self) -> float:
        v = C.c_float()
        _ck(lib().ompimg_timer_elapsed_ms(self.id, C.byref(v)), "timer_elapsed")
        return float(v.value)


# ------------------------------------------------------------------------------------------------
# the reference's functions, on host (numpy) buffers
# ------------------------------------------------------------------------------------------------
def convert_to_grayscale(data: np.ndarray, width: int | None = None, height: int | None = None, channels: int | None = None):
    """monolithic/include/parallel_to_grayscale.h:4 -- in place on an HxWxC (or flat) uint8 array."""
    _u8c(data, "data")
    if width is None:
        if data.ndim != 3:
            raise ValueError("data must be H x W x C when width/height/channels are not given")
        height, width, channels = data.shape
    elif data.size < width * height * channels:
        raise ValueError("data holds fewer than width*height*channels bytes")
    _ck(lib().ompimg_gray_inplace_host(_ptr(data), width, height, channels), "convert_to_grayscale")
    return data


def sobel_edge(src: np.ndarray, dst: np.ndarray | None = None) -> np.ndarray:
    """monolithic/include/sobel.h:3-5 -- dst border bytes are left exactly as they were."""
    _u8c(src, "src")
    if src.ndim != 2:
        raise ValueError("src must be an H x W plane")
    h, w = src.shape
    if dst is None:
        dst = np.zeros((h, w), np.uint8)
    _u8c(dst, "dst")
    if dst.shape != src.shape:
        raise ValueError("src and dst must have the same shape")
    _ck(lib().ompimg_sobel_host(_ptr(src), _ptr(dst), w, h), "sobel_edge")
    return dst


def gray_sobel_pass(img: np.ndarray, edge: np.ndarray | None = None):
    """One iteration of monolithic/src/main_with_sobel.c:51-74, in place on img (HxWxC)."""
    _u8c(img, "img")
    if img.ndim != 3 or img.shape[2] not in (3, 4):
        raise ValueError("img must be H x W x 3 or H x W x 4")
    h, w, ch = img.shape
    if edge is not None:
        _u8c(edge, "edge")
        if edge.shape != (h, w):
            raise ValueError("edge must be HxW")
    _ck(lib().ompimg_gray_sobel_pass_host(_ptr(img), w, h, ch, _ptr(edge) if edge is not None else None), "gray_sobel_pass")
    return img, edge


def rgb_to_edge_batch(frames: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """Fused gray->Sobel on a batch NxHxWxC; returns NxHxW edge planes with zero borders."""
    _u8c(frames, "frames")
    if frames.ndim != 4 or frames.shape[3] not in (3, 4):
        raise ValueError("frames must be N x H x W x 3 or N x H x W x 4")
    n, h, w, ch = frames.shape
    if out is None:
        out = np.empty((n, h, w), np.uint8)
    _ck(lib().ompimg_rgb_to_edge_batch_host(_ptr(frames), h * w * ch, ch, _ptr(out), h * w, w, h, n), "rgb_to_edge_batch")
    return out


def _rgb3(a: np.ndarray, name: str) -> np.ndarray:
    _u8c(a, name)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError(f"{name} must be rows x cols x 3 (got shape {a.shape})")
    return a


def parallel_convolution(padded: np.ndarray, kernel: np.ndarray, norm: int, out: np.ndarray | None = None) -> np.ndarray:
    """old/parallel_convolution.c:8-24 on uint8 storage; returns int32 rows x cols x 3.  With `out` (int32, rows x cols x 3)
    the reference's own accumulate-then-divide statement applies: out = (out + sum) / norm (:17,:20)."""
    _rgb3(padded, "padded")
    kernel = np.asarray(kernel)
    if kernel.ndim != 2 or kernel.shape[0] != kernel.shape[1]:
        raise ValueError("kernel must be K x K")
    K = int(kernel.shape[0])
    rows, cols = padded.shape[0] - K + 1, padded.shape[1] - K + 1
    if rows <= 0 or cols <= 0:
        raise ValueError("padded must be (rows+K-1) x (cols+K-1) x 3")
    taps = np.ascontiguousarray(kernel, np.int32)
    if out is None:
        out = np.empty((rows, cols, 3), np.int32)
        _ck(lib().ompimg_conv_i32_host(_ptr(padded), rows, cols, K, _ptr(taps), int(norm), _ptr(out)), "parallel_convolution")
        return out
    if not (isinstance(out, np.ndarray) and out.dtype == np.int32 and out.flags.c_contiguous and out.shape == (rows, cols, 3)):
        raise ValueError("out must be a C-contiguous int32 array of shape rows x cols x 3")
    _ck(lib().ompimg_conv_i32_accum_host(_ptr(padded), rows, cols, K, _ptr(taps), int(norm), _ptr(out)), "parallel_convolution (accumulate)")
    return out


def parallel_avg_pixel(img: np.ndarray, avgs=(0, 0, 0)) -> np.ndarray:
    """old/parallel_avg_pixel.c:5-42 (accumulates onto avgs, then divides by rows*cols*3)."""
    _rgb3(img, "img")
    a = np.array(avgs, np.int64)
    _ck(lib().ompimg_avg_pixel_host(_ptr(img), img.shape[0], img.shape[1], _ptr(a)), "parallel_avg_pixel")
    return a


def parallel_to_grayscale(img: np.ndarray):
    """old/parallel_to_grayscale.c:7-38: (r+g+b)/3 to all channels and {min, max}."""
    _rgb3(img, "img")
    out = np.empty_like(img)
    mm = np.zeros(2, np.int32)
    _ck(lib().ompimg_gray_avg3_host(_ptr(img), _ptr(out), img.shape[0], img.shape[1], _ptr(mm)), "parallel_to_grayscale")
    return out, mm


def luma_lut(mode: int | str = LUMA_STRICT) -> np.ndarray:
    m = {"strict": LUMA_STRICT, "fma": LUMA_FMA}.get(mode, mode)
    lut = np.empty(1 << 24, np.uint8)
    _ck(lib().ompimg_selftest_luma_lut(_ptr(lut), int(m)), "selftest_luma_lut")
    return lut


def isqrt_table(count: int, bias: float = 0.5) -> np.ndarray:
    out = np.empty(count, np.uint8)
    _ck(lib().ompimg_selftest_isqrt(_ptr(out), count, bias), "selftest_isqrt")
    return out


def synth_dev(buf: DeviceBuffer, nbytes: int, seed: int, offset: int = 0, dst_offset: int = 0, stream: int = 0):
    _ck(lib().ompimg_synth_fill_dev(buf.ptr + dst_offset, nbytes, seed & 0xFFFFFFFF, offset, stream), "synth_fill")


# ------------------------------------------------------------------------------------------------
# JPEG decode in front of the kernels (monolithic/src/main.c:21: stbi_load)
# ------------------------------------------------------------------------------------------------
class Jpeg:
    """A parsed JPEG: entropy decoding done on the host (works without a GPU), pixels produced on the device."""

    def __init__(self, data: bytes):
        self._buf = np.frombuffer(data, np.uint8)
        self.h_ = lib().ompimg_jpeg_parse(_ptr(self._buf), self._buf.size)
        if not self.h_:
            raise OmpimgError("ompimg_jpeg_parse: " + lib().ompimg_last_error().decode())
        w, h, ch = C.c_int(), C.c_int(), C.c_int()
        _ck(lib().ompimg_jpeg_info(self.h_, C.byref(w), C.byref(h), C.byref(ch)), "jpeg_info")
        self.width, self.height, self.channels = w.value, h.value, ch.value

    def component(self, c: int) -> dict:
        """sampling factors, block grid, coefficient blocks (copy) and quantisation table of component c (tests)"""
        v = [C.c_int() for _ in range(6)]
        pc, pq = _vp(), _vp()
        _ck(lib().ompimg_jpeg_component(self.h_, c, *[C.byref(x) for x in v], C.byref(pc), C.byref(pq)), "jpeg_component")
        hs, vs, bw, bh, sx, sy = [x.value for x in v]
        coeff = np.ctypeslib.as_array((C.c_int16 * (bw * bh * 64)).from_address(pc.value)).reshape(bh, bw, 64).copy()
        dq = np.ctypeslib.as_array((C.c_uint16 * 64).from_address(pq.value)).copy()
        return {"h": hs, "v": vs, "bw": bw, "bh": bh, "x": sx, "y": sy, "coeff": coeff, "dequant": dq}

    @property
    def ycbcr(self) -> bool:
        return lib().ompimg_jpeg_color_transform(self.h_) == 1

    def decode(self) -> np.ndarray:
        """H x W x 3 (or H x W) uint8, equal to stbi_load(path, ..., 0)"""
        shape = (self.height, self.width, 3) if self.channels == 3 else (self.height, self.width)
        out = np.empty(shape, np.uint8)
        _ck(lib().ompimg_jpeg_to_rgb_host(self.h_, _ptr(out)), "jpeg_to_rgb_host")
        return out

    def close(self):
        if self.h_:
            lib().ompimg_jpeg_free(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def jpeg_gray(data: bytes, passes: int = 1) -> np.ndarray:
    """main.c:21-35: decode + `passes` x convert_to_grayscale, one device round trip"""
    buf = np.frombuffer(data, np.uint8)
    w, h, ch = C.c_int(), C.c_int(), C.c_int()
    _ck(lib().ompimg_jpeg_probe(_ptr(buf), buf.size, C.byref(w), C.byref(h), C.byref(ch)), "jpeg_probe")
    out = np.empty((h.value, w.value, ch.value) if ch.value == 3 else (h.value, w.value), np.uint8)
    _ck(lib().ompimg_jpeg_gray_host(_ptr(buf), buf.size, passes, _ptr(out), out.nbytes, C.byref(w), C.byref(h), C.byref(ch)), "jpeg_gray_host")
    return out


# ------------------------------------------------------------------------------------------------
# row bands over several GPUs of this process (BASELINE config 5)
# ------------------------------------------------------------------------------------------------
class Bands:
    def __init__(self, width: int, height: int, channels: int, devices):
        devs = (C.c_int * len(devices))(*devices)
        self.w, self.h, self.ch, self.n = width, height, channels, len(devices)
        self.h_ = lib().ompimg_bands_create(width, height, channels, len(devices), devs)
        if not self.h_:
            raise OmpimgError("ompimg_bands_create failed")

    def rows(self, i: int):
        r0, n = C.c_int(), C.c_int()
        _ck(lib().ompimg_bands_rows(self.h_, i, C.byref(r0), C.byref(n)), "bands_rows")
        return r0.value, n.value

    def synth(self, seed: int):
        _ck(lib().ompimg_bands_synth(self.h_, seed & 0xFFFFFFFF), "bands_synth")

    def upload(self, rgb: np.ndarray):
        _ck(lib().ompimg_bands_upload(self.h_, _ptr(_u8c(rgb, "rgb"))), "bands_upload")

    def rgb_to_edge(self) -> float:
        ms = C.c_float()
        _ck(lib().ompimg_bands_rgb_to_edge(self.h_, C.byref(ms)), "bands_rgb_to_edge")
        return float(ms.value)

    def download_edge(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.zeros((self.h, self.w), np.uint8)
        _ck(lib().ompimg_bands_download_edge(self.h_, _ptr(_u8c(out, "out"))), "bands_download_edge")
        return out

    def checksum(self) -> int:
        s = C.c_ulonglong()
        _ck(lib().ompimg_bands_checksum(self.h_, C.byref(s)), "bands_checksum")
        return int(s.value)

    def convolution(self, kernel: np.ndarray, norm: int) -> float:
        """a6 over the bands (old/parallel_convolution.c:12-23): the banded image is the pre-padded input; returns the device ms"""
        k = np.ascontiguousarray(kernel, np.int32)
        if k.ndim != 2 or k.shape[0] != k.shape[1]:
            raise OmpimgError("kernel must be K x K")
        ms = C.c_float()
        _ck(lib().ompimg_bands_conv_u8(self.h_, k.shape[0], k.ctypes.data, int(norm), C.byref(ms)), "bands_conv_u8")
        self.k_ = k.shape[0]
        return float(ms.value)

    def download_convolution(self) -> np.ndarray:
        out = np.zeros((self.h - self.k_ + 1, self.w - self.k_ + 1, 3), np.uint8)
        _ck(lib().ompimg_bands_download_conv(self.h_, _ptr(out)), "bands_download_conv")
        return out

    def convolution_checksum(self) -> int:
        s = C.c_ulonglong()
        _ck(lib().ompimg_bands_conv_checksum(self.h_, C.byref(s)), "bands_conv_checksum")
        return int(s.value)

    def channel_sums(self):
        """a7 over the bands (old/parallel_avg_pixel.c:17-30): per-device partial sums added on the host; (sums[3], device ms)"""
        s3 = (C.c_ulonglong * 3)()
        ms = C.c_float()
        _ck(lib().ompimg_bands_channel_sums(self.h_, s3, C.byref(ms)), "bands_channel_sums")
        return [int(v) for v in s3], float(ms.value)

    def close(self):
        if self.h_:
            lib().ompimg_bands_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
