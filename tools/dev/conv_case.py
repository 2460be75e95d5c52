"""development probe: one convolution case through the device entry point, repeated, mismatches vs the oracle located"""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import oracle as O
O.build()
gpu = importlib.import_module("openmp-parallel-computing_b200")
gpu.init(0)
L = gpu.lib()
from test_parity_gpu import _sep_kernel
K, kind, rows, cols, in_pad, in_off, out_pad = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 20
kern, norm = _sep_kernel(K, kind)
taps = np.ascontiguousarray(kern.reshape(-1), np.int32)
pr, pc = rows + K - 1, cols + K - 1
padded = O.synth(pr * pc * 3, 0x7A50 + K + rows).reshape(pr, pc, 3)
want = O.convolution_u8(padded, kern, norm)
in_row = pc * 3
in_pitch = (in_row + 15) // 16 * 16 + in_pad
src = gpu.DeviceBuffer(in_off + in_pitch * pr + 64); src.fill(0xEE)
assert L.ompimg_copy2d_h2d(src.ptr + in_off, in_pitch, padded.ctypes.data, in_row, in_row, pr, 0) == 0
for u8 in (True, False):
    ob = 1 if u8 else 4
    out_row = cols * 3 * ob; out_pitch = out_row + out_pad * ob
    dst = gpu.DeviceBuffer(out_pitch * rows + 64)
    bad = 0
    for rep in range(reps):
        dst.fill(0x5A)
        fn = L.ompimg_conv_u8_dev if u8 else L.ompimg_conv_i32_dev
        assert fn(src.ptr + in_off, in_pitch, rows, cols, K, taps.ctypes.data, norm, dst.ptr, out_pitch, 0) == 0
        gpu.sync(0)
        raw = dst.download(out_pitch * rows).reshape(rows, out_pitch)
        got = raw[:, :out_row]
        exp = want.astype(np.uint8).reshape(rows, out_row) if u8 else want.reshape(rows, cols * 3).view(np.uint8).reshape(rows, out_row)
        if not np.array_equal(got, exp):
            bad += 1
            ys, xs = np.nonzero(got != exp)
            print(f"u8={u8} rep={rep}: {len(ys)} bytes differ; rows {sorted(set(ys.tolist()))[:12]} cols(px) {sorted(set((xs // (3 * ob)).tolist()))[:16]} got {got[ys[0], xs[0]]} want {exp[ys[0], xs[0]]}")
    print(f"u8={u8}: {bad}/{reps} bad")
