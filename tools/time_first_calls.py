"""time the first calls of a fresh process (context creation, allocations, pinned staging)"""
import importlib, time, sys, numpy as np
sys.path.insert(0, ".")
t0 = time.perf_counter()
pkg = importlib.import_module("openmp-parallel-computing_b200"); L = pkg.lib()
t1 = time.perf_counter(); print("dlopen+bind", round((t1 - t0) * 1e3, 1), "ms")
pkg.init(0); t2 = time.perf_counter(); print("ompimg_init (context + sqrt check)", round((t2 - t1) * 1e3, 1), "ms")
img = np.random.default_rng(0).integers(0, 256, (3000, 2000, 3), dtype=np.uint8)
for i in range(3):
    t = time.perf_counter(); pkg.convert_to_grayscale(img); print("convert_to_grayscale call", i, round((time.perf_counter() - t) * 1e3, 2), "ms")
plane = np.ascontiguousarray(img[:, :, 0]); dst = np.zeros_like(plane)
for i in range(3):
    t = time.perf_counter(); pkg.sobel_edge(plane, dst); print("sobel_edge call", i, round((time.perf_counter() - t) * 1e3, 2), "ms")
