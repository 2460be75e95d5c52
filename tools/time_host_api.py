import importlib, time, sys, numpy as np
sys.path.insert(0, ".")
pkg = importlib.import_module("openmp-parallel-computing_b200"); pkg.init(0)
img = np.random.default_rng(0).integers(0,256,(1080,1920,3),dtype=np.uint8)
for name, arr in (("pageable", img.copy()),):
    pkg.convert_to_grayscale(arr)
    for rep in range(3):
        t0=time.perf_counter()
        for _ in range(50): pkg.convert_to_grayscale(arr)
        print(name, "gray ms/call", (time.perf_counter()-t0)/50*1e3)
pb = pkg.PinnedBuffer(img.nbytes); parr = pb.array.reshape(img.shape); parr[:] = img
pkg.convert_to_grayscale(parr)
t0=time.perf_counter()
for _ in range(50): pkg.convert_to_grayscale(parr)
print("pinned gray ms/call", (time.perf_counter()-t0)/50*1e3)
plane = np.ascontiguousarray(img[:,:,0]); dst=np.zeros_like(plane)
pkg.sobel_edge(plane,dst)
t0=time.perf_counter()
for _ in range(50): pkg.sobel_edge(plane,dst)
print("pageable sobel ms/call", (time.perf_counter()-t0)/50*1e3)
t0=time.perf_counter()
for _ in range(20): pkg.gray_sobel_pass(img)
print("pageable pass ms/call", (time.perf_counter()-t0)/20*1e3)
L=pkg.lib(); d=pkg.DeviceBuffer(img.nbytes); 
t0=time.perf_counter()
for _ in range(50):
    L.ompimg_copy_h2d(d.ptr, parr.ctypes.data, img.nbytes, 0); L.ompimg_gray_inplace_dev(d.ptr, 1920*3, 1920,1080,3,1,0,0); L.ompimg_copy_d2h(parr.ctypes.data, d.ptr, img.nbytes, 0); pkg.sync(0)
print("manual pinned h2d+kernel+d2h ms", (time.perf_counter()-t0)/50*1e3)
t0=time.perf_counter()
for _ in range(50):
    L.ompimg_gray_inplace_dev(d.ptr, 1920*3, 1920,1080,3,1,0,0); pkg.sync(0)
print("kernel only + sync ms", (time.perf_counter()-t0)/50*1e3)
