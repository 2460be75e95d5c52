#!/usr/bin/env python
"""bench.py -- the headline benchmark of the hot path (BASELINE.json config 3):

    fused grayscale->Sobel (RGB in, edge plane out; monolithic/src/main_with_sobel.c:51-74 without the
    HBM round trips) on a synthetic batch of 1024 RGB 3840x2160 frames, sharded over the N GPUs of one
    box with no collective (frames are independent).  metric: Gpixel/s, whole job.

    python bench.py [--gpus N] [--steps K] [--warmup W]                   (N=1 runs in-process)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU)
    python bench.py --impl reference ...    the reference's CPU implementation on the host cores

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over the whole shard, device
resident (`value`); `e2e` is the same metric through the public host API with pinned host buffers,
H2D and D2H inside the timed region.  torch is used only for the multi-rank barrier / max-over-ranks.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H, CH = 3840, 2160, 3
TOTAL_FRAMES = 1024
PX = W * H
BYTES_PER_PX = 4.0          # algorithmic: 3 B RGB read + 1 B edge written
SEED = 0xB2000003


_REAL_STDOUT = None


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def shard(total: int, rank: int, world: int):
    """contiguous block of frames for `rank`: no frame is split, none is shared"""
    lo = total * rank // world
    hi = total * (rank + 1) // world
    return lo, hi - lo


class Dist:
    """barrier + max-over-ranks; torch.distributed only when world > 1"""

    def __init__(self, backend: str):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.td = None
        if self.world > 1 and backend:
            import torch
            import torch.distributed as td
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29533")
            if backend == "nccl":
                torch.cuda.set_device(self.local)
            td.init_process_group(backend=backend, rank=self.rank, world_size=self.world)
            self.td, self.torch, self.backend = td, torch, backend

    def _dev(self):
        return f"cuda:{self.local}" if self.backend == "nccl" else "cpu"

    def barrier(self):
        if self.td:
            t = self.torch.zeros(1, device=self._dev())
            self.td.all_reduce(t)
            if self.backend == "nccl":
                self.torch.cuda.synchronize()

    def max(self, v: float) -> float:
        if not self.td:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self._dev())
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def sum(self, v: float) -> float:
        if not self.td:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self._dev())
        self.td.all_reduce(t, op=self.td.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.td:
            self.td.destroy_process_group()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self) -> dict:
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        out, _ = self.p.communicate(timeout=10)
        sm, mx, reasons = [], 0, set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # "under load": the upper half of the samples (the sampler also sees the idle edges of the region)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx or None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if there is one"""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("sobel_warp_kernel<3,0>")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_pass_factory():
    """One iteration of main_with_sobel.c:51-74 on the host cores.  Uses the UNMODIFIED reference objects
    (oracle/_ref/libref_strict.so: convert_to_grayscale, sobel_edge) when they were built, else the oracle port."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as O
    O.build()
    L = O.lib()
    u8 = O._u8p
    if O.have_ref("strict"):
        R = O.Ref("strict").L

        def run(img, gray, edge, w, h, ch):
            R.convert_to_grayscale(img.ctypes.data_as(u8), w, h, ch)                       # main_with_sobel.c:54
            L.oracle_extract_plane(img.ctypes.data_as(u8), w * h, ch, gray.ctypes.data_as(u8))   # :57-59 (loop lives in main())
            R.sobel_edge(gray.ctypes.data_as(u8), edge.ctypes.data_as(u8), w, h)           # :62
            L.oracle_replicate_plane(edge.ctypes.data_as(u8), w * h, ch, img.ctypes.data_as(u8))  # :65-73
        return run, "reference", O
    def run(img, gray, edge, w, h, ch):
        L.oracle_gray_sobel_pass(img.ctypes.data_as(u8), w, h, ch, gray.ctypes.data_as(u8), edge.ctypes.data_as(u8), 0)
    return run, "port", O


def time_cpu(frames: int, reps: int, threads: int | None = None):
    """Gpx/s of the CPU pass over `frames` 4K frames, best of `reps`"""
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    run, kind, O = cpu_pass_factory()
    src = O.synth(frames * PX * CH, SEED).reshape(frames, H, W, CH)
    gray = np.empty((H, W), np.uint8)
    edge = np.zeros((H, W), np.uint8)
    best = None
    for _ in range(reps):
        work = src.copy()
        t0 = time.perf_counter()
        for f in range(frames):
            run(work[f], gray, edge, W, H, CH)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return frames * PX / best / 1e9, kind


def cpu_baseline_subprocess(frames: int, reps: int, threads: int):
    """the OpenMP thread count is fixed at library load, so each thread count runs in its own process"""
    code = (f"import sys; sys.path.insert(0, {str(ROOT)!r}); import bench, json; "
            f"v, k = bench.time_cpu({frames}, {reps}, {threads}); print(json.dumps([v, k]))")
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    if r.returncode != 0:
        raise RuntimeError("cpu baseline failed: " + r.stderr[-2000:])
    v, k = json.loads(r.stdout.strip().splitlines()[-1])
    return float(v), k


def run_reference(args, dist: Dist):
    """--impl reference: the reference's CPU implementation of the same pass, all host threads"""
    if dist.rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = args.cpu_frames
    run, kind, O = cpu_pass_factory()
    src = O.synth(frames * PX * CH, SEED).reshape(frames, H, W, CH)
    gray, edge = np.empty((H, W), np.uint8), np.zeros((H, W), np.uint8)
    work = src.copy()
    def step():
        np.copyto(work, src)
        t0 = time.perf_counter()
        for f in range(frames):
            run(work[f], gray, edge, W, H, CH)
        return time.perf_counter() - t0
    for _ in range(args.warmup):
        step()
    total = sum(step() for _ in range(args.steps))
    value = frames * PX * args.steps / total / 1e9
    sample = f"{frames} synthetic 3840x2160 RGB frames per step through the main_with_sobel.c pass ({kind}), OMP threads = {cores}"
    line = {
        "impl": "reference", "metric": "Gpixel/s", "value": value, "unit": "Gpixel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8 (fp32 luma)", "data": "synthetic",
        "config": {"workload": "cfg3: fused grayscale->Sobel, 1024 x 3840x2160 RGB frames (bounded CPU sample per step)", "frames_per_step": frames},
        "cpu_baseline": {"value": value, "unit": "Gpixel/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(local_rank: int):
    """Run this rank (and first-touch its pinned host buffers) on the CPU socket its GPU hangs off: with several ranks per
    box the end-to-end leg is bound by host memory and the inter-socket link, not by PCIe alone.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        bus = bus[4:] if len(bus.split(":")[0]) == 8 else bus           # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        return None
    return None


def run_ours(args, dist: Dist):
    pkg = importlib.import_module("openmp-parallel-computing_b200")
    numa = bind_to_gpu_numa_node(dist.local) if dist.world > 1 and not os.environ.get("OMPIMG_NO_NUMA") else None
    pkg.init(dist.local)
    L = pkg.lib()
    f0, nf = shard(args.frames, dist.rank, dist.world)
    fb, eb = PX * CH, PX
    src = pkg.DeviceBuffer(nf * fb)
    dst = pkg.DeviceBuffer(nf * eb)
    pkg.synth_dev(src, nf * fb, SEED, offset=f0 * fb)        # frame f of the job is bytes [f*fb, (f+1)*fb) of one stream
    dst.fill(0)
    pkg.sync()

    def step():
        rc = L.ompimg_rgb_to_edge_dev(src.ptr, W * CH, CH, dst.ptr, W, W, H, nf, fb, eb, 0)
        if rc:
            raise pkg.OmpimgError(L.ompimg_last_error().decode())

    for _ in range(max(args.warmup, 3)):
        step()
    pkg.sync()
    dist.barrier()
    sampler = ClockSampler(dist.local) if dist.rank == 0 else None
    t = pkg.Timer()
    n0 = pkg.launch_count()
    dist.barrier()
    pkg.sync()
    t.start(0)
    for _ in range(args.steps):
        step()
    t.stop(0)
    ms = t.ms()
    pkg.sync()
    dist.barrier()
    launches = pkg.launch_count() - n0
    ms_max = dist.max(ms)
    clocks = sampler.stop() if sampler else None
    value = args.frames * PX * args.steps / (ms_max * 1e-3) / 1e9

    # checksum so that a run that skipped the work cannot pass for one that did it
    edge0 = dst.download(eb).reshape(H, W)
    assert edge0[1:-1, 1:-1].any() and not edge0[0].any()

    peak, peak_src = measured_peak()
    kernel_ms = ms / args.steps                               # this rank's launches: one per step
    achieved = BYTES_PER_PX * nf * PX / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "sobel_warp_kernel<3,0> (fused RGB->luma->Sobel, warp-autonomous bulk-copy rings)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(),
                "algorithmic_bytes_per_launch": BYTES_PER_PX * nf * PX, "frac_of_nominal_8TBs": achieved / 8000.0}

    # ---- end to end through the public host API: pinned host frames -> H2D -> kernel -> D2H edge planes
    ne = min(args.e2e_frames, nf)
    hin, hout = pkg.PinnedBuffer(ne * fb), pkg.PinnedBuffer(ne * eb)
    pkg.synth_dev(src, ne * fb, SEED, offset=f0 * fb)         # same bytes as the shard's first frames
    rc = L.ompimg_copy_d2h(hin.ptr, src.ptr, ne * fb, 0); assert rc == 0
    pkg.sync()
    frames_host = hin.array.reshape(ne, H, W, CH)
    edges_host = hout.array.reshape(ne, H, W)
    for _ in range(2):
        pkg.rgb_to_edge_batch(frames_host, edges_host)
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        pkg.rgb_to_edge_batch(frames_host, edges_host)
    dt = dist.max(time.perf_counter() - t0)
    assert np.array_equal(edges_host[0][1:-1, 1:-1], edge0[1:-1, 1:-1])       # same result as the resident path
    e2e = {"value": dist.world * ne * PX * args.e2e_steps / dt / 1e9, "unit": "Gpixel/s", "h2d_bytes_per_step": ne * fb, "d2h_bytes_per_step": ne * eb,
           "frames_per_step_per_gpu": ne, "steps": args.e2e_steps, "api": "ompimg_rgb_to_edge_batch_host (pinned host buffers, 8 streams)",
           "timing": "host wall clock around the synchronous call, max over ranks",
           "numa_node_rank0": numa}
    hin.free(); hout.free()

    cpu = None
    if dist.rank == 0 and dist.world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v_all, kind = cpu_baseline_subprocess(args.cpu_frames, 3, cores)
        v_one, _ = cpu_baseline_subprocess(max(1, args.cpu_frames // 4), 2, 1)
        cpu = {"value": v_all, "unit": "Gpixel/s", "cores": cores, "kind": kind, "value_1_thread": v_one,
               "sample": f"{args.cpu_frames} synthetic 3840x2160 RGB frames through the main_with_sobel.c pass, best of 3 (1 thread: {max(1, args.cpu_frames // 4)} frames, best of 2)"}

    extras = None
    if dist.rank == 0 and dist.world == 1 and not args.no_extras:
        extras = other_kernels(pkg, L, peak)

    if dist.rank == 0:
        line = {
            "metric": "Gpixel/s", "value": value, "unit": "Gpixel/s", "n_gpus": dist.world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 (fp32 luma, f16x2 Sobel, exact)", "data": "synthetic",
            "config": {"workload": "cfg3: fused grayscale->Sobel, 1024 x 3840x2160 RGB frames sharded over the GPUs, no collective",
                       "frames_total": args.frames, "frames_per_gpu": nf, "width": W, "height": H, "channels": CH, "luma_mode": "strict",
                       "l2": "inputs larger than L2 (shard >= 3.1 GB read per step, 126 MB L2)", "parallelism": f"batch-shard x{dist.world}"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if extras:
            line["other_kernels"] = extras
        emit(line)


def other_kernels(pkg, L, peak):
    """device-resident throughput of the other kernels of the path, each on a working set larger than L2
    (explanatory numbers; the headline is the line's `value`)"""
    out = {}
    t = pkg.Timer()

    def timeit(fn, iters=10):
        for _ in range(3):
            fn()
        pkg.sync()
        t.start(0)
        for _ in range(iters):
            fn()
        t.stop(0)
        return t.ms() / iters * 1e-3

    def entry(px, bpp, sec):
        gbs = bpp * px / sec / 1e9
        return {"gpixel_s": px / sec / 1e9, "gbs": gbs, "frac_of_measured_peak": gbs / peak, "bytes_per_px": bpp}

    n = 128
    # cfg1 shape: in-place grayscale, 1920x1080 RGB x 512 frames (3.2 GB)
    n1, w1, h1 = 512, 1920, 1080
    b = pkg.DeviceBuffer(n1 * w1 * h1 * 3)
    pkg.synth_dev(b, b.nbytes, 1)
    out["gray_inplace_1920x1080x512"] = entry(n1 * w1 * h1, 6, timeit(lambda: L.ompimg_gray_inplace_dev(b.ptr, w1 * 3, w1, h1, 3, n1, w1 * h1 * 3, 0)))
    # RGB -> packed luma plane (north star (2): 3 B read + 1 B written per pixel), same frames
    lum = pkg.DeviceBuffer(n1 * w1 * h1)
    out["rgb_to_luma_1920x1080x512"] = entry(n1 * w1 * h1, 4, timeit(lambda: L.ompimg_rgb_to_luma_dev(b.ptr, w1 * 3, 3, lum.ptr, w1, w1, h1, n1, w1 * h1 * 3, w1 * h1, 0)))
    # a8: legacy (r+g+b)/3 grayscale with min/max, one 1920 x (1080*512) image
    mm = pkg.DeviceBuffer(64); g3 = pkg.DeviceBuffer(n1 * w1 * h1 * 3)
    out["gray_avg3_minmax_1920x552960"] = entry(n1 * w1 * h1, 6, timeit(lambda: L.ompimg_gray_avg3_dev(b.ptr, w1 * 3, g3.ptr, w1 * 3, h1 * n1, w1, mm.ptr, 0)))
    lum.free(); mm.free(); g3.free()
    b.free()
    # cfg2 shape: Sobel plane->plane, 2000x3000 (pitch 2048) x 128 planes
    w2, h2, p2 = 2000, 3000, 2048
    s = pkg.DeviceBuffer(n * p2 * h2); d = pkg.DeviceBuffer(n * p2 * h2)
    pkg.synth_dev(s, s.nbytes, 2); d.fill(0)
    out["sobel_plane_2000x3000x128"] = entry(n * w2 * h2, 2, timeit(lambda: L.ompimg_sobel_dev(s.ptr, p2, d.ptr, p2, w2, h2, n, p2 * h2, p2 * h2, 0)))
    # same kernel on 3840x2160 planes
    w3, h3 = 3840, 2160
    n3 = min(n, s.nbytes // (w3 * h3))
    out[f"sobel_plane_3840x2160x{n3}"] = entry(n3 * w3 * h3, 2, timeit(lambda: L.ompimg_sobel_dev(s.ptr, w3, d.ptr, w3, w3, h3, n3, w3 * h3, w3 * h3, 0)))
    s.free(); d.free()
    # a5 semantics: RGB in place -> RGB edge (fused kernel + replicate), 3840x2160 x 64
    n5 = 64
    img = pkg.DeviceBuffer(n5 * PX * 3); e = pkg.DeviceBuffer(n5 * PX)
    pkg.synth_dev(img, img.nbytes, 3); e.fill(0)
    out["gray_sobel_pass_rgb_inplace_3840x2160x64"] = entry(n5 * PX, 6, timeit(lambda: L.ompimg_gray_sobel_pass_dev(img.ptr, W * 3, 3, e.ptr, W, W, H, n5, PX * 3, PX, 0)))
    img.free(); e.free()
    # cfg4: 8192x8192 RGB box / binomial blur (uint8 out)
    rows = cols = 8192
    for K in (5, 9):
        pr, pc = rows + K - 1, cols + K - 1
        src = pkg.DeviceBuffer(pr * pc * 3); dst = pkg.DeviceBuffer(rows * cols * 3)
        pkg.synth_dev(src, src.nbytes, 4)
        taps = np.ones(K * K, np.int32)
        out[f"conv_box{K}x{K}_8192x8192"] = entry(rows * cols, 6, timeit(lambda: L.ompimg_conv_u8_dev(src.ptr, pc * 3, rows, cols, K, taps.ctypes.data, K * K, dst.ptr, cols * 3, 0), 5))
        bn = np.array([1.0])
        for _ in range(K - 1):
            bn = np.convolve(bn, [1, 1])
        g = np.outer(bn, bn).astype(np.int32).ravel().copy()
        out[f"conv_binomial{K}x{K}_8192x8192"] = entry(rows * cols, 6, timeit(lambda: L.ompimg_conv_u8_dev(src.ptr, pc * 3, rows, cols, K, g.ctypes.data, int(g.sum()), dst.ptr, cols * 3, 0), 5))
        src.free(); dst.free()
    # a7: channel sums, 8192x8192
    src = pkg.DeviceBuffer(rows * cols * 3); sums = pkg.DeviceBuffer(64)
    pkg.synth_dev(src, src.nbytes, 5); sums.fill(0)
    out["channel_sums_8192x8192"] = entry(rows * cols, 3, timeit(lambda: L.ompimg_channel_sums_dev(src.ptr, cols * 3, rows, cols, sums.ptr, 0)))
    src.free(); sums.free()
    pkg.sync()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=TOTAL_FRAMES, help="frames in the whole job (BASELINE config 3: 1024)")
    ap.add_argument("--e2e-frames", type=int, default=64, help="frames per step per GPU on the end-to-end leg")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-frames", type=int, default=16, help="4K frames per CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--backend", default="nccl")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) go to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    dist = Dist(None if args.impl == "reference" else args.backend)      # the reference arm runs on rank 0 alone: no process group
    try:
        if args.impl == "reference":
            run_reference(args, dist)
        else:
            run_ours(args, dist)
    finally:
        dist.close()


if __name__ == "__main__":
    main()
