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
H2D and D2H inside the timed region.  The line also carries

    bands          BASELINE config 5: one synthetic 32768x32768 RGB image split into row bands over the N
                   GPUs (one process per GPU, the neighbours' bands mapped by CUDA IPC through the C ABI, the
                   stencil kernel fetching its halo rows from the peer GPU over NVLink), checked against
                   oracle-derived known answers (tests/golden/bands_cfg5.json) and, on the seams, byte for
                   byte against the CPU oracle
    cpu_baseline   the reference's OpenMP code on the box's host cores (1 thread and all), N=1 only
    other_kernels  the other kernels of the path, each with its CPU partner, N=1 only

torch is used only for the multi-rank barrier / max-over-ranks / the exchange of 64-byte IPC handles.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import importlib
import json
import os
import subprocess
import sys
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
BANDS_W = BANDS_H = 32768   # BASELINE config 5
BANDS_SEED = 0xB2000005


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
        self.cpu_group = None
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

    def host_barrier(self):
        """a barrier that keeps the GPUs idle while ranks wait (gloo): used around legs in which one rank drives
        several devices, where an NCCL kernel spinning on a waiting rank's GPU would compete with the measured work"""
        if not self.td:
            return
        if self.backend == "gloo":
            return self.td.barrier()
        if self.cpu_group is None:
            self.cpu_group = self.td.new_group(backend="gloo")
        self.td.barrier(group=self.cpu_group)

    def _reduce(self, v: float, op) -> float:
        if not self.td:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self._dev())
        self.td.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v: float) -> float:
        return self._reduce(v, self.td.ReduceOp.MAX) if self.td else v

    def sum(self, v: float) -> float:
        return self._reduce(v, self.td.ReduceOp.SUM) if self.td else v

    def gather(self, obj):
        """every rank's object, indexed by rank"""
        if not self.td:
            return [obj]
        out = [None] * self.world
        self.td.all_gather_object(out, obj)
        return out

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
        sm, mx, reasons, power = [], 0, set(), []
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            try:
                power.append(float(f[2]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # "under load": the upper half of the samples (the sampler also sees the idle edges of the region)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_mhz_min": (sm[0] if sm else None), "sm_max_mhz": mx or None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(nframes: int):
    """DRAM bytes of THIS launch (nframes 4K frames) from the committed `ncu --set full` capture of the same kernel,
    which records bytes per frame (profiles/traffic.json); None without a capture."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        per_frame = json.loads(p.read_text())["fused_rgb_to_edge_3840x2160"]["dram_bytes_per_frame"]
        return per_frame * nframes
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ CPU legs
# Everything below this line up to "GPU arm" runs the reference's code on the host cores: it is the baseline and
# the checker, never part of the product path.  The OpenMP thread count is fixed when libgomp loads, so every
# thread count runs in its own subprocess with OMP_NUM_THREADS pinned (torchrun exports OMP_NUM_THREADS=1).
def _oracle():
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as O
    O.build()
    return O


def cpu_pass_factory():
    """One iteration of main_with_sobel.c:51-74 on the host cores.  Uses the UNMODIFIED reference objects
    (oracle/_ref/libref_strict.so: convert_to_grayscale, sobel_edge) when they were built, else the oracle port."""
    O = _oracle()
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


def _best(fn, reps):
    best = None
    for _ in range(reps):
        dt = fn()
        best = dt if best is None or dt < best else best
    return best


def cpu_job(job: str, **kw) -> dict:
    """runs INSIDE the pinned-thread subprocess; returns a small dict"""
    if job == "pass":          # cfg3: the main_with_sobel.c pass over `frames` 4K frames
        run, kind, O = cpu_pass_factory()
        frames = kw["frames"]
        src = O.synth(frames * PX * CH, SEED).reshape(frames, H, W, CH)
        gray, edge = np.empty((H, W), np.uint8), np.zeros((H, W), np.uint8)
        work = src.copy()

        def step():
            np.copyto(work, src)
            t0 = time.perf_counter()
            for f in range(frames):
                run(work[f], gray, edge, W, H, CH)
            return time.perf_counter() - t0
        times = [step() for _ in range(kw.get("warmup", 0) + kw["steps"])][kw.get("warmup", 0):]
        # frame 0's edge plane, for the GPU end-to-end leg to be compared with (the scratch plane keeps a zero border)
        run(src[0].copy(), gray, edge, W, H, CH)
        return {"gpx_s_best": frames * PX / min(times) / 1e9, "gpx_s_mean": frames * PX * len(times) / sum(times) / 1e9, "ms_per_step": sum(times) / len(times) * 1e3,
                "kind": kind, "threads": O.omp_max_threads(), "frame0_edge_sha256": hashlib.sha256(edge.tobytes()).hexdigest()}
    if job == "gray":          # cfg1: convert_to_grayscale on the 1920x1080 test image (real pixels when the decoded fixture travelled)
        O = _oracle()
        w, h = 1920, 1080
        fx, gold = ROOT / "oracle" / "_ref" / "fixtures" / "test_rgb.raw", ROOT / "oracle" / "_ref" / "fixtures" / "test_gray_golden.raw"
        real = fx.exists() and fx.stat().st_size == w * h * 3
        src = np.fromfile(fx, np.uint8).reshape(h, w, 3) if real else O.synth(w * h * 3, 0xB2000001).reshape(h, w, 3)
        ref = O.Ref("strict") if O.have_ref("strict") else None
        work = src.copy()

        def step():
            np.copyto(work, src)
            t0 = time.perf_counter()
            ref.gray(work) if ref else O.gray_inplace(work, 3)
            return time.perf_counter() - t0
        best = _best(step, kw.get("reps", 20))
        exact = None
        if real and gold.exists():
            exact = bool(np.array_equal(work, np.fromfile(gold, np.uint8).reshape(h, w, 3)))     # images/test_gray.png, decoded
        return {"gpx_s_best": w * h / best / 1e9, "kind": "reference" if ref else "port", "threads": O.omp_max_threads(),
                "image": "images/test.jpg (decoded)" if real else "synthetic 1920x1080", "equals_test_gray_png": exact, "out_sha256": hashlib.sha256(work.tobytes()).hexdigest()}
    if job == "sobel":         # cfg2: sobel_edge on a 2000x3000 plane
        O = _oracle()
        w, h = 2000, 3000
        src = O.synth(w * h, 0xB2000002).reshape(h, w)
        dst = np.zeros((h, w), np.uint8)
        ref = O.Ref("strict") if O.have_ref("strict") else None

        def step():
            t0 = time.perf_counter()
            ref.sobel(src, dst) if ref else O.sobel(src, dst)
            return time.perf_counter() - t0
        best = _best(step, kw.get("reps", 8))
        return {"gpx_s_best": w * h / best / 1e9, "kind": "reference" if ref else "port", "threads": O.omp_max_threads(), "out_sha256": hashlib.sha256(dst.tobytes()).hexdigest()}
    if job == "conv":          # cfg4 at reduced size: old/parallel_convolution.c, `long` layout, K x K box
        O = _oracle()
        K, n = kw["K"], 1024
        padded = O.synth((n + K - 1) * (n + K - 1) * 3, 0xB2000004).reshape(n + K - 1, n + K - 1, 3).astype(np.int64)
        taps = np.ones((K, K), np.int64)
        out = np.zeros((n, n, 3), np.int64)
        ref = O.RefOld(n, n, K, K * K) if O.have_ref_old(n, n, K, K * K) else None

        def step():
            out[:] = 0
            t0 = time.perf_counter()
            ref.convolution(padded, taps, out) if ref else O.convolution_long(padded, taps, K * K, out)
            return time.perf_counter() - t0
        best = _best(step, kw.get("reps", 3))
        return {"gpx_s_best": n * n / best / 1e9, "kind": "reference" if ref else "port", "threads": O.omp_max_threads(),
                "sample": f"{n}x{n} RGB in the reference's `long` layout (reduced from 8192x8192: the layout is 8x larger), {K}x{K} box"}
    if job == "seams":         # checker of the bands leg: edge rows either side of the band seams, byte for byte
        O = _oracle()
        d = np.load(kw["npz"])
        rows, got = d["rows"], d["edge"]
        w, h = int(d["width"]), int(d["height"])
        bad = []
        for i, y in enumerate(rows.tolist()):
            src = O.synth(3 * w * 3, BANDS_SEED, (y - 1) * w * 3).reshape(3, w, 3)
            want = O.rgb_to_edge(src)[1]
            if not np.array_equal(got[i, 1:w - 1], want[1:w - 1]):
                bad.append(y)
        return {"rows_checked": len(rows), "rows_differing": bad}
    raise ValueError(job)


def cpu_subprocess(job: str, threads: int, timeout: int = 900, **kw) -> dict:
    code = (f"import sys, json; sys.path.insert(0, {str(ROOT)!r}); import bench; "
            f"print('@@' + json.dumps(bench.cpu_job({job!r}, **{kw!r})))")
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="false", OMP_DYNAMIC="false")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"cpu job {job} failed: " + r.stderr[-2000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("@@")][-1]
    return json.loads(line[2:])


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, dist: Dist):
    """--impl reference: the reference's CPU implementation of the same pass with all the host threads it can use"""
    if dist.rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = args.cpu_frames
    r = cpu_subprocess("pass", cores, frames=frames, steps=args.steps, warmup=args.warmup)
    value = r["gpx_s_mean"]
    sample = (f"{frames} synthetic 3840x2160 RGB frames per step through the main_with_sobel.c pass ({r['kind']}), "
              f"omp_get_max_threads() = {r['threads']} in a subprocess with OMP_NUM_THREADS={cores}")
    line = {
        "impl": "reference", "metric": "Gpixel/s", "value": value, "unit": "Gpixel/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u8 (fp32 luma)", "data": "synthetic",
        "config": {"workload": "cfg3: fused grayscale->Sobel, 1024 x 3840x2160 RGB frames (bounded CPU sample per step)", "frames_per_step": frames},
        "cpu_baseline": {"value": value, "unit": "Gpixel/s", "cores": r["threads"], "kind": r["kind"], "sample": sample},
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


def ck(pkg, rc, what=""):
    if rc:
        raise pkg.OmpimgError(f"{what}: " + pkg.lib().ompimg_last_error().decode())


def e2e_leg(args, dist, pkg, L, src, f0, nf, edge0):
    """end to end through the public host API: pinned host frames -> H2D -> kernel -> D2H edge planes, then the same
    transfers without the kernel (the platform's ceiling for this traffic pattern on this box at this N)"""
    fb, eb = PX * CH, PX
    ne = min(args.e2e_frames, nf)
    hin, hout = pkg.PinnedBuffer(ne * fb), pkg.PinnedBuffer(ne * eb)
    pkg.synth_dev(src, ne * fb, SEED, offset=f0 * fb)         # same bytes as the shard's first frames
    ck(pkg, L.ompimg_copy_d2h(hin.ptr, src.ptr, ne * fb, 0), "copy_d2h")
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
    frame0_sha = hashlib.sha256(edges_host[0].tobytes()).hexdigest() if f0 == 0 else None
    # ---- the ceiling: the same bytes over the same 8 streams, no kernel in between
    ns = 8
    dsl = [pkg.DeviceBuffer(fb) for _ in range(ns)]
    esl = [pkg.DeviceBuffer(eb) for _ in range(ns)]

    def copies_only():
        for f in range(ne):
            s = f % ns
            L.ompimg_copy_h2d(dsl[s].ptr, hin.ptr + f * fb, fb, s)
            L.ompimg_copy_d2h(hout.ptr + f * eb, esl[s].ptr, eb, s)
        pkg.sync()
    copies_only()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        copies_only()
    dtc = dist.max(time.perf_counter() - t0)
    for b in dsl + esl:
        b.free()
    value = dist.world * ne * PX * args.e2e_steps / dt / 1e9
    ceiling = dist.world * ne * PX * args.e2e_steps / dtc / 1e9
    e2e = {"value": value, "unit": "Gpixel/s", "h2d_bytes_per_step": ne * fb, "d2h_bytes_per_step": ne * eb,
           "frames_per_step_per_gpu": ne, "steps": args.e2e_steps, "api": "ompimg_rgb_to_edge_batch_host (pinned host buffers, 8 streams)",
           "timing": "host wall clock around the synchronous call, max over ranks",
           "h2d_gbs_whole_box": dist.world * ne * fb * args.e2e_steps / dt / 1e9, "d2h_gbs_whole_box": dist.world * ne * eb * args.e2e_steps / dt / 1e9,
           "platform_ceiling": {"value": ceiling, "unit": "Gpixel/s", "h2d_gbs_whole_box": dist.world * ne * fb * args.e2e_steps / dtc / 1e9,
                                "how": "the same H2D + D2H copies (ompimg_copy_h2d / ompimg_copy_d2h, same buffers, 8 streams, all ranks at once) with no kernel"},
           "frac_of_ceiling": value / ceiling}
    hin.free(); hout.free()
    return e2e, frame0_sha


def bands_leg(args, dist, pkg, L, peak):
    """BASELINE config 5: one 32768x32768 synthetic RGB image, fused gray->Sobel, one row band per GPU.  One process per
    GPU; every rank maps its neighbours' bands (CUDA IPC through the C ABI) and the kernel reads the two halo rows out of
    the peer GPU's memory over NVLink.  Known answers: tests/golden/bands_cfg5.json (derived from the CPU oracle alone)."""
    bmp = importlib.import_module("openmp-parallel-computing_b200.bands_mp")
    Wb, Hb = args.bands_size, args.bands_size
    rows_of = bmp.band_rows(Hb, dist.world)
    r0, nr = rows_of[dist.rank]
    pitch, epitch = Wb * 3, Wb
    assert pitch % 128 == 0
    band = pkg.DeviceBuffer(nr * pitch)
    edge = pkg.DeviceBuffer(nr * epitch)
    edge.fill(0)
    pkg.synth_dev(band, nr * pitch, BANDS_SEED, offset=r0 * pitch)
    pkg.sync()
    nb = bmp.NeighbourBands(pkg, dist.gather, band.ptr, pitch, rows_of, dist.rank, dist.world) if dist.world > 1 else None
    above = nb.row_above if nb else None
    below = nb.row_below if nb else None
    dist.barrier()                                            # every band is filled before anybody reads a halo row

    def step():
        ck(pkg, L.ompimg_rgb_to_edge_band_dev(band.ptr, pitch, 3, edge.ptr, epitch, Wb, nr, above, below, 0), "rgb_to_edge_band_dev")
    for _ in range(3):
        step()
    pkg.sync()
    t = pkg.Timer()
    dist.barrier()
    t.start(0)
    for _ in range(args.bands_steps):
        step()
    t.stop(0)
    ms = dist.max(t.ms() / args.bands_steps)
    pkg.sync()
    dist.barrier()
    # ---- known answers: interior checksum (whole image), rows either side of the seams
    y0, y1 = max(r0, 1), min(r0 + nr, Hb - 1)
    rows_local = sorted({y for y in (r0, r0 + 1, r0 + nr - 2, r0 + nr - 1, 1, 2, Hb - 3, Hb - 2) if y0 <= y < y1})
    seam_rows = np.stack([edge.download(epitch, offset=(y - r0) * epitch) for y in rows_local]) if rows_local else np.zeros((0, epitch), np.uint8)
    sums = pkg.DeviceBuffer(8)
    sums.fill(0)
    ck(pkg, L.ompimg_byte_sum_dev(edge.ptr + (y0 - r0) * epitch + 1, epitch, Wb - 2, y1 - y0, sums.ptr, 0), "byte_sum")
    checksum = sum(dist.gather(int(np.frombuffer(sums.download(8).tobytes(), np.uint64)[0])))
    gathered = dist.gather((rows_local, seam_rows))
    if nb:
        nb.close()
    band.free(); edge.free(); sums.free()
    # ---- the same image on ONE GPU (rank 0 alone), for the efficiency of the split
    n1_ms = None
    if dist.world > 1:
        if dist.rank == 0:
            b1, e1 = pkg.DeviceBuffer(Hb * pitch), pkg.DeviceBuffer(Hb * epitch)
            pkg.synth_dev(b1, Hb * pitch, BANDS_SEED)
            for _ in range(2):
                ck(pkg, L.ompimg_rgb_to_edge_dev(b1.ptr, pitch, 3, e1.ptr, epitch, Wb, Hb, 1, 0, 0, 0), "rgb_to_edge_dev")
            pkg.sync()
            t.start(0)
            for _ in range(args.bands_steps):
                ck(pkg, L.ompimg_rgb_to_edge_dev(b1.ptr, pitch, 3, e1.ptr, epitch, Wb, Hb, 1, 0, 0, 0), "rgb_to_edge_dev")
            t.stop(0)
            n1_ms = t.ms() / args.bands_steps
            b1.free(); e1.free()
        dist.host_barrier()
    if dist.rank != 0:
        return None
    px = Wb * Hb
    rec = {"workload": f"cfg5: fused grayscale->Sobel of one synthetic {Wb}x{Hb} RGB image, one row band per GPU",
           "halo": "none (one GPU)" if dist.world == 1 else "1 RGB row each side, read by the stencil kernel's bulk copies straight from the neighbour GPU's band (CUDA IPC mapping, NVLink); no staging copy, no collective",
           "processes": dist.world, "rows_per_gpu": nr, "steps": args.bands_steps, "ms": ms, "gpixel_s": px / ms / 1e6,
           "gbs_algorithmic_per_gpu": 4.0 * px / dist.world / ms / 1e6, "frac_of_measured_peak_per_gpu": 4.0 * px / dist.world / ms / 1e6 / peak,
           "ms_one_gpu_same_run": n1_ms, "efficiency_vs_n1": (n1_ms / (dist.world * ms)) if n1_ms else (1.0 if dist.world == 1 else None),
           "checksum": checksum, "timing": "CUDA events around the band's launches on its stream, max over ranks"}
    gold_p = ROOT / "tests" / "golden" / "bands_cfg5.json"
    if gold_p.exists() and Wb == BANDS_W:
        gold = json.loads(gold_p.read_text())
        rec["checksum_expected"] = gold["interior_sum"]
        rec["checksum_ok"] = checksum == gold["interior_sum"]
        bad = []
        nchecked = 0
        for rows_r, edge_r in gathered:
            for i, y in enumerate(rows_r):
                if str(y) in gold["row_sums"]:
                    nchecked += 1
                    if int(edge_r[i, 1:Wb - 1].astype(np.int64).sum()) != gold["row_sums"][str(y)]:
                        bad.append(y)
        rec["seam_row_sums"] = {"rows_checked": nchecked, "rows_differing": bad, "source": "tests/golden/bands_cfg5.json (CPU oracle)"}
        assert rec["checksum_ok"] and not bad, rec
    if not args.no_cpu:
        # byte-for-byte against the CPU oracle on the rows either side of every seam (the checker, in its own process)
        rows_all = [y for rows_r, _ in gathered for y in rows_r]
        edge_all = np.concatenate([e for _, e in gathered if len(e)]) if rows_all else np.zeros((0, epitch), np.uint8)
        npz = f"/tmp/ompimg_bench_seams_{os.getpid()}.npz"
        np.savez(npz, rows=np.array(rows_all, np.int64), edge=edge_all, width=Wb, height=Hb)
        try:
            rec["seams_vs_oracle"] = cpu_subprocess("seams", host_cores(), npz=npz)
        finally:
            os.unlink(npz)
        assert not rec["seams_vs_oracle"]["rows_differing"], rec
    return rec


def bands_single_process_leg(args, dist, pkg, peak):
    """the same image with ONE process driving all N GPUs (ompimg_bands_*: one launching thread per device, halo rows by
    bulk copies out of the peer allocation).  Rank 0 runs it while the other ranks wait on a host-side barrier."""
    rec = None
    if dist.rank == 0:
        try:
            b = pkg.Bands(args.bands_size, args.bands_size, 3, list(range(dist.world)))
            b.synth(BANDS_SEED)
            for _ in range(3):
                b.rgb_to_edge()
            ms = sorted(b.rgb_to_edge() for _ in range(args.bands_steps))
            px = args.bands_size ** 2
            rec = {"ms_median": ms[len(ms) // 2], "ms_best": ms[0], "gpixel_s": px / ms[len(ms) // 2] / 1e6, "checksum": b.checksum(),
                   "mode": f"one process, {dist.world} devices, peer access",
                   "timing": "per-device CUDA events around each band's launch, max over devices, median of the steps"}
            b.close()
        except Exception as e:       # a second context per GPU may not fit beside the ranks' own: report, do not fail the line
            rec = {"error": repr(e)[:300]}
    dist.host_barrier()
    return rec


def bands_old_kernels_leg(args, dist, pkg, peak):
    """row-band forms of the old/ kernels on the cfg5 image (old/parallel_convolution.c:12-23 as a 5x5 box blur with the image as the
    pre-padded input: K-1 halo rows per band by P2P copies; old/parallel_avg_pixel.c:17-30: per-GPU partial channel sums added on
    the host).  One process (rank 0) drives all N GPUs; results are checked against constants derived from the CPU oracle
    (tools/make_bands_checksum.py --extras -> tests/golden/bands_cfg5.json)."""
    rec = None
    if dist.rank == 0:
        try:
            gold = {}
            gp = ROOT / "tests" / "golden" / "bands_cfg5.json"
            if gp.exists() and args.bands_size == BANDS_W:
                gold = json.loads(gp.read_text())
            n = args.bands_size
            b = pkg.Bands(n, n, 3, list(range(dist.world)))
            b.synth(BANDS_SEED)
            sums, _ = b.channel_sums()
            ms_s = sorted(b.channel_sums()[1] for _ in range(5))
            K = 5
            kern = np.ones((K, K), np.int32)
            for _ in range(2):
                b.convolution(kern, K * K)
            ms_c = sorted(b.convolution(kern, K * K) for _ in range(max(5, args.bands_steps // 2)))
            csum = b.convolution_checksum()
            opx = (n - K + 1) ** 2
            rec = {"devices": dist.world,
                   "channel_sums": {"ms_median": ms_s[len(ms_s) // 2], "gpixel_s": n * n / ms_s[len(ms_s) // 2] / 1e6, "sums": sums,
                                    "equals_oracle": (sums == gold["channel_sums"]) if "channel_sums" in gold else None},
                   "conv_box5x5": {"ms_median": ms_c[len(ms_c) // 2], "ms_best": ms_c[0], "gpixel_s": opx / ms_c[len(ms_c) // 2] / 1e6,
                                   "frac_of_measured_peak_per_gpu": 6 * opx / (ms_c[len(ms_c) // 2] * 1e-3) / 1e9 / dist.world / peak,
                                   "halo": "4 rows below each band, cudaMemcpyPeerAsync from the next GPU on the band's stream (inside the timed region)",
                                   "checksum": csum, "equals_oracle": (csum == gold["conv5_box_sum"]) if "conv5_box_sum" in gold else None},
                   "timing": "per-device CUDA events around each band's copies + launch, max over devices, median of the steps"}
            b.close()
        except Exception as e:
            rec = {"error": repr(e)[:300]}
    dist.host_barrier()
    return rec


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
        ck(pkg, L.ompimg_rgb_to_edge_dev(src.ptr, W * CH, CH, dst.ptr, W, W, H, nf, fb, eb, 0), "rgb_to_edge_dev")

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
    roofline = {"bound": "hbm", "kernel": "sobel_tma_kernel<3,0,0,0> (fused RGB->luma->Sobel; one warp per band x 512-pixel strip, 3-D tensor-map copies)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src, "traffic": ncu_traffic(nf),
                "algorithmic_bytes_per_launch": BYTES_PER_PX * nf * PX, "frac_of_nominal_8TBs": achieved / 8000.0,
                "launch": f"rank 0's launch: {nf} frames"}

    e2e, frame0_sha = e2e_leg(args, dist, pkg, L, src, f0, nf, edge0)
    e2e["numa_node_rank0"] = numa
    src.free(); dst.free()

    bands = None
    if not args.no_bands:
        bands = bands_leg(args, dist, pkg, L, peak)
        if dist.world > 1 and not args.no_bands_single:
            sp = bands_single_process_leg(args, dist, pkg, peak)
            if bands is not None:
                bands["single_process"] = sp
        if not args.no_bands_single:
            ok = bands_old_kernels_leg(args, dist, pkg, peak)
            if bands is not None:
                bands["old_kernels"] = ok

    cpu = None
    if dist.rank == 0 and dist.world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        r_all = cpu_subprocess("pass", cores, frames=args.cpu_frames, steps=3)
        r_one = cpu_subprocess("pass", 1, frames=max(1, args.cpu_frames // 4), steps=2)
        cpu = {"value": r_all["gpx_s_best"], "unit": "Gpixel/s", "cores": r_all["threads"], "kind": r_all["kind"], "value_1_thread": r_one["gpx_s_best"],
               "sample": f"{args.cpu_frames} synthetic 3840x2160 RGB frames through the main_with_sobel.c pass, best of 3 (1 thread: {max(1, args.cpu_frames // 4)} frames, best of 2)",
               "e2e_frame0_equals_cpu": frame0_sha == r_all["frame0_edge_sha256"]}
        assert cpu["e2e_frame0_equals_cpu"], "the end-to-end leg's first edge plane differs from the CPU reference's"

    extras = None
    if dist.rank == 0 and dist.world == 1 and not args.no_extras:
        extras = other_kernels(pkg, L, peak, with_cpu=not args.no_cpu)

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
        if bands:
            line["bands"] = bands
        if extras:
            line["other_kernels"] = extras
        emit(line)


def other_kernels(pkg, L, peak, with_cpu=True):
    """device-resident throughput of the other kernels of the path, each on a working set larger than L2, each with the
    reference's CPU code timed beside it (1 thread and all cores; cfg1 also 2/4/6 as the reference's own bench does)
    (explanatory numbers; the headline is the line's `value`)"""
    out = {}
    t = pkg.Timer()
    cores = os.cpu_count() or 1

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

    def partner(job, threads_list, **kw):
        if not with_cpu:
            return None
        res = {}
        for th in threads_list:
            try:
                r = cpu_subprocess(job, th, **kw)
            except Exception as e:
                res[f"threads_{th}"] = {"error": repr(e)[:200]}
                continue
            d = {"gpixel_s": r["gpx_s_best"], "omp_threads": r["threads"], "kind": r["kind"]}
            for k in ("image", "equals_test_gray_png", "sample"):
                if r.get(k) is not None:
                    d[k] = r[k]
            res[f"threads_{th}"] = d
        return res

    # frame-level parity inside the bench: the first and last frame of each batch against the CPU oracle on the same synthetic bytes
    # (the oracle is the checker here, as in the cpu_baseline leg; full-size and edge-case parity is tests/)
    O = None
    if with_cpu:
        try:
            O = _oracle()
        except Exception:
            O = None

    def same(got, want):
        return bool(np.array_equal(np.asarray(got).reshape(-1), np.asarray(want).reshape(-1)))

    n = 128
    # cfg1 shape: in-place grayscale, 1920x1080 RGB x 512 frames (3.2 GB)
    n1, w1, h1 = 512, 1920, 1080
    b = pkg.DeviceBuffer(n1 * w1 * h1 * 3)
    pkg.synth_dev(b, b.nbytes, 1)
    out["gray_inplace_1920x1080x512"] = entry(n1 * w1 * h1, 6, timeit(lambda: L.ompimg_gray_inplace_dev(b.ptr, w1 * 3, w1, h1, 3, n1, w1 * h1 * 3, 0)))
    out["gray_inplace_1920x1080x512"]["cpu"] = partner("gray", sorted({1, 2, 4, 6, cores}))
    fb1 = w1 * h1 * 3
    gray_ref = {}
    if O:
        ok = True
        for f in (0, n1 - 1):                        # 13 in-place applications so far (3 warm-up + 10 timed)
            ref = O.synth(fb1, 1, f * fb1)
            for _ in range(13):
                O.gray_inplace(ref, 3)
            gray_ref[f] = ref
            ok = ok and same(b.download(fb1, offset=f * fb1), ref)
        out["gray_inplace_1920x1080x512"]["equals_oracle"] = ok
    # RGB -> packed luma plane (north star (2): 3 B read + 1 B written per pixel), same frames
    lum = pkg.DeviceBuffer(n1 * w1 * h1)
    out["rgb_to_luma_1920x1080x512"] = entry(n1 * w1 * h1, 4, timeit(lambda: L.ompimg_rgb_to_luma_dev(b.ptr, w1 * 3, 3, lum.ptr, w1, w1, h1, n1, w1 * h1 * 3, w1 * h1, 0)))
    # a8: legacy (r+g+b)/3 grayscale with min/max, one 1920 x (1080*512) image
    mm = pkg.DeviceBuffer(64); g3 = pkg.DeviceBuffer(n1 * w1 * h1 * 3)
    out["gray_avg3_minmax_1920x552960"] = entry(n1 * w1 * h1, 6, timeit(lambda: L.ompimg_gray_avg3_dev(b.ptr, w1 * 3, g3.ptr, w1 * 3, h1 * n1, w1, mm.ptr, 0)))
    if O:
        f = n1 - 1
        img_f = gray_ref[f].reshape(h1, w1, 3)
        out["rgb_to_luma_1920x1080x512"]["equals_oracle"] = same(lum.download(w1 * h1, offset=f * w1 * h1), O.rgb_to_luma(img_f))
        out["gray_avg3_minmax_1920x552960"]["equals_oracle"] = same(g3.download(fb1, offset=f * fb1), O.gray_avg3_u8(img_f)[0])
    lum.free(); mm.free(); g3.free()
    b.free()
    # cfg2 shape: Sobel plane->plane, 2000x3000 (pitch 2048) x 128 planes
    w2, h2, p2 = 2000, 3000, 2048
    s = pkg.DeviceBuffer(n * p2 * h2); d = pkg.DeviceBuffer(n * p2 * h2)
    pkg.synth_dev(s, s.nbytes, 2); d.fill(0)
    out["sobel_plane_2000x3000x128"] = entry(n * w2 * h2, 2, timeit(lambda: L.ompimg_sobel_dev(s.ptr, p2, d.ptr, p2, w2, h2, n, p2 * h2, p2 * h2, 0)))
    out["sobel_plane_2000x3000x128"]["cpu"] = partner("sobel", sorted({1, cores}))
    if O:
        ok = True
        for f in (0, n - 1):
            src_f = O.synth(p2 * h2, 2, f * p2 * h2).reshape(h2, p2)[:, :w2]
            got = d.download(p2 * h2, offset=f * p2 * h2).reshape(h2, p2)[:, :w2]
            ok = ok and same(got, O.sobel(np.ascontiguousarray(src_f)))          # dst was zeroed: the untouched border is zero on both sides
        out["sobel_plane_2000x3000x128"]["equals_oracle"] = ok
    # same kernel on 3840x2160 planes
    w3, h3 = 3840, 2160
    n3 = min(n, s.nbytes // (w3 * h3))
    d.fill(0)
    out[f"sobel_plane_3840x2160x{n3}"] = entry(n3 * w3 * h3, 2, timeit(lambda: L.ompimg_sobel_dev(s.ptr, w3, d.ptr, w3, w3, h3, n3, w3 * h3, w3 * h3, 0)))
    if O:
        f = n3 - 1
        src_f = O.synth(w3 * h3, 2, f * w3 * h3).reshape(h3, w3)
        out[f"sobel_plane_3840x2160x{n3}"]["equals_oracle"] = same(d.download(w3 * h3, offset=f * w3 * h3), O.sobel(src_f))
    s.free(); d.free()
    # a5 semantics: RGB in place -> RGB edge, 3840x2160 x 64, the driver's loop of 10 passes per call (main_with_sobel.c:51-74 x passes):
    # one preparation kernel per call, then one fused launch per pass (each pass writes the next one's seam data)
    n5, npass = 64, 10
    img = pkg.DeviceBuffer(n5 * PX * 3); e = pkg.DeviceBuffer(n5 * PX)
    pkg.synth_dev(img, img.nbytes, 3); e.fill(0)
    l0 = pkg.launch_count()
    sec = timeit(lambda: L.ompimg_gray_sobel_passes_dev(img.ptr, W * 3, 3, e.ptr, W, W, H, n5, PX * 3, PX, npass, 0), 3) / npass
    out["gray_sobel_pass_rgb_inplace_3840x2160x64"] = entry(n5 * PX, 6, sec)
    out["gray_sobel_pass_rgb_inplace_3840x2160x64"]["launches_per_pass"] = (pkg.launch_count() - l0) / (6 * npass)
    out["gray_sobel_pass_rgb_inplace_3840x2160x64"]["passes_per_call"] = npass
    sec1 = timeit(lambda: L.ompimg_gray_sobel_pass_dev(img.ptr, W * 3, 3, e.ptr, W, W, H, n5, PX * 3, PX, 0))
    out["gray_sobel_pass_rgb_inplace_3840x2160x64"]["single_pass_call"] = entry(n5 * PX, 6, sec1)
    if O:
        f = n5 - 1                                   # 6 calls x 10 passes + 13 single-pass calls on the same image
        ref = O.synth(PX * 3, 3, f * PX * 3).reshape(H, W, 3)
        edge_ref = np.zeros((H, W), np.uint8)
        for _ in range(6 * npass + 13):
            ref, edge_ref = O.gray_sobel_pass(ref, edge_ref)
        out["gray_sobel_pass_rgb_inplace_3840x2160x64"]["equals_oracle"] = same(img.download(PX * 3, offset=f * PX * 3), ref)
    img.free(); e.free()
    # cfg4: 8192x8192 RGB box / binomial blur (uint8 out).  Device layout: rows 128-byte aligned like every device copy the host entry
    # points make (conv_tma.cuh: tensor-map copies); the reference's dense (cols+K-1)*3 pitch (conv_sep.cuh) is timed beside it.
    rows = cols = 8192
    gold = {}
    gp = ROOT / "tests" / "golden" / "conv_cfg4.json"          # output byte sums from the CPU oracle (tools/make_conv_checksums.py)
    if gp.exists():
        gold = json.loads(gp.read_text()).get("sums", {})
    csum = pkg.DeviceBuffer(64)

    def conv_check(which):
        """sum of the output bytes of the last launch against the oracle's constant for this kernel and layout (None without one)"""
        if which not in gold:
            return None
        csum.fill(0)
        ck(pkg, L.ompimg_byte_sum_dev(dst.ptr, cols * 3, cols * 3, rows, csum.ptr, 0), "byte_sum")
        pkg.sync()
        return int(np.frombuffer(csum.download(8).tobytes(), np.uint64)[0]) == gold[which]["out_byte_sum"]

    for K in (5, 9):
        pr, pc = rows + K - 1, cols + K - 1
        ip = (pc * 3 + 127) // 128 * 128
        src = pkg.DeviceBuffer(pr * ip); dst = pkg.DeviceBuffer(rows * cols * 3)
        pkg.synth_dev(src, src.nbytes, 4)
        taps = np.ones(K * K, np.int32)
        bn = np.array([1.0])
        for _ in range(K - 1):
            bn = np.convolve(bn, [1, 1])
        g = np.outer(bn, bn).astype(np.int32).ravel().copy()
        for name, tp, nm in (("box", taps, K * K), ("binomial", g, int(g.sum()))):
            key = f"conv_{name}{K}x{K}_8192x8192"
            out[key] = entry(rows * cols, 6, timeit(lambda: L.ompimg_conv_u8_dev(src.ptr, ip, rows, cols, K, tp.ctypes.data, nm, dst.ptr, cols * 3, 0), 10))
            out[key]["in_pitch"] = ip
            out[key]["equals_oracle"] = conv_check(f"{name}{K}x{K}_aligned128")
            out[key]["dense_pitch"] = entry(rows * cols, 6, timeit(lambda: L.ompimg_conv_u8_dev(src.ptr, pc * 3, rows, cols, K, tp.ctypes.data, nm, dst.ptr, cols * 3, 0), 10))
            out[key]["dense_pitch"]["in_pitch"] = pc * 3
            out[key]["dense_pitch"]["equals_oracle"] = conv_check(f"{name}{K}x{K}_dense")
        out[f"conv_box{K}x{K}_8192x8192"]["cpu"] = partner("conv", sorted({1, cores}), K=K)
        src.free(); dst.free()
    csum.free()
    # a7: channel sums, 8192x8192
    src = pkg.DeviceBuffer(rows * cols * 3); sums = pkg.DeviceBuffer(64)
    pkg.synth_dev(src, src.nbytes, 5); sums.fill(0)
    out["channel_sums_8192x8192"] = entry(rows * cols, 3, timeit(lambda: L.ompimg_channel_sums_dev(src.ptr, cols * 3, rows, cols, sums.ptr, 0)))
    if O:                                            # the 13 launches accumulated onto the same three sums
        ref = O.synth(rows * cols * 3, 5).reshape(-1, 3).sum(axis=0, dtype=np.uint64)
        got = np.frombuffer(sums.download(24).tobytes(), np.uint64)
        out["channel_sums_8192x8192"]["equals_oracle"] = bool((got == 13 * ref).all())
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
    ap.add_argument("--bands-size", type=int, default=BANDS_W, help="side of the cfg5 image")
    ap.add_argument("--bands-steps", type=int, default=20)
    ap.add_argument("--no-bands", action="store_true")
    ap.add_argument("--no-bands-single", action="store_true", help="skip the one-process-all-devices form of cfg5")
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
