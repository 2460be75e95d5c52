#!/usr/bin/env python
"""tools/bench_bands.py -- BASELINE config 5: one 32768x32768 synthetic RGB image, fused gray->Sobel, split
into row bands over the GPUs of one box.

    python tools/bench_bands.py --gpus N                      one process, N devices: the kernel's producer
                                                              warp reads the halo rows from the peer GPU (NVLink)
    torchrun --nproc-per-node N tools/bench_bands.py --mp     one process per GPU: halo rows by NCCL send/recv

Prints one JSON line.  The checksum (sum of all interior edge bytes) is independent of N: compare it
with the --gpus 1 run to check that the decomposition is bit-identical.
"""
import argparse
import importlib
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
SEED = 0xB2000005


def single_process(a):
    pkg = importlib.import_module("openmp-parallel-computing_b200")
    pkg.init(0)
    b = pkg.Bands(a.width, a.height, 3, list(range(a.gpus)))
    b.synth(SEED)
    for _ in range(3):
        b.rgb_to_edge()
    ms = [b.rgb_to_edge() for _ in range(a.iters)]          # per-device CUDA events, max over devices
    best, med = min(ms), sorted(ms)[len(ms) // 2]
    px = a.width * a.height
    print(json.dumps({"workload": f"cfg5 {a.width}x{a.height} RGB fused gray->Sobel, row bands, peer-load halo", "mode": "one process, peer access",
                      "n_gpus": a.gpus, "ms_median": med, "ms_best": best, "gpixel_s": px / med / 1e6, "gbs_algorithmic": 4 * px / med / 1e6,
                      "checksum": b.checksum()}))
    b.close()


def multi_process(a):
    import torch
    import torch.distributed as td
    pkg = importlib.import_module("openmp-parallel-computing_b200")
    bmp = importlib.import_module("openmp-parallel-computing_b200.bands_mp")
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl")
    pkg.init(local)
    W, H = a.width, a.height
    r0, nr = bmp.band_rows(H, world)[rank]
    pitch, epitch = W * 3, W
    assert pitch % 128 == 0
    dev = torch.device("cuda", local)
    band = torch.empty((nr, pitch), dtype=torch.uint8, device=dev)
    edge = torch.zeros((nr, epitch), dtype=torch.uint8, device=dev)
    L = pkg.lib()
    assert L.ompimg_synth_fill_dev(band.data_ptr(), nr * pitch, SEED, r0 * pitch, 0) == 0
    pkg.sync()

    def step():
        keep = bmp.rgb_to_edge_band(pkg, td, torch, band, edge, W, 3, rank, world)
        pkg.sync(0)
        return keep

    for _ in range(3):
        step()
    times = []
    for _ in range(a.iters):
        td.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()                                                # halo exchange + kernel, synchronised
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        times.append(float(t.item()) * 1e3)
    y0, y1 = max(r0, 1), min(r0 + nr, H - 1)
    s = edge[y0 - r0:y1 - r0, 1:W - 1].to(torch.int64).sum() if y1 > y0 else torch.zeros((), dtype=torch.int64, device=dev)
    td.all_reduce(s)
    if rank == 0:
        med = sorted(times)[len(times) // 2]
        px = W * H
        print(json.dumps({"workload": f"cfg5 {W}x{H} RGB fused gray->Sobel, row bands, NCCL send/recv halo", "mode": "one process per GPU",
                          "n_gpus": world, "ms_median": med, "ms_best": min(times), "gpixel_s": px / med / 1e6, "gbs_algorithmic": 4 * px / med / 1e6,
                          "checksum": int(s.item()), "timing": "host wall clock incl. halo exchange, max over ranks"}))
    td.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--width", type=int, default=32768)
    ap.add_argument("--height", type=int, default=32768)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--mp", action="store_true")
    args = ap.parse_args()
    multi_process(args) if args.mp else single_process(args)
