"""CPU tests of the multi-rank host logic with world_size 2 and 3 over gloo: the batch sharding of
bench.py (no frame lost, none duplicated, checksum of checksums equal to the single-rank result) and
the row-band halo exchange protocol of bands_mp.py (per-band results with exchanged halo rows equal the
whole-image result).  The pixel math in these tests is the CPU oracle -- they test the plumbing, not
the kernels (the kernels' band mode is covered by the GPU tests)."""
import importlib
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as td
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as O
    import bench
    bmp = importlib.import_module("openmp-parallel-computing_b200.bands_mp")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist = bench.Dist("gloo")
    try:
        # ---- batch sharding (BASELINE config 3): frames are independent, contiguous block per rank
        total, h, w = 11, 40, 56
        f0, nf = bench.shard(total, rank, world)
        frames = O.synth(total * h * w * 3, 0xB2000003).reshape(total, h, w, 3)
        mine = np.array([int(O.rgb_to_edge(frames[f]).astype(np.int64).sum()) * (f + 1) for f in range(f0, f0 + nf)], np.float64)
        covered = dist.sum(float(nf))
        checksum = dist.sum(float(mine.sum()))
        slowest = dist.max(float(rank + 1))
        dist.barrier()
        # ---- row bands (BASELINE config 5): halo rows travel by send/recv, then every band is computed on its own
        H, W = 97, 130
        img = O.synth(H * W * 3, 0xB2000005).reshape(H, W, 3)
        r0, nr = bmp.band_rows(H, world)[rank]
        band = torch.from_numpy(img[r0:r0 + nr].reshape(nr, W * 3).copy())
        above, below = torch.zeros(W * 3, dtype=torch.uint8), torch.zeros(W * 3, dtype=torch.uint8)
        has_a, has_b = bmp.exchange_halo_rows(td, band[0], band[nr - 1], above, below, rank, world)
        rows = ([above.numpy()] if has_a else []) + [band.numpy()] + ([below.numpy()] if has_b else [])
        ext = np.ascontiguousarray(np.vstack(rows)).reshape(-1, W, 3)
        e = O.rgb_to_edge(ext)                                   # interior rows of the extended band
        lo = 1 if has_a else 0
        np.save(os.path.join(out_dir, f"band_{rank}.npy"), e[lo:lo + nr])
        np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([covered, checksum, slowest, r0, nr, int(has_a), int(has_b)], np.float64))
    finally:
        dist.close()


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_and_halo_exchange_over_gloo(world, tmp_path, O):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT))
    total, h, w = 11, 40, 56
    frames = O.synth(total * h * w * 3, 0xB2000003).reshape(total, h, w, 3)
    want = float(sum(int(O.rgb_to_edge(frames[f]).astype(np.int64).sum()) * (f + 1) for f in range(total)))
    H, W = 97, 130
    whole = O.rgb_to_edge(O.synth(H * W * 3, 0xB2000005).reshape(H, W, 3))
    got = np.zeros((H, W), np.uint8)
    for r in range(world):
        covered, checksum, slowest, r0, nr, has_a, has_b = np.load(tmp_path / f"meta_{r}.npy")
        assert covered == total and checksum == want and slowest == world
        assert bool(has_a) == (r > 0) and bool(has_b) == (r + 1 < world)
        band = np.load(tmp_path / f"band_{r}.npy")
        r0, nr = int(r0), int(nr)
        ys = slice(max(r0, 1), min(r0 + nr, H - 1))              # rows the band is responsible for (image border rows excluded)
        got[ys] = band[ys.start - r0:ys.stop - r0]
    assert np.array_equal(got[1:-1, 1:-1], whole[1:-1, 1:-1])


def test_shard_covers_every_frame_once():
    sys.path.insert(0, str(ROOT))
    import bench
    for total in (1, 7, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            blocks = [bench.shard(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == total
            for (a, n), (b, _) in zip(blocks, blocks[1:]):
                assert a + n == b
            assert max(n for _, n in blocks) - min(n for _, n in blocks) <= 1
