"""Row bands with ONE PROCESS PER GPU (BASELINE config 5, the launcher's form).

CPU (not gpu): the side-channel protocol of bands_mp.NeighbourBands -- every rank's 64-byte handle reaches
every rank, each rank opens exactly its neighbours' handles and derives the halo-row addresses from the
neighbours' band heights -- against a fake C ABI, over gloo with world 2 and 3; and the bench's seam checker.

GPU: two (or more) real processes, each with its own CUDA context, export their bands with ompimg_ipc_export,
map the neighbour's band with ompimg_ipc_open and run ompimg_rgb_to_edge_band_dev with the mapped halo-row
addresses: the stencil kernel's bulk copies then read another process's memory (another GPU's, over NVLink,
when the box has several).  The assembled image must equal the CPU oracle byte for byte."""
import ctypes as C
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


class _FakeLib:
    """stands in for libompimg_b200.so on the CPU: 'exports' a handle naming (rank, base) and 'opens' it to base"""

    def __init__(self, rank):
        self.rank, self.opened, self.closed = rank, [], []

    def ompimg_ipc_export(self, ptr, buf):
        raw = f"{self.rank}:{ptr}".encode().ljust(64, b"\0")
        C.memmove(buf, raw, 64)
        return 0

    def ompimg_ipc_open(self, buf):
        r, base = bytes(buf).rstrip(b"\0").decode().split(":")
        self.opened.append(int(r))
        return int(base)

    def ompimg_ipc_close(self, p):
        self.closed.append(p)
        return 0


class _FakePkg:
    class OmpimgError(RuntimeError):
        pass

    def __init__(self, rank):
        self._lib = _FakeLib(rank)

    def lib(self):
        return self._lib


def _cpu_worker(rank, world, port, out_dir):
    import torch.distributed as td
    sys.path.insert(0, str(ROOT))
    import bench
    bmp = importlib.import_module("openmp-parallel-computing_b200.bands_mp")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist = bench.Dist("gloo")
    try:
        H, pitch = 1001, 4096
        rows_of = bmp.band_rows(H, world)
        base = 0x10000000 * (rank + 1)
        pkg = _FakePkg(rank)
        nb = bmp.NeighbourBands(pkg, dist.gather, base, pitch, rows_of, rank, world)
        dist.host_barrier()
        np.save(os.path.join(out_dir, f"nb_{rank}.npy"),
                np.array([nb.row_above or 0, nb.row_below or 0] + sorted(pkg.lib().opened) + [-1] * (2 - len(pkg.lib().opened)), np.int64))
        nb.close()
        assert len(pkg.lib().closed) == len(pkg.lib().opened)
    finally:
        dist.close()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_band_protocol_over_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    bmp = importlib.import_module("openmp-parallel-computing_b200.bands_mp")
    mp.spawn(_cpu_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    H, pitch = 1001, 4096
    rows_of = bmp.band_rows(H, world)
    assert rows_of[0][0] == 0 and sum(n for _, n in rows_of) == H
    for r in range(world):
        above, below, o1, o2 = np.load(tmp_path / f"nb_{r}.npy").tolist()
        want_open = sorted(x for x in (r - 1, r + 1) if 0 <= x < world)
        assert [x for x in (o1, o2) if x >= 0] == want_open
        assert above == (0x10000000 * r + (rows_of[r - 1][1] - 1) * pitch if r > 0 else 0)        # last row of the band above
        assert below == (0x10000000 * (r + 2) if r + 1 < world else 0)                           # first row of the band below


def test_bench_seam_checker_catches_a_wrong_row(tmp_path, O):
    """bench.py's `seams` checker (the bands leg's byte-for-byte comparison with the oracle)"""
    sys.path.insert(0, str(ROOT))
    import bench
    w, h = 96, 40
    img = O.synth(w * h * 3, bench.BANDS_SEED).reshape(h, w, 3)
    e = O.rgb_to_edge(img)
    rows = [1, 9, 10, 19, 20, h - 2]
    npz = tmp_path / "s.npz"
    np.savez(npz, rows=np.array(rows), edge=e[rows], width=w, height=h)
    r = bench.cpu_job("seams", npz=str(npz))
    assert r == {"rows_checked": len(rows), "rows_differing": []}
    bad = e[rows].copy()
    bad[2, 17] ^= 1
    np.savez(npz, rows=np.array(rows), edge=bad, width=w, height=h)
    assert bench.cpu_job("seams", npz=str(npz))["rows_differing"] == [10]


def test_cfg5_golden_is_self_consistent():
    import json
    g = json.loads((ROOT / "tests" / "golden" / "bands_cfg5.json").read_text())
    assert g["width"] == g["height"] == 32768 and g["seed"] == 0xB2000005
    H = g["height"]
    for n in (2, 4, 8):
        for r in range(1, n):
            s = H * r // n
            for y in (s - 2, s - 1, s, s + 1):
                assert str(y) in g["row_sums"]
    assert 0 < g["interior_sum"] < 255 * 32766 * 32766


# ------------------------------------------------------------------------------------------------ GPU
def _gpu_worker(rank, world, port, out_dir, w, h, ndev):
    sys.path.insert(0, str(ROOT))
    import bench
    pkg = importlib.import_module("openmp-parallel-computing_b200")
    bmp = importlib.import_module("openmp-parallel-computing_b200.bands_mp")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist = bench.Dist("gloo")
    try:
        pkg.init(rank % ndev)
        L = pkg.lib()
        rows_of = bmp.band_rows(h, world)
        r0, nr = rows_of[rank]
        pitch, epitch = (w * 3 + 127) // 128 * 128, (w + 127) // 128 * 128
        band, edge = pkg.DeviceBuffer(nr * pitch), pkg.DeviceBuffer(nr * epitch)
        edge.fill(0x77)
        sys.path.insert(0, str(ROOT / "oracle"))
        import oracle as O
        img = O.synth(w * h * 3, bench.BANDS_SEED).reshape(h, w * 3)
        assert L.ompimg_copy2d_h2d(band.ptr, pitch, np.ascontiguousarray(img[r0:r0 + nr]).ctypes.data, w * 3, w * 3, nr, 0) == 0
        pkg.sync()
        nb = bmp.NeighbourBands(pkg, dist.gather, band.ptr, pitch, rows_of, rank, world)
        dist.host_barrier()                      # every band uploaded before a neighbour reads a halo row
        for _ in range(2):                       # twice: the result must not depend on what the first pass left behind
            rc = L.ompimg_rgb_to_edge_band_dev(band.ptr, pitch, 3, edge.ptr, epitch, w, nr, nb.row_above, nb.row_below, 0)
            assert rc == 0, L.ompimg_last_error()
        pkg.sync()
        out = np.empty((nr, w), np.uint8)
        assert L.ompimg_copy2d_d2h(out.ctypes.data, w, edge.ptr, epitch, w, nr, 0) == 0
        pkg.sync()
        np.save(os.path.join(out_dir, f"edge_{rank}.npy"), out)
        dist.host_barrier()                      # nobody unmaps or frees while a neighbour may still be reading
        nb.close()
        band.free(); edge.free()
    finally:
        dist.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,w,h", [(2, 1500, 257), (3, 4112, 100), (2, 8192, 520)])
def test_row_bands_across_processes_equal_oracle(world, w, h, tmp_path, O, pkg):
    """IPC-mapped neighbour bands (one process per band; distinct GPUs when the box has them)"""
    import torch.multiprocessing as mp
    sys.path.insert(0, str(ROOT))
    import bench
    bmp = importlib.import_module("openmp-parallel-computing_b200.bands_mp")
    ndev = pkg.lib().ompimg_device_count()
    assert ndev >= 1
    mp.spawn(_gpu_worker, args=(world, _free_port(), str(tmp_path), w, h, ndev), nprocs=world, join=True)
    want = O.rgb_to_edge(O.synth(w * h * 3, bench.BANDS_SEED).reshape(h, w, 3), np.full((h, w), 0x77, np.uint8))
    got = np.vstack([np.load(tmp_path / f"edge_{r}.npy") for r in range(world)])
    assert np.array_equal(got, want)
