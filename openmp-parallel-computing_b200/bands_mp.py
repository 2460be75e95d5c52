"""Row-band decomposition of one very large image with ONE PROCESS PER GPU (BASELINE config 5).

Each rank owns a band of rows.  The fused gray->Sobel kernel needs the source row just above and
just below the band, which live in the neighbouring ranks' allocations.  Two ways to get them, both
ending in the same C-ABI call (ompimg_rgb_to_edge_band_dev with row_above / row_below pointers), both
without a collective and bit-identical to the single-GPU run:

  * `NeighbourBands` (the product path): every rank exports the CUDA IPC handle of its band through
    the C ABI (ompimg_ipc_export), the 64-byte handles are all-gathered once over the launcher's side
    channel, and each rank maps its two neighbours' bands (ompimg_ipc_open).  The stencil kernel then
    fetches the halo rows itself, with bulk copies straight out of the peer GPU's memory over NVLink:
    no staging buffer, no exchange step, no torch tensor anywhere near the data.
  * `exchange_halo_rows` / `rgb_to_edge_band`: the rows travel by point-to-point send/recv over
    torch.distributed (NCCL on GPUs, gloo in the CPU tests) into two small halo buffers first.

(The single-process form, where one process drives all GPUs, is ompimg_bands_* in the C ABI / `Bands`.)
"""
from __future__ import annotations

import ctypes as C

IPC_HANDLE_BYTES = 64


def band_rows(height: int, world: int):
    """[(row0, nrows)] for every rank: contiguous, disjoint, covering [0, height)"""
    bounds = [height * r // world for r in range(world + 1)]
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def exchange_halo_rows(td, first_row, last_row, above_buf, below_buf, rank: int, world: int):
    """Send this band's first/last source row to the neighbours and receive theirs.

    first_row / last_row / above_buf / below_buf are 1-D uint8 tensors (views of the band and of the two
    halo buffers) on the device the process group communicates on.  Returns (has_above, has_below).
    """
    ops = []
    if rank > 0:
        ops.append(td.P2POp(td.isend, first_row, rank - 1))
        ops.append(td.P2POp(td.irecv, above_buf, rank - 1))
    if rank + 1 < world:
        ops.append(td.P2POp(td.isend, last_row, rank + 1))
        ops.append(td.P2POp(td.irecv, below_buf, rank + 1))
    if ops:
        for req in td.batch_isend_irecv(ops):
            req.wait()
    return rank > 0, rank + 1 < world


def rgb_to_edge_band(pkg, td, torch, rgb_band, edge_band, width: int, channels: int, rank: int, world: int, stream: int = 0):
    """rgb_band: (rows, pitch) uint8 CUDA tensor holding this rank's band; edge_band: (rows, epitch) uint8 CUDA tensor.
    Exchanges the halo rows, then runs the fused kernel on the band.  Returns after enqueueing the kernel."""
    rows, pitch = rgb_band.shape
    above = torch.empty(pitch, dtype=torch.uint8, device=rgb_band.device)
    below = torch.empty(pitch, dtype=torch.uint8, device=rgb_band.device)
    has_above, has_below = exchange_halo_rows(td, rgb_band[0], rgb_band[rows - 1], above, below, rank, world)
    torch.cuda.synchronize(rgb_band.device)          # the exchange ran on torch's stream, the kernel runs on the library's
    rc = pkg.lib().ompimg_rgb_to_edge_band_dev(rgb_band.data_ptr(), pitch, channels, edge_band.data_ptr(), edge_band.shape[1], width, rows,
                                              above.data_ptr() if has_above else None, below.data_ptr() if has_below else None, stream)
    if rc:
        raise pkg.OmpimgError(pkg.lib().ompimg_last_error().decode())
    return above, below                               # keep alive until the stream is synchronised


def gather_handles(gather, my_handle: bytes, world: int):
    """all ranks' IPC handles, indexed by rank.  `gather(obj) -> [obj of rank 0, ...]` is the launcher's side channel
    (bench.py: torch.distributed.all_gather_object; 64 bytes per rank, once -- the only traffic that is not the kernel's)"""
    assert len(my_handle) == IPC_HANDLE_BYTES
    out = gather(bytes(my_handle))
    assert len(out) == world and all(len(h) == IPC_HANDLE_BYTES for h in out)
    return out


def neighbour_ranks(rank: int, world: int):
    """(rank holding the band above or None, rank holding the band below or None)"""
    return (rank - 1 if rank > 0 else None), (rank + 1 if rank + 1 < world else None)


class NeighbourBands:
    """Maps the neighbouring ranks' bands into this process (CUDA IPC through the C ABI) and hands out the addresses of
    the two halo rows.  `band_ptr` is this rank's band from ompimg_dev_alloc, `pitch` its row pitch, `rows_of` the
    (row0, nrows) list of all ranks (band_rows)."""

    def __init__(self, pkg, gather, band_ptr: int, pitch: int, rows_of, rank: int, world: int):
        self.pkg, self.L = pkg, pkg.lib()
        h = (C.c_ubyte * IPC_HANDLE_BYTES)()
        rc = self.L.ompimg_ipc_export(band_ptr, h)
        if rc:
            raise pkg.OmpimgError("ompimg_ipc_export failed")
        handles = gather_handles(gather, bytes(h), world)
        up, down = neighbour_ranks(rank, world)
        self.mapped = []
        self.row_above = self.row_below = None

        def open_(r):
            buf = (C.c_ubyte * IPC_HANDLE_BYTES).from_buffer_copy(handles[r])
            p = self.L.ompimg_ipc_open(buf)
            if not p:
                raise pkg.OmpimgError(f"ompimg_ipc_open failed for the band of rank {r}")
            self.mapped.append(p)
            return p
        if up is not None:
            self.row_above = open_(up) + (rows_of[up][1] - 1) * pitch      # the last row of the band above
        if down is not None:
            self.row_below = open_(down)                                   # the first row of the band below

    def close(self):
        for p in self.mapped:
            self.L.ompimg_ipc_close(p)
        self.mapped = []
