"""Row-band decomposition of one very large image with ONE PROCESS PER GPU (BASELINE config 5).

Each rank owns a band of rows.  The fused gray->Sobel kernel needs the source row just above and
just below the band, which live on the neighbouring ranks: they are exchanged with point-to-point
send/recv over torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) into two small
halo buffers, and the kernel is then launched through the C ABI with those buffers as its
row_above / row_below pointers (ompimg_rgb_to_edge_band_dev).  No collective is involved and the
result is bit-identical to the single-GPU run.

(The single-process form, where the kernel's producer warp loads the halo rows straight out of the
peer GPU's memory, is ompimg_bands_* in the C ABI / `Bands` in the package.)
"""
from __future__ import annotations


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
