"""Persistent worker mode (csrc/worker.cu).

CPU tests: the CLIENT side of the frozen prototypes -- with OMPIMG_WORKER set they must not touch CUDA at all, so
it can be exercised here without a GPU against a fake worker that speaks the wire format and answers with the
oracle's results (the oracle is the checker, never the product).  GPU tests: the real `ompimg_worker` process
serving the reference's unmodified CLI drivers.
"""
import hashlib
import mmap
import os
import socket
import struct
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import pytest

from common import ROOT, sha

REQ = struct.Struct("<IIiiiiIIQ")        # magic, op, width, height, channels, luma_mode, flags, reserved, nbytes
WK_SHM, WK_NEWFD = 1, 2
REP = struct.Struct("<IiQ232s")          # magic, status, nbytes, msg
MAGIC = 0x4F4D5042
LIB = ROOT / "openmp-parallel-computing_b200" / "lib" / "libompimg_b200.so"
WORKER = ROOT / "openmp-parallel-computing_b200" / "lib" / "ompimg_worker"

CLIENT = r"""
import ctypes as C, hashlib, sys, numpy as np
L = C.CDLL(sys.argv[1])
h, w, ch = 37, 53, int(sys.argv[2])
rng = np.random.default_rng(5)
img = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
for _ in range(int(sys.argv[3])):
    L.convert_to_grayscale(C.c_void_p(img.ctypes.data), w, h, ch)
plane = np.ascontiguousarray(img[:, :, 0])
dst = np.full((h, w), 0xAB, np.uint8)
L.sobel_edge(C.c_void_p(plane.ctypes.data), C.c_void_p(dst.ctypes.data), w, h)
print(hashlib.sha256(img.tobytes()).hexdigest(), hashlib.sha256(dst.tobytes()).hexdigest(), L.ompimg_launch_count())
"""


def recv_full(c, n):
    buf = bytearray()
    while len(buf) < n:
        part = c.recv(n - len(buf))
        if not part:
            return None
        buf += part
    return bytes(buf)


class FakeWorker:
    """Speaks worker.cu's wire format; computes with the oracle.  `fail_op` makes that op answer with an error."""

    def __init__(self, O, path, fail_op=None):
        self.O, self.path, self.fail_op = O, str(path), fail_op
        self.requests = []
        self.srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        self.srv.bind(self.path)
        self.srv.listen(8)
        self.thread = threading.Thread(target=self.run, daemon=True)
        self.thread.start()

    def run(self):
        while True:
            try:
                c, _ = self.srv.accept()
            except OSError:
                return
            threading.Thread(target=self.serve, args=(c,), daemon=True).start()

    def serve(self, c):
        shm = None
        with c:
            while True:
                hdr, fds, _, _ = socket.recv_fds(c, REQ.size, 4)           # the mapping's descriptor rides on a header (SCM_RIGHTS)
                if not hdr:
                    return
                if len(hdr) < REQ.size:
                    hdr += recv_full(c, REQ.size - len(hdr))
                if fds:
                    shm = mmap.mmap(fds[0], 0)
                    for f in fds:
                        os.close(f)
                magic, op, w, h, ch, mode, flags, reserved, n = REQ.unpack(hdr)   # reserved = passes for op 4
                assert magic == MAGIC
                assert bool(flags & WK_NEWFD) == bool(fds)
                if flags & WK_SHM:
                    data = bytes(shm[:n])
                else:
                    data = recv_full(c, n)
                self.requests.append((op, w, h, ch, mode, n, flags))
                if op == self.fail_op:
                    c.sendall(REP.pack(MAGIC, -3, 0, b"injected failure"))
                    continue
                if op == 1:
                    img = np.frombuffer(data, np.uint8).reshape(h, w, ch).copy()
                    out, off = self.O.gray_inplace(img, ch, fma=bool(mode)), 0
                elif op == 4:                                       # a JPEG file in, decoded + grayscaled pixels out (`reserved` = passes)
                    import tempfile
                    with tempfile.NamedTemporaryFile(suffix=".jpg") as t:
                        t.write(data); t.flush()
                        img = np.ascontiguousarray(self.O.stb_decode(t.name))      # the reference's own decoder
                    assert img.shape[:2] == (h, w)
                    for _p in range(reserved):
                        img = self.O.gray_inplace(img, ch, fma=bool(mode)) if ch == 3 else img
                    out, off = img, (n + 4095) & ~4095
                    n_out = out.nbytes
                    if flags & WK_SHM:
                        shm[off:off + n_out] = out.tobytes()
                        c.sendall(REP.pack(MAGIC, 0, n_out, b""))
                    else:
                        c.sendall(REP.pack(MAGIC, 0, n_out, b"") + out.tobytes())
                    continue
                else:
                    out, off = self.O.sobel(np.frombuffer(data, np.uint8).reshape(h, w).copy(), np.zeros((h, w), np.uint8)), (n + 4095) & ~4095
                if flags & WK_SHM:
                    shm[off:off + n] = out.tobytes()
                    c.sendall(REP.pack(MAGIC, 0, n, b""))
                else:
                    c.sendall(REP.pack(MAGIC, 0, n, b"") + out.tobytes())

    def close(self):
        self.srv.close()


def run_client(sock, ch=3, passes=1, extra_env=None):
    env = dict(os.environ, OMPIMG_WORKER=str(sock), CUDA_VISIBLE_DEVICES="", **(extra_env or {}))
    return subprocess.run([sys.executable, "-c", CLIENT, str(LIB), str(ch), str(passes)], capture_output=True, text=True, env=env)


@pytest.mark.parametrize("ch,passes,mode", [(3, 1, "strict"), (4, 2, "strict"), (3, 1, "fma")])
def test_frozen_prototypes_go_through_the_worker_without_cuda(O, tmp_path, ch, passes, mode):
    """CUDA_VISIBLE_DEVICES is empty in the client: any attempt to initialise CUDA there would abort it."""
    fw = FakeWorker(O, tmp_path / "w.sock")
    try:
        r = run_client(tmp_path / "w.sock", ch, passes, {"OMPIMG_LUMA": mode})
        assert r.returncode == 0, r.stderr
        got_img, got_edge, launches = r.stdout.split()
        rng = np.random.default_rng(5)
        img = rng.integers(0, 256, (37, 53, ch), dtype=np.uint8)
        for _ in range(passes):
            img = O.gray_inplace(img, ch, fma=(mode == "fma"))
        want_edge = O.sobel(np.ascontiguousarray(img[:, :, 0]), np.full((37, 53), 0xAB, np.uint8))    # border bytes stay 0xAB (sobel.c:11-12)
        assert got_img == sha(img) and got_edge == sha(want_edge)
        assert launches == "0"                                     # nothing ran in the client process
        assert [q[0] for q in fw.requests] == [1] * passes + [2]
        # the payload travels through the shared mapping; its descriptor is passed when the mapping is (re)created
        assert all(q[6] & WK_SHM for q in fw.requests) and fw.requests[0][6] & WK_NEWFD
        assert all(q[4] == (1 if mode == "fma" else 0) for q in fw.requests if q[0] == 1)
    finally:
        fw.close()


def test_worker_error_aborts_like_any_other_failure(O, tmp_path):
    """The frozen functions return void: a worker-side failure ends the process with a message (HTTP 500 upstream)."""
    fw = FakeWorker(O, tmp_path / "w.sock", fail_op=1)
    try:
        r = run_client(tmp_path / "w.sock")
        assert r.returncode != 0
        assert "convert_to_grayscale failed" in r.stderr and "injected failure" in r.stderr
    finally:
        fw.close()


def test_unreachable_worker_means_in_process_cuda_not_a_cpu_path(tmp_path):
    r = run_client(tmp_path / "nobody.sock")
    assert "unreachable" in r.stderr
    if r.returncode != 0:                                          # no GPU visible to the client: it must fail loudly
        assert "no usable CUDA device" in r.stderr and "no CPU fallback" in r.stderr


def test_reference_cli_through_a_worker_matches_the_golden_png(O, golden, tmp_path):
    """The reference's unmodified main.c linked against the library, served by the (fake) worker: test_gray.png."""
    exe = ROOT / "oracle" / "_ref" / "bin" / "grayscale_b200"
    img = ROOT / "oracle" / "_ref" / "fixtures" / "test.jpg"
    if not exe.exists() or not img.exists():
        pytest.skip("drop-in binaries not built (oracle/Makefile `dropin`, needs /root/reference at build time)")
    fw = FakeWorker(O, tmp_path / "w.sock")
    try:
        out = tmp_path / "o.png"
        r = subprocess.run([str(exe), str(img), str(out), "1"], capture_output=True, text=True,
                           env=dict(os.environ, OMPIMG_WORKER=str(tmp_path / "w.sock"), CUDA_VISIBLE_DEVICES=""))
        assert r.returncode == 0, r.stderr
        assert "Compute kernel" in r.stdout
        assert hashlib.md5(out.read_bytes()).hexdigest() == golden["test_jpg"]["golden_png_md5"]
    finally:
        fw.close()


# ------------------------------------------------------------------------------------------------ real worker
@pytest.fixture()
def real_worker(tmp_path):
    sock = tmp_path / "gpu.sock"
    p = subprocess.Popen([str(WORKER), "--socket", str(sock)], stderr=subprocess.PIPE, text=True)
    t0 = time.time()
    while not sock.exists():
        assert p.poll() is None, p.stderr.read()
        assert time.time() - t0 < 120, "worker did not come up"
        time.sleep(0.05)
    time.sleep(0.2)
    yield sock
    p.terminate()
    p.wait(timeout=10)


@pytest.mark.gpu
def test_real_worker_serves_the_frozen_prototypes(O, real_worker):
    r = run_client(real_worker, 3, 2)
    assert r.returncode == 0, r.stderr
    got_img, got_edge, launches = r.stdout.split()
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    for _ in range(2):
        img = O.gray_inplace(img, 3)
    want_edge = O.sobel(np.ascontiguousarray(img[:, :, 0]), np.full((37, 53), 0xAB, np.uint8))
    assert got_img == sha(img) and got_edge == sha(want_edge) and launches == "0"


@pytest.mark.gpu
def test_real_worker_behind_the_reference_cli(golden, real_worker, tmp_path):
    """BASELINE config 1 and 2 through the worker: byte-identical PNGs, and the per-request cost without the ~3 s of
    CUDA initialisation a fresh process pays (printed)."""
    fix = ROOT / "oracle" / "_ref" / "fixtures"
    exe = ROOT / "oracle" / "_ref" / "bin"
    if not (exe / "grayscale_b200").exists():
        pytest.skip("drop-in binaries not built")
    env = dict(os.environ, OMPIMG_WORKER=str(real_worker), CUDA_VISIBLE_DEVICES="")

    def cli(binary, image, passes, e):
        out = tmp_path / f"{binary}.png"
        t0 = time.time()
        r = subprocess.run([str(exe / binary), str(fix / image), str(out), str(passes)], capture_output=True, text=True, env=e)
        assert r.returncode == 0, r.stderr
        return hashlib.md5(out.read_bytes()).hexdigest(), time.time() - t0, r.stdout.strip()

    md5, t_w, line = cli("grayscale_b200", "test.jpg", 1, env)
    assert md5 == golden["test_jpg"]["golden_png_md5"]
    md5, t_w3, line3 = cli("grayscale_b200", "test.jpg", 3, env)
    assert md5 == golden["test_jpg"]["strict_p3_png_md5"]
    md5, t_s, _ = cli("sobel_b200", "more_than_one_mega_photo.jpg", 1, env)
    assert md5 == golden["photo_jpg"]["strict_p1_png_md5"]
    md5, t_own, line_own = cli("grayscale_b200", "test.jpg", 1, dict(os.environ))      # own CUDA context
    assert md5 == golden["test_jpg"]["golden_png_md5"]
    print(f"\ngrayscale test.jpg x1: worker {t_w:.3f} s [{line}], x3 [{line3}] | own context {t_own:.3f} s [{line_own}] | sobel photo via worker {t_s:.3f} s")


@pytest.mark.gpu
def test_real_worker_serves_concurrent_clients(O, real_worker):
    """Several CLI-like clients at once (the services run one process per request): every one gets its own answer."""
    procs = [subprocess.Popen([sys.executable, "-c", CLIENT, str(LIB), str(3 + (i & 1)), str(1 + i % 3)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True, env=dict(os.environ, OMPIMG_WORKER=str(real_worker), CUDA_VISIBLE_DEVICES="")) for i in range(8)]
    for i, p in enumerate(procs):
        out, err = p.communicate(timeout=120)
        assert p.returncode == 0, err
        ch, passes = 3 + (i & 1), 1 + i % 3
        rng = np.random.default_rng(5)
        img = rng.integers(0, 256, (37, 53, ch), dtype=np.uint8)
        for _ in range(passes):
            img = O.gray_inplace(img, ch)
        want_edge = O.sobel(np.ascontiguousarray(img[:, :, 0]), np.full((37, 53), 0xAB, np.uint8))
        assert out.split()[:2] == [sha(img), sha(want_edge)], i


# ------------------------------------------------------------------------------------------------ hardening
def test_client_mapping_is_size_sealed(O, tmp_path):
    """The client's memfd must be sealed against resizing (the worker refuses anything else): a truncation after the
    worker's bounds check would otherwise SIGBUS the shared worker."""
    import fcntl
    seen = []

    class SealWorker(FakeWorker):
        def serve(self, c):
            with c:
                hdr, fds, _, _ = socket.recv_fds(c, REQ.size, 4)
                if fds:
                    seen.append(fcntl.fcntl(fds[0], fcntl.F_GET_SEALS))
                    try:
                        os.ftruncate(fds[0], 4096)                  # a sealed memfd cannot be resized by anyone
                        seen.append("truncated")
                    except OSError:
                        seen.append("sealed")
                c.sendall(REP.pack(MAGIC, -3, 0, b"stop"))

    fw = SealWorker(O, tmp_path / "w.sock")
    try:
        run_client(tmp_path / "w.sock", 3, 1)
    finally:
        fw.close()
    assert seen and isinstance(seen[0], int)
    want = fcntl.F_SEAL_SHRINK | fcntl.F_SEAL_GROW
    assert seen[0] & want == want and seen[1] == "sealed"


def _raw_request(sock_path, header, payload=b"", fd=None):
    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    c.settimeout(30)
    c.connect(str(sock_path))
    with c:
        if fd is not None:
            socket.send_fds(c, [header], [fd])
        else:
            c.sendall(header)
        if payload:
            c.sendall(payload)
        rep = recv_full(c, REP.size)
        if rep is None:
            return None, None
        magic, status, n, msg = REP.unpack(rep)
        body = recv_full(c, n) if n and status == 0 and fd is None else b""
        return status, (msg.rstrip(b"\0").decode(), body)


@pytest.mark.gpu
def test_real_worker_refuses_what_could_hurt_it(O, real_worker):
    assert oct(os.stat(real_worker).st_mode & 0o777) == "0o600"                 # born private
    # (1) an unsealed memfd: refused before it is mapped
    fd = os.memfd_create("evil", os.MFD_CLOEXEC)
    os.ftruncate(fd, 1 << 20)
    st, (msg, _) = _raw_request(real_worker, REQ.pack(MAGIC, 1, 16, 16, 3, 0, WK_SHM | WK_NEWFD, 0, 16 * 16 * 3), fd=fd)
    os.close(fd)
    assert st == -2 and "sealed" in msg
    # (2) a payload beyond the cap: refused before anything is allocated
    st, (msg, _) = _raw_request(real_worker, REQ.pack(MAGIC, 1, 65536, 65536, 4, 0, 0, 0, 65536 * 65536 * 4))
    assert st == -2
    # (3) the worker is still alive and a pooled Sobel plane comes back with a ZERO border, not an earlier client's bytes
    h, w = 40, 64
    noise = np.random.default_rng(1).integers(1, 256, (h, w), dtype=np.uint8)
    for plane in (noise, np.zeros((h, w), np.uint8)):
        st, (msg, body) = _raw_request(real_worker, REQ.pack(MAGIC, 2, w, h, 1, 0, 0, 0, w * h), plane.tobytes())
        assert st == 0, msg
        got = np.frombuffer(body, np.uint8).reshape(h, w)
        assert np.array_equal(got, O.sobel(plane, np.zeros((h, w), np.uint8)))


@pytest.mark.gpu
def test_second_worker_does_not_steal_a_live_socket(real_worker):
    r = subprocess.run([str(WORKER), "--socket", str(real_worker)], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "another worker" in r.stderr
    assert run_client(real_worker, 3, 1).returncode == 0                       # the first worker still serves


# ------------------------------------------------------------------------------------------------ the CLI with the codec steps replaced
CLI = ROOT / "openmp-parallel-computing_b200" / "lib" / "ompimg_grayscale"


def _png_pixels(path):
    from PIL import Image
    return np.ascontiguousarray(np.asarray(Image.open(path)))


def test_cli_client_side_and_png_writer_against_a_fake_worker(O, golden, tmp_path):
    """ompimg_grayscale (main.c's contract with decode and encode replaced) without a GPU: the JPEG goes to a fake worker that
    answers with the reference's decoder + the oracle; the PNG this process writes must decode to the golden pixels."""
    img = ROOT / "oracle" / "_ref" / "fixtures" / "test.jpg"
    if not img.exists() or O.stb_decode(img) is None:
        pytest.skip("reference image / stb_decode not on this box")
    fw = FakeWorker(O, tmp_path / "w.sock")
    try:
        for passes, key in ((1, "golden_gray_sha256"), (3, "strict_p2_sha256")):
            out = tmp_path / f"o{passes}.png"
            r = subprocess.run([str(CLI), str(img), str(out), str(passes)], capture_output=True, text=True,
                               env=dict(os.environ, OMPIMG_WORKER=str(tmp_path / "w.sock"), CUDA_VISIBLE_DEVICES="", OMPIMG_PNG_THREADS=str(passes)))
            assert r.returncode == 0, r.stderr
            assert r.stdout.startswith("Compute kernel \u00d7%d: " % passes) and r.stdout.rstrip().endswith(" s")
            assert sha(_png_pixels(out)) == golden["test_jpg"][key]
        assert [q[0] for q in fw.requests] == [4, 4]
        assert fw.requests[0][5] == img.stat().st_size                  # the FILE travelled, not the pixels
    finally:
        fw.close()
    # argv and error contract of main.c:15-25
    r = subprocess.run([str(CLI)], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Uso: ")
    r = subprocess.run([str(CLI), str(tmp_path / "missing.jpg"), str(tmp_path / "x.png")], capture_output=True, text=True)
    assert r.returncode == 1 and "Errore caricando immagine" in r.stderr


@pytest.mark.gpu
def test_cli_with_codec_steps_behind_the_real_worker(O, golden, real_worker, tmp_path):
    """BASELINE config 1 end to end, the reference's real workload: `grayscale test.jpg out.png 1` as a whole process."""
    fix = ROOT / "oracle" / "_ref" / "fixtures"
    if not (fix / "test.jpg").exists():
        pytest.skip("reference images did not travel")
    env = dict(os.environ, OMPIMG_WORKER=str(real_worker), CUDA_VISIBLE_DEVICES="")

    def run(exe, image, passes, e):
        out = tmp_path / "o.png"
        t0 = time.time()
        r = subprocess.run([str(exe), str(fix / image), str(out), str(passes)], capture_output=True, text=True, env=e)
        dt = time.time() - t0
        assert r.returncode == 0, r.stderr
        return _png_pixels(out), dt, r.stdout.strip()

    px, t_first, _ = run(CLI, "test.jpg", 1, env)
    assert sha(px) == golden["test_jpg"]["golden_gray_sha256"]
    times = [run(CLI, "test.jpg", 1, env)[1] for _ in range(5)]
    px3, _, line3 = run(CLI, "test.jpg", 3, env)
    assert sha(px3) == golden["test_jpg"]["strict_p2_sha256"]
    for name, key in (("more_than_one_mega_photo.jpg", "photo_jpg"), ("half_of_a_mega_photo.jpg", "half_jpg")):
        if (fix / name).exists():
            px, _, _ = run(CLI, name, 1, env)
            rgb = O.stb_decode(fix / name)
            assert np.array_equal(px, O.gray_inplace(np.ascontiguousarray(rgb), 3))
    ref = ROOT / "oracle" / "_ref" / "bin" / "grayscale_strict"
    msg = f"\nompimg_grayscale test.jpg x1 behind the worker: first {t_first:.3f} s, then {min(times):.3f}..{max(times):.3f} s (median {sorted(times)[2]:.3f}) [{line3}]"
    if ref.exists():
        ref_times = []
        for _ in range(5):
            out = tmp_path / "r.png"
            t0 = time.time()
            assert subprocess.run([str(ref), str(fix / "test.jpg"), str(out), "1"], capture_output=True, env=dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count()))).returncode == 0
            ref_times.append(time.time() - t0)
        msg += f" | reference CLI (all host threads): {min(ref_times):.3f}..{max(ref_times):.3f} s (median {sorted(ref_times)[2]:.3f})"
        assert sha(_png_pixels(tmp_path / "r.png")) == golden["test_jpg"]["golden_gray_sha256"]
    print(msg)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "cli_codec_times.txt").write_text(msg.strip() + "\n")
