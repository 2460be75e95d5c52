// worker.cu -- persistent worker mode.
//
// The reference's services exec the `grayscale` binary once per request
// (microservices/grayscale/app.py:24-34, event-driven/grayscale_service/app.py:54-65).  A fresh
// process pays ~3 s of CUDA driver + context initialisation before the first kernel, which is more
// than the whole CPU run it replaces.  With OMPIMG_WORKER=<unix socket path> in the environment the
// two frozen prototypes (convert_to_grayscale, sobel_edge) do NOT create a CUDA context: they send
// the image to a long-lived `ompimg_worker` process that holds the context (ompimg_worker_serve
// below) and read the result back.  The CLI contract above the boundary (argv, stdout, exit code)
// is untouched.  Still no CPU path: an unreachable worker means this process initialises CUDA itself.
//
// Trust model: the worker serves local processes of its own user (socket mode 0600, peer credentials checked with
// SO_PEERCRED; OMPIMG_WORKER_SOCKET_MODE / OMPIMG_WORKER_ALLOW_UIDS widen that deliberately).  A request can neither
// crash nor starve it: shared mappings must be size-sealed memfds (F_SEAL_SHRINK|F_SEAL_GROW, so no SIGBUS by
// truncation), mappings and payloads are capped (OMPIMG_WORKER_MAX_BYTES, default 1 GiB), nothing is pre-faulted, and
// a pooled result buffer never carries another client's bytes back (Sobel's untouched border is zeroed).
//
// Wire format (host byte order, one machine): request header, payload; reply header, payload.  The payload normally
// does not cross the socket: the client keeps a memfd mapping, passes the descriptor once (SCM_RIGHTS) and both sides
// work in that shared memory (socket copies cost ~8 ms per 6 MB image each way; a memcpy into the mapping 0.5 ms).
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include "../../include/ompimg_b200.h"
#include "worker.h"

namespace ompimg {

namespace {

constexpr uint32_t WK_MAGIC = 0x4F4D5042u;      // "OMPB"
enum : uint32_t { WK_GRAY = 1, WK_SOBEL = 2, WK_PING = 3, WK_JPEG_GRAY = 4 };   // 4: payload = a JPEG file, reply = decoded + grayscaled pixels; `reserved` = passes

enum : uint32_t { WK_SHM = 1, WK_NEWFD = 2 };     // flags: payload in the shared mapping / its descriptor rides on this message
struct WkRequest {
    uint32_t magic, op;
    int32_t width, height, channels, luma_mode;
    uint32_t flags, reserved;
    uint64_t nbytes;                             // payload (on the socket, or at offset 0 of the mapping)
};
inline uint64_t shm_out_offset(uint64_t nbytes) { return (nbytes + 4095u) & ~(uint64_t)4095u; }   // Sobel's result plane in the mapping
struct WkReply {
    uint32_t magic;
    int32_t status;                              // OMPIMG_OK or a negative OMPIMG_ERR_*
    uint64_t nbytes;                             // payload that follows (0 on error; then `msg` is set)
    char msg[232];
};

bool read_full(int fd, void *buf, size_t n) {
    uint8_t *p = static_cast<uint8_t *>(buf);
    while (n) {
        const ssize_t r = ::recv(fd, p, n, 0);
        if (r == 0) return false;
        if (r < 0) { if (errno == EINTR) continue; return false; }
        p += r; n -= (size_t)r;
    }
    return true;
}
bool write_full(int fd, const void *buf, size_t n) {
    const uint8_t *p = static_cast<const uint8_t *>(buf);
    while (n) {
        const ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        p += r; n -= (size_t)r;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ client
std::mutex g_client_mu;
int g_client_fd = -1;
bool g_client_warned = false;
int g_shm_fd = -1;                               // the client's shared mapping (grow-only) ...
uint8_t *g_shm = nullptr;
size_t g_shm_cap = 0;
bool g_shm_sent = false;                         // ... and whether the current connection has its descriptor
bool g_shm_off = false;

bool shm_reserve(size_t n) {
    if (g_shm_off) return false;
    if (g_shm_cap >= n) return true;
    if (g_shm) { munmap(g_shm, g_shm_cap); g_shm = nullptr; g_shm_cap = 0; }
    if (g_shm_fd >= 0) { ::close(g_shm_fd); g_shm_fd = -1; }
    const size_t want = (n + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
    // sealed against resizing: the worker refuses a mapping whose size could change under it
    const int fd = memfd_create("ompimg_b200", MFD_CLOEXEC | MFD_ALLOW_SEALING);
    if (fd < 0 || ftruncate(fd, (off_t)want) != 0 || fcntl(fd, F_ADD_SEALS, F_SEAL_SHRINK | F_SEAL_GROW) != 0) {
        if (fd >= 0) ::close(fd);
        g_shm_off = true;
        return false;
    }
    void *m = mmap(nullptr, want, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) { ::close(fd); g_shm_off = true; return false; }
    g_shm_fd = fd; g_shm = static_cast<uint8_t *>(m); g_shm_cap = want; g_shm_sent = false;
    return true;
}
bool send_with_fd(int sock, const void *buf, size_t n, int fd_to_pass) {
    msghdr mh{};
    iovec iov{const_cast<void *>(buf), n};
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))] = {};
    mh.msg_iov = &iov; mh.msg_iovlen = 1; mh.msg_control = ctl; mh.msg_controllen = sizeof ctl;
    cmsghdr *c = CMSG_FIRSTHDR(&mh);
    c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(c), &fd_to_pass, sizeof(int));
    for (;;) {
        const ssize_t r = ::sendmsg(sock, &mh, MSG_NOSIGNAL);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        return (size_t)r == n || write_full(sock, static_cast<const uint8_t *>(buf) + r, n - (size_t)r);
    }
}
// header of a request; a descriptor that rides on it comes back in *passed_fd (else -1)
bool recv_header(int sock, void *buf, size_t n, int *passed_fd) {
    *passed_fd = -1;
    msghdr mh{};
    iovec iov{buf, n};
    alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))] = {};
    mh.msg_iov = &iov; mh.msg_iovlen = 1; mh.msg_control = ctl; mh.msg_controllen = sizeof ctl;
    ssize_t r;
    do r = ::recvmsg(sock, &mh, MSG_CMSG_CLOEXEC); while (r < 0 && errno == EINTR);
    if (r <= 0) return false;
    for (cmsghdr *c = CMSG_FIRSTHDR(&mh); c; c = CMSG_NXTHDR(&mh, c))
        if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) memcpy(passed_fd, CMSG_DATA(c), sizeof(int));
    return (size_t)r == n || read_full(sock, static_cast<uint8_t *>(buf) + r, n - (size_t)r);
}

int client_connect(const char *path) {
    if (g_client_fd >= 0) return g_client_fd;
    sockaddr_un a{};
    a.sun_family = AF_UNIX;
    if (strlen(path) >= sizeof a.sun_path) return -1;
    strcpy(a.sun_path, path);
    const int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return -1;
    if (::connect(fd, reinterpret_cast<sockaddr *>(&a), sizeof a) != 0) { ::close(fd); return -1; }
    g_client_fd = fd;
    g_shm_sent = false;
    return fd;
}

}  // namespace

// 1: served by the worker (result in `out`), 0: no worker configured or reachable, < 0: the worker answered with an error
int worker_call(uint32_t op, int w, int h, int ch, int luma_mode, const void *in, size_t in_bytes, void *out, size_t out_bytes,
                char *err, size_t err_cap, uint32_t reserved = 0) {
    const char *path = getenv("OMPIMG_WORKER");
    if (!path || !*path) return 0;
    std::lock_guard<std::mutex> lk(g_client_mu);
    const int fd = client_connect(path);
    if (fd < 0) {
        if (!g_client_warned) {
            fprintf(stderr, "libompimg_b200: worker %s unreachable (%s); using this process's own CUDA context\n", path, strerror(errno));
            g_client_warned = true;
        }
        return 0;
    }
    // payload through the shared mapping when it can be had (input at 0; Sobel's result at shm_out_offset), else on the socket
    const bool in_place = in == out;
    const uint64_t out_off = in_place ? 0 : shm_out_offset(in_bytes);
    const bool shm = shm_reserve((size_t)out_off + out_bytes);
    WkRequest rq{WK_MAGIC, op, w, h, ch, luma_mode, 0u, reserved, (uint64_t)in_bytes};
    WkReply rp{};
    bool ok;
    if (shm) {
        memcpy(g_shm, in, in_bytes);
        rq.flags = WK_SHM | (g_shm_sent ? 0u : WK_NEWFD);
        ok = g_shm_sent ? write_full(fd, &rq, sizeof rq) : send_with_fd(fd, &rq, sizeof rq, g_shm_fd);
        g_shm_sent = ok;
        ok = ok && read_full(fd, &rp, sizeof rp);
    } else {
        ok = write_full(fd, &rq, sizeof rq) && write_full(fd, in, in_bytes) && read_full(fd, &rp, sizeof rp);
    }
    if (ok && rp.magic == WK_MAGIC && rp.status == OMPIMG_OK && rp.nbytes == out_bytes) {
        if (shm) memcpy(out, g_shm + out_off, out_bytes);
        else ok = read_full(fd, out, out_bytes);
    } else if (ok && rp.magic == WK_MAGIC && rp.status != OMPIMG_OK) {
        rp.msg[sizeof rp.msg - 1] = 0;
        snprintf(err, err_cap, "worker: %s", rp.msg);
        return rp.status;
    } else ok = false;
    if (!ok) {
        ::close(g_client_fd);
        g_client_fd = -1;
        snprintf(err, err_cap, "worker %s: connection lost or malformed reply", path);
        return OMPIMG_ERR_CUDA;
    }
    return 1;
}

int worker_call_gray(unsigned char *data, int w, int h, int ch, int luma_mode, char *err, size_t err_cap) {
    const size_t n = (size_t)w * h * ch;
    return worker_call(WK_GRAY, w, h, ch, luma_mode, data, n, data, n, err, err_cap);
}

// main.c:21-35 through the worker: the JPEG file goes over (a tenth of the pixels' size), the grayscaled pixels come back
int worker_call_jpeg_gray(const unsigned char *jpeg, size_t nbytes, int passes, int luma_mode, unsigned char *out, size_t out_bytes, int w, int h, int ch,
                          char *err, size_t err_cap) {
    return worker_call(WK_JPEG_GRAY, w, h, ch, luma_mode, jpeg, nbytes, out, out_bytes, err, err_cap, (uint32_t)(passes < 1 ? 1 : passes));
}

// the reference never writes dst's border (sobel.c:11-12): the plane comes back whole, only its interior is copied
int worker_call_sobel(const unsigned char *src, unsigned char *dst, int w, int h, char *err, size_t err_cap) {
    const size_t n = (size_t)w * h;
    unsigned char *tmp = static_cast<unsigned char *>(malloc(n));
    if (!tmp) { snprintf(err, err_cap, "out of host memory"); return OMPIMG_ERR_NOMEM; }
    const int rc = worker_call(WK_SOBEL, w, h, 1, 0, src, n, tmp, n, err, err_cap);
    if (rc == 1)
        for (int y = 1; y < h - 1; ++y) memcpy(dst + (size_t)y * w + 1, tmp + (size_t)y * w + 1, (size_t)w - 2);
    free(tmp);
    return rc;
}

// ------------------------------------------------------------------------------------------------ server
namespace {

std::mutex g_serve_mu;                           // one request on the GPU at a time (the host pipelines share staging buffers)

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n) {
        if (cap >= n) return true;
        if (p) ompimg_host_free(p);
        const size_t want = (n + (1 << 20) - 1) & ~(size_t)((1 << 20) - 1);
        p = ompimg_host_alloc(want);
        cap = p ? want : 0;
        return p != nullptr;
    }
};
// Pinned buffers outlive connections (page-locking 6 MB costs milliseconds, a request should not pay it):
// a connection borrows a pair from the pool and gives it back, grown to the largest image it saw.
std::mutex g_pool_mu;
std::vector<PinBuf> g_pool;
PinBuf pool_take() {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool.empty()) return PinBuf{};
    size_t best = 0;
    for (size_t i = 1; i < g_pool.size(); ++i) if (g_pool[i].cap > g_pool[best].cap) best = i;
    PinBuf b = g_pool[best];
    g_pool.erase(g_pool.begin() + (long)best);
    return b;
}
void pool_give(PinBuf b) {
    if (!b.p) return;
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.push_back(b);
}

uint64_t max_request_bytes() {
    static uint64_t v = 0;
    if (!v) {
        const char *e = getenv("OMPIMG_WORKER_MAX_BYTES");
        v = e ? strtoull(e, nullptr, 0) : 0;
        if (!v) v = 1ull << 30;                    // 1 GiB: a 16384 x 16384 RGBA image
    }
    return v;
}
// may this peer use the worker?  default: the worker's own user (and root); OMPIMG_WORKER_ALLOW_UIDS="1000,33" or "*" widens it
bool peer_allowed(int fd) {
    ucred cr{};
    socklen_t len = sizeof cr;
    if (getsockopt(fd, SOL_SOCKET, SO_PEERCRED, &cr, &len) != 0) return false;
    if (cr.uid == geteuid() || cr.uid == 0) return true;
    const char *e = getenv("OMPIMG_WORKER_ALLOW_UIDS");
    if (!e) return false;
    if (!strcmp(e, "*")) return true;
    for (const char *q = e; *q;) {                 // comma-separated decimal uids
        char *end;
        const unsigned long u = strtoul(q, &end, 10);
        if (end == q) break;
        if (u == cr.uid) return true;
        q = *end == ',' ? end + 1 : end;
    }
    return false;
}
// a descriptor a client passed: map it only if it is a memfd whose size is sealed and within the cap
bool map_client_memfd(int fd, uint8_t **map, size_t *len) {
    struct stat st{};
    const int seals = fcntl(fd, F_GET_SEALS);
    if (seals < 0 || (seals & (F_SEAL_SHRINK | F_SEAL_GROW)) != (F_SEAL_SHRINK | F_SEAL_GROW)) return false;
    if (fstat(fd, &st) != 0 || st.st_size <= 0 || (uint64_t)st.st_size > 2 * max_request_bytes() + 8192) return false;
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) return false;
    *map = static_cast<uint8_t *>(m);
    *len = (size_t)st.st_size;
    return true;
}

void reply_error(int fd, int status, const char *msg) {
    WkReply rp{};
    rp.magic = WK_MAGIC; rp.status = status; rp.nbytes = 0;
    snprintf(rp.msg, sizeof rp.msg, "%s", msg);
    write_full(fd, &rp, sizeof rp);
}

void serve_connection(int fd) {
    PinBuf in = pool_take(), out = pool_take();   // socket payloads: recv() lands in pinned memory, where the H2D copy starts
    uint8_t *map = nullptr;                       // the client's shared mapping, once its descriptor has arrived
    size_t map_len = 0;
    for (;;) {
        WkRequest rq{};
        int newfd = -1;
        if (!recv_header(fd, &rq, sizeof rq, &newfd)) break;
        if (newfd >= 0) {
            if (map) { munmap(map, map_len); map = nullptr; map_len = 0; }
            const bool mapped = map_client_memfd(newfd, &map, &map_len);
            ::close(newfd);
            if (!mapped) { reply_error(fd, OMPIMG_ERR_ARG, "shared mapping refused: it must be a memfd sealed with F_SEAL_SHRINK|F_SEAL_GROW, within the size cap"); break; }
        }
        if (rq.magic != WK_MAGIC) break;
        const long long px = (long long)rq.width * rq.height;
        const bool jpeg = rq.op == WK_JPEG_GRAY;
        const uint64_t want = rq.op == WK_GRAY ? (uint64_t)px * (uint64_t)rq.channels : rq.op == WK_SOBEL ? (uint64_t)px : jpeg ? rq.nbytes : 0;
        if (rq.op == WK_PING) { WkReply rp{}; rp.magic = WK_MAGIC; if (!write_full(fd, &rp, sizeof rp)) break; continue; }
        if (rq.width <= 0 || rq.height <= 0 || rq.nbytes != want || rq.nbytes == 0 || (rq.op != WK_GRAY && rq.op != WK_SOBEL && !jpeg) ||
            rq.nbytes > max_request_bytes() || (jpeg && ((rq.channels != 1 && rq.channels != 3) || (uint64_t)px * (uint64_t)rq.channels > max_request_bytes()))) {
            reply_error(fd, OMPIMG_ERR_ARG, "malformed request");
            break;
        }
        const uint64_t out_bytes = jpeg ? (uint64_t)px * (uint64_t)rq.channels : rq.nbytes;      // the client announces the decoded size it expects
        const bool separate_out = rq.op == WK_SOBEL || jpeg;
        const bool shm = (rq.flags & WK_SHM) != 0;
        uint8_t *src, *dst;
        if (shm) {
            const uint64_t out_off = separate_out ? shm_out_offset(rq.nbytes) : 0;
            if (!map || out_off + out_bytes > map_len) { reply_error(fd, OMPIMG_ERR_ARG, "shared mapping missing or too small"); break; }
            src = map; dst = map + out_off;
        } else {
            if (!in.reserve(rq.nbytes) || (separate_out && !out.reserve(out_bytes))) {
                reply_error(fd, OMPIMG_ERR_NOMEM, "pinned allocation failed");
                break;
            }
            if (!read_full(fd, in.p, rq.nbytes)) break;
            src = static_cast<uint8_t *>(in.p);
            dst = separate_out ? static_cast<uint8_t *>(out.p) : src;
            if (rq.op == WK_SOBEL) {
                // sobel_edge never writes the border and the pooled plane held an earlier client's result: zero the border
                const size_t w = (size_t)rq.width, h = (size_t)rq.height;
                memset(dst, 0, w);
                memset(dst + (h - 1) * w, 0, w);
                for (size_t y = 1; y + 1 < h; ++y) dst[y * w] = dst[y * w + w - 1] = 0;
            }
        }
        int rc;
        {
            std::lock_guard<std::mutex> lk(g_serve_mu);
            if (rq.op == WK_GRAY) {
                rc = ompimg_set_luma_mode(rq.luma_mode);
                if (rc == OMPIMG_OK) rc = ompimg_gray_inplace_host(src, rq.width, rq.height, rq.channels);
            } else if (jpeg) {
                int w = 0, h = 0, ch = 0;
                rc = ompimg_set_luma_mode(rq.luma_mode);
                if (rc == OMPIMG_OK) rc = ompimg_jpeg_gray_host(src, rq.nbytes, (int)rq.reserved, dst, out_bytes, &w, &h, &ch);
                if (rc == OMPIMG_OK && (w != rq.width || h != rq.height || ch != rq.channels)) { reply_error(fd, OMPIMG_ERR_ARG, "the JPEG is not the size the request announced"); continue; }
            } else {
                rc = (rq.width < 3 || rq.height < 3) ? OMPIMG_OK : ompimg_sobel_host(src, dst, rq.width, rq.height);
            }
            if (rc != OMPIMG_OK) { reply_error(fd, rc, ompimg_last_error()); continue; }
        }
        WkReply rp{};
        rp.magic = WK_MAGIC; rp.status = OMPIMG_OK; rp.nbytes = out_bytes;
        if (!write_full(fd, &rp, sizeof rp) || (!shm && !write_full(fd, dst, out_bytes))) break;
    }
    if (map) munmap(map, map_len);
    pool_give(in);
    pool_give(out);
    ::close(fd);
}

}  // namespace
}  // namespace ompimg

// Blocks: initialises the library on `device` (< 0: OMPIMG_DEVICE or 0), listens on `socket_path` and serves requests
// until the process is signalled.  Returns a negative OMPIMG_ERR_* if it cannot start.
extern "C" int ompimg_worker_serve(const char *socket_path, int device) {
    using namespace ompimg;
    if (!socket_path || !*socket_path) return OMPIMG_ERR_ARG;
    if (device < 0) { const char *d = getenv("OMPIMG_DEVICE"); device = d ? atoi(d) : 0; }
    int rc = ompimg_init(device);
    if (rc != OMPIMG_OK) return rc;
    sockaddr_un a{};
    a.sun_family = AF_UNIX;
    if (strlen(socket_path) >= sizeof a.sun_path) return OMPIMG_ERR_ARG;
    strcpy(a.sun_path, socket_path);
    {
        // a path left behind by a dead worker is removed; a path another worker still answers on is not stolen
        const int probe = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
        if (probe >= 0) {
            const bool live = ::connect(probe, reinterpret_cast<sockaddr *>(&a), sizeof a) == 0;
            ::close(probe);
            if (live) { fprintf(stderr, "ompimg_worker: %s is served by another worker\n", socket_path); return OMPIMG_ERR_ARG; }
        }
        struct stat st{};
        if (lstat(socket_path, &st) == 0 && S_ISSOCK(st.st_mode)) ::unlink(socket_path);
    }
    const int ls = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (ls < 0) return OMPIMG_ERR_ARG;
    const mode_t old_umask = umask(0177);          // the socket is born 0600: no window in which it is world-connectable
    const bool bound = ::bind(ls, reinterpret_cast<sockaddr *>(&a), sizeof a) == 0 && ::listen(ls, 64) == 0;
    umask(old_umask);
    if (!bound) { ::close(ls); return OMPIMG_ERR_ARG; }
    if (const char *m = getenv("OMPIMG_WORKER_SOCKET_MODE")) ::chmod(socket_path, (mode_t)(strtoul(m, nullptr, 8) & 0777));
    // warm the pipelines and the pinned pool once so that the first request pays neither module loading nor page-locking
    {
        PinBuf a, b;
        if (a.reserve(32u << 20) && b.reserve(32u << 20)) {
            memset(a.p, 0, 64 * 64 * 3);
            ompimg_gray_inplace_host(static_cast<uint8_t *>(a.p), 64, 64, 3);
            ompimg_sobel_host(static_cast<const uint8_t *>(a.p), static_cast<uint8_t *>(b.p), 64, 64);
        }
        pool_give(a); pool_give(b);
    }
    // requests normally arrive in pageable shared mappings: run that path once at 4K size, so that its pinned staging
    // buffers and helper threads exist before the first request (they cost tens of milliseconds to set up)
    {
        const int w = 3840, h = 2160;
        uint8_t *warm = static_cast<uint8_t *>(malloc((size_t)w * h * 4));
        if (warm) {
            memset(warm, 0, (size_t)w * h * 4);
            ompimg_gray_inplace_host(warm, w, h, 3);
            ompimg_sobel_host(warm, warm + (size_t)w * h * 3, w, h);
            free(warm);
        }
    }
    fprintf(stderr, "ompimg_worker: serving %s on device %d\n", socket_path, ompimg_device());
    for (;;) {
        const int fd = ::accept4(ls, nullptr, nullptr, SOCK_CLOEXEC);
        if (fd < 0) { if (errno == EINTR) continue; break; }
        if (!peer_allowed(fd)) { reply_error(fd, OMPIMG_ERR_ARG, "peer uid not allowed (OMPIMG_WORKER_ALLOW_UIDS)"); ::close(fd); continue; }
        std::thread([fd, device] { ompimg_init(device); serve_connection(fd); }).detach();
    }
    ::close(ls);
    ::unlink(socket_path);
    return OMPIMG_OK;
}
