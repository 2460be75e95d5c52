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
// Wire format (host byte order, one machine): request header, payload; reply header, payload.
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <signal.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include "../../include/ompimg_b200.h"
#include "worker.h"

namespace ompimg {

namespace {

constexpr uint32_t WK_MAGIC = 0x4F4D5042u;      // "OMPB"
enum : uint32_t { WK_GRAY = 1, WK_SOBEL = 2, WK_PING = 3 };

struct WkRequest {
    uint32_t magic, op;
    int32_t width, height, channels, luma_mode;
    uint64_t nbytes;                             // payload that follows
};
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
    return fd;
}

}  // namespace

// 1: served by the worker (result in `out`), 0: no worker configured or reachable, < 0: the worker answered with an error
int worker_call(uint32_t op, int w, int h, int ch, int luma_mode, const void *in, size_t in_bytes, void *out, size_t out_bytes,
                char *err, size_t err_cap) {
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
    WkRequest rq{WK_MAGIC, op, w, h, ch, luma_mode, (uint64_t)in_bytes};
    WkReply rp{};
    bool ok = write_full(fd, &rq, sizeof rq) && write_full(fd, in, in_bytes) && read_full(fd, &rp, sizeof rp);
    if (ok && rp.magic == WK_MAGIC && rp.status == OMPIMG_OK && rp.nbytes == out_bytes) ok = read_full(fd, out, out_bytes);
    else if (ok && rp.magic == WK_MAGIC && rp.status != OMPIMG_OK) {
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

void reply_error(int fd, int status, const char *msg) {
    WkReply rp{};
    rp.magic = WK_MAGIC; rp.status = status; rp.nbytes = 0;
    snprintf(rp.msg, sizeof rp.msg, "%s", msg);
    write_full(fd, &rp, sizeof rp);
}

void serve_connection(int fd) {
    PinBuf in = pool_take(), out = pool_take();   // pinned: recv() lands where the H2D copy starts
    for (;;) {
        WkRequest rq{};
        if (!read_full(fd, &rq, sizeof rq)) break;
        if (rq.magic != WK_MAGIC) break;
        const long long px = (long long)rq.width * rq.height;
        const uint64_t want = rq.op == WK_GRAY ? (uint64_t)px * (uint64_t)rq.channels : rq.op == WK_SOBEL ? (uint64_t)px : 0;
        if (rq.op == WK_PING) { WkReply rp{}; rp.magic = WK_MAGIC; if (!write_full(fd, &rp, sizeof rp)) break; continue; }
        if (rq.width <= 0 || rq.height <= 0 || rq.nbytes != want || (rq.op != WK_GRAY && rq.op != WK_SOBEL) || rq.nbytes > (1ull << 36)) {
            reply_error(fd, OMPIMG_ERR_ARG, "malformed request");
            break;
        }
        if (!in.reserve(rq.nbytes) || (rq.op == WK_SOBEL && !out.reserve(rq.nbytes))) {
            reply_error(fd, OMPIMG_ERR_NOMEM, "pinned allocation failed");
            break;
        }
        if (!read_full(fd, in.p, rq.nbytes)) break;
        int rc;
        const void *result = in.p;
        {
            std::lock_guard<std::mutex> lk(g_serve_mu);
            if (rq.op == WK_GRAY) {
                rc = ompimg_set_luma_mode(rq.luma_mode);
                if (rc == OMPIMG_OK) rc = ompimg_gray_inplace_host(static_cast<uint8_t *>(in.p), rq.width, rq.height, rq.channels);
            } else {
                memset(out.p, 0, rq.nbytes);
                rc = (rq.width < 3 || rq.height < 3) ? OMPIMG_OK
                                                     : ompimg_sobel_host(static_cast<const uint8_t *>(in.p), static_cast<uint8_t *>(out.p), rq.width, rq.height);
                result = out.p;
            }
            if (rc != OMPIMG_OK) { reply_error(fd, rc, ompimg_last_error()); continue; }
        }
        WkReply rp{};
        rp.magic = WK_MAGIC; rp.status = OMPIMG_OK; rp.nbytes = rq.nbytes;
        if (!write_full(fd, &rp, sizeof rp) || !write_full(fd, result, rq.nbytes)) break;
    }
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
    ::unlink(socket_path);
    const int ls = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (ls < 0) return OMPIMG_ERR_ARG;
    if (::bind(ls, reinterpret_cast<sockaddr *>(&a), sizeof a) != 0 || ::listen(ls, 64) != 0) { ::close(ls); return OMPIMG_ERR_ARG; }
    ::chmod(socket_path, 0666);
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
    fprintf(stderr, "ompimg_worker: serving %s on device %d\n", socket_path, ompimg_device());
    for (;;) {
        const int fd = ::accept4(ls, nullptr, nullptr, SOCK_CLOEXEC);
        if (fd < 0) { if (errno == EINTR) continue; break; }
        std::thread([fd, device] { ompimg_init(device); serve_connection(fd); }).detach();
    }
    ::close(ls);
    ::unlink(socket_path);
    return OMPIMG_OK;
}
