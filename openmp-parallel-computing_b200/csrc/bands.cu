// bands.cu -- one very large image split into row bands over several GPUs of one process
// (BASELINE config 5: 32768 x 32768 RGB Sobel over 8 B200).
//
// Every device holds only its own band (RGB in, edge plane out).  The fused gray->Sobel kernel
// needs one luma row above and below the band; the producer warp of the stencil kernel reads those
// two RGB rows straight from the neighbouring device's band through the NVLink peer mapping
// (cudaDeviceEnablePeerAccess), so there is no staging copy and no collective: the halo exchange
// is fused into the kernel.  Where peer access is unavailable the two rows are copied with
// cudaMemcpyPeerAsync into a local halo buffer first (same kernel, local pointers).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/ompimg_b200.h"
#include "kernels.h"

using namespace ompimg;

namespace ompimg {

__global__ void __launch_bounds__(256) synth2d_kernel(uint8_t *dst, size_t pitch, unsigned long long row_bytes, int rows, uint32_t seed,
                                                      unsigned long long off) {
    const unsigned long long total = row_bytes * rows;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long y = i / row_bytes, x = i - y * row_bytes;
        const unsigned long long gi = off + i;
        uint32_t v = seed + (uint32_t)(gi >> 2) * 0x9E3779B9u;
        v ^= v >> 16; v *= 0x7feb352du; v ^= v >> 15; v *= 0x846ca68bu; v ^= v >> 16;
        dst[y * pitch + x] = (uint8_t)(v >> (8 * (gi & 3)));
    }
}

// sum of the bytes of a 2-D region (checksum of an edge plane's interior)
__global__ void __launch_bounds__(256) byte_sum_kernel(const uint8_t *src, size_t pitch, int x0, int x1, int y0, int y1, unsigned long long *out) {
    unsigned long long s = 0;
    const long long wr = x1 - x0, total = wr * (long long)(y1 - y0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long y = i / wr, x = i - y * wr;
        s += src[(size_t)(y0 + y) * pitch + x0 + x];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

cudaError_t launch_byte_sum(const uint8_t *src, size_t pitch, int width, int rows, unsigned long long *sum_dev, int sm_count, cudaStream_t st) {
    count_launch();
    byte_sum_kernel<<<sm_count * 8, 256, 0, st>>>(src, pitch, 0, width, 0, rows, sum_dev);
    return cudaGetLastError();
}

}  // namespace ompimg

constexpr int BANDS_TAIL_ROWS = 14;   // K <= 15

struct ompimg_bands {
    int w, h, ch, ndev;
    int dev[16];
    int row0[17];
    int sm_count[16];
    size_t rpitch, epitch;
    uint8_t *rgb[16];
    uint8_t *edge[16];
    uint8_t *halo[16];          // 2 rows (above, below) when peer access is unavailable
    uint8_t *conv[16];          // convolution output of the band (allocated by the first ompimg_bands_conv_u8)
    size_t cpitch; int conv_k;  // its pitch and the K it was sized for
    unsigned long long *sum[16];
    bool direct[16][2];         // can device i read its (above, below) neighbour directly?
    cudaStream_t st[16];
    cudaEvent_t e0[16], e1[16];
    int home;                   // device current before create()
    char err[256];
};

#define BCK(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            fprintf(stderr, "libompimg_b200: bands: %s failed: %s (%s:%d)\n", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return OMPIMG_ERR_CUDA;                                                                     \
        }                                                                                               \
    } while (0)

extern "C" ompimg_bands *ompimg_bands_create(int w, int h, int ch, int ndev, const int *devices) {
    if (w < 3 || h < 3 || !(ch == 1 || ch == 3 || ch == 4) || ndev < 1 || ndev > 16 || h < 2 * ndev) return nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return nullptr;
    ompimg_bands *b = (ompimg_bands *)calloc(1, sizeof(ompimg_bands));
    if (!b) return nullptr;
    b->w = w; b->h = h; b->ch = ch; b->ndev = ndev;
    cudaGetDevice(&b->home);
    b->rpitch = ((size_t)w * ch + 127) / 128 * 128;
    b->epitch = ((size_t)w + 127) / 128 * 128;
    for (int i = 0; i <= ndev; ++i) b->row0[i] = (int)((long long)h * i / ndev);
    bool ok = true;
    for (int i = 0; i < ndev && ok; ++i) {
        b->dev[i] = devices ? devices[i] : i % count;
        if (b->dev[i] < 0 || b->dev[i] >= count) { ok = false; break; }
        const int rows = b->row0[i + 1] - b->row0[i];
        ok = cudaSetDevice(b->dev[i]) == cudaSuccess &&
             cudaMalloc(&b->rgb[i], b->rpitch * (rows + BANDS_TAIL_ROWS)) == cudaSuccess &&        // + room for the K-1 rows below (convolution)
             cudaMalloc(&b->edge[i], b->epitch * rows) == cudaSuccess &&
             cudaMalloc(&b->halo[i], 2 * b->rpitch) == cudaSuccess &&
             cudaMalloc(&b->sum[i], 8) == cudaSuccess &&
             cudaMemset(b->edge[i], 0, b->epitch * rows) == cudaSuccess &&
             cudaStreamCreateWithFlags(&b->st[i], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreate(&b->e0[i]) == cudaSuccess && cudaEventCreate(&b->e1[i]) == cudaSuccess &&
             cudaDeviceGetAttribute(&b->sm_count[i], cudaDevAttrMultiProcessorCount, b->dev[i]) == cudaSuccess;
    }
    for (int i = 0; i < ndev && ok; ++i) {
        cudaSetDevice(b->dev[i]);
        for (int side = 0; side < 2; ++side) {
            const int j = side == 0 ? i - 1 : i + 1;
            b->direct[i][side] = false;
            if (j < 0 || j >= ndev) continue;
            if (b->dev[j] == b->dev[i]) { b->direct[i][side] = true; continue; }
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, b->dev[i], b->dev[j]) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b->dev[j], 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) b->direct[i][side] = true;
                cudaGetLastError();
            }
        }
    }
    cudaSetDevice(b->home);
    if (!ok) { ompimg_bands_destroy(b); return nullptr; }
    return b;
}

extern "C" int ompimg_bands_destroy(ompimg_bands *b) {
    if (!b) return OMPIMG_OK;
    for (int i = 0; i < b->ndev; ++i) {
        cudaSetDevice(b->dev[i]);
        if (b->st[i]) { cudaStreamSynchronize(b->st[i]); cudaStreamDestroy(b->st[i]); }
        if (b->e0[i]) cudaEventDestroy(b->e0[i]);
        if (b->e1[i]) cudaEventDestroy(b->e1[i]);
        cudaFree(b->rgb[i]); cudaFree(b->edge[i]); cudaFree(b->halo[i]); cudaFree(b->sum[i]); cudaFree(b->conv[i]);
    }
    cudaGetLastError();
    cudaSetDevice(b->home);
    free(b);
    return OMPIMG_OK;
}

extern "C" int ompimg_bands_rows(const ompimg_bands *b, int i, int *row0, int *nrows) {
    if (!b || i < 0 || i >= b->ndev) return OMPIMG_ERR_ARG;
    if (row0) *row0 = b->row0[i];
    if (nrows) *nrows = b->row0[i + 1] - b->row0[i];
    return OMPIMG_OK;
}

static int bands_sync(ompimg_bands *b) {
    for (int i = 0; i < b->ndev; ++i) { BCK(cudaSetDevice(b->dev[i])); BCK(cudaStreamSynchronize(b->st[i])); }
    BCK(cudaSetDevice(b->home));
    return OMPIMG_OK;
}

extern "C" int ompimg_bands_synth(ompimg_bands *b, unsigned int seed) {
    if (!b) return OMPIMG_ERR_ARG;
    const unsigned long long row = (unsigned long long)b->w * b->ch;
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        const int rows = b->row0[i + 1] - b->row0[i];
        const unsigned long long off = row * b->row0[i];
        if (b->rpitch == row) BCK(launch_synth(b->rgb[i], row * rows, seed, off, b->sm_count[i], b->st[i]));
        else {
            count_launch();
            synth2d_kernel<<<b->sm_count[i] * 8, 256, 0, b->st[i]>>>(b->rgb[i], b->rpitch, row, rows, seed, off);
            BCK(cudaGetLastError());
        }
    }
    return bands_sync(b);
}

extern "C" int ompimg_bands_upload(ompimg_bands *b, const uint8_t *rgb_host) {
    if (!b || !rgb_host) return OMPIMG_ERR_ARG;
    const size_t row = (size_t)b->w * b->ch;
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        BCK(cudaMemcpy2DAsync(b->rgb[i], b->rpitch, rgb_host + row * b->row0[i], row, row, b->row0[i + 1] - b->row0[i],
                              cudaMemcpyHostToDevice, b->st[i]));
    }
    return bands_sync(b);
}

// enqueue one band's work on its device: start event, (staged halo copies where peer access is missing,) kernel, stop event
static int bands_enqueue(ompimg_bands *b, int i, int mode, int no_bias) {
    BCK(cudaSetDevice(b->dev[i]));
    const int rows = b->row0[i + 1] - b->row0[i];
    BCK(cudaEventRecord(b->e0[i], b->st[i]));
    const uint8_t *above = nullptr, *below = nullptr;
    if (i > 0) {
        const uint8_t *peer_row = b->rgb[i - 1] + (size_t)(b->row0[i] - b->row0[i - 1] - 1) * b->rpitch;
        if (b->direct[i][0]) above = peer_row;
        else { BCK(cudaMemcpyPeerAsync(b->halo[i], b->dev[i], peer_row, b->dev[i - 1], b->rpitch, b->st[i])); above = b->halo[i]; }
    }
    if (i + 1 < b->ndev) {
        const uint8_t *peer_row = b->rgb[i + 1];
        if (b->direct[i][1]) below = peer_row;
        else { BCK(cudaMemcpyPeerAsync(b->halo[i] + b->rpitch, b->dev[i], peer_row, b->dev[i + 1], b->rpitch, b->st[i])); below = b->halo[i] + b->rpitch; }
    }
    StencilArgs a{};
    a.src = b->rgb[i]; a.dst = b->edge[i]; a.src_pitch = b->rpitch; a.dst_pitch = b->epitch;
    a.width = b->w; a.height = rows; a.channels = b->ch; a.nframes = 1; a.fma_mode = mode;
    a.halo_above = above; a.halo_below = below;
    a.no_bias = no_bias;
    BCK(launch_sobel(a, b->sm_count[i], b->st[i]));
    BCK(cudaEventRecord(b->e1[i], b->st[i]));
    return OMPIMG_OK;
}

// The bands' kernels are short (0.1 ms on 8 GPUs for 32768^2), so the launches must not queue up behind one host
// thread: every device gets its own launching thread (env OMPIMG_BANDS_THREADS=0: one thread, for A/B runs).
extern "C" int ompimg_bands_rgb_to_edge(ompimg_bands *b, float *ms_max) {
    if (!b) return OMPIMG_ERR_ARG;
    const int mode = ompimg_get_luma_mode(), no_bias = sqrt_exact_flag();
    static int threaded = -1;
    if (threaded < 0) { const char *e = getenv("OMPIMG_BANDS_THREADS"); threaded = e ? atoi(e) : 1; }
    int rcs[16] = {};
    if (threaded && b->ndev > 1) {
        std::vector<std::thread> th;
        for (int i = 0; i < b->ndev; ++i) th.emplace_back([&, i] { rcs[i] = bands_enqueue(b, i, mode, no_bias); });
        for (auto &t : th) t.join();
    } else {
        for (int i = 0; i < b->ndev; ++i) rcs[i] = bands_enqueue(b, i, mode, no_bias);
    }
    for (int i = 0; i < b->ndev; ++i) if (rcs[i]) return rcs[i];
    float mx = 0.f;
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        BCK(cudaEventSynchronize(b->e1[i]));
        float ms = 0.f;
        BCK(cudaEventElapsedTime(&ms, b->e0[i], b->e1[i]));
        if (ms > mx) mx = ms;
    }
    BCK(cudaSetDevice(b->home));
    if (ms_max) *ms_max = mx;
    return OMPIMG_OK;
}

extern "C" int ompimg_bands_download_edge(ompimg_bands *b, uint8_t *edge_host) {
    if (!b || !edge_host) return OMPIMG_ERR_ARG;
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        int y0 = b->row0[i], y1 = b->row0[i + 1];
        if (y0 < 1) y0 = 1;
        if (y1 > b->h - 1) y1 = b->h - 1;
        if (y1 <= y0) continue;
        BCK(cudaMemcpy2DAsync(edge_host + (size_t)y0 * b->w + 1, b->w, b->edge[i] + (size_t)(y0 - b->row0[i]) * b->epitch + 1, b->epitch,
                              b->w - 2, y1 - y0, cudaMemcpyDeviceToHost, b->st[i]));
    }
    return bands_sync(b);
}

extern "C" int ompimg_bands_checksum(ompimg_bands *b, unsigned long long *sum) {
    if (!b || !sum) return OMPIMG_ERR_ARG;
    unsigned long long total = 0;
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        int y0 = b->row0[i], y1 = b->row0[i + 1];
        if (y0 < 1) y0 = 1;
        if (y1 > b->h - 1) y1 = b->h - 1;
        BCK(cudaMemsetAsync(b->sum[i], 0, 8, b->st[i]));
        if (y1 > y0) {
            count_launch();
            byte_sum_kernel<<<b->sm_count[i] * 8, 256, 0, b->st[i]>>>(b->edge[i], b->epitch, 1, b->w - 1, y0 - b->row0[i], y1 - b->row0[i], b->sum[i]);
            BCK(cudaGetLastError());
        }
    }
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        unsigned long long s = 0;
        BCK(cudaMemcpyAsync(&s, b->sum[i], 8, cudaMemcpyDeviceToHost, b->st[i]));
        BCK(cudaStreamSynchronize(b->st[i]));
        total += s;
    }
    BCK(cudaSetDevice(b->home));
    *sum = total;
    return OMPIMG_OK;
}

// ---- row-band forms of the old/ kernels ------------------------------------------------------------------------
// a6 (old/parallel_convolution.c:12-23) over bands.  The banded image is the reference's pre-padded, top-left anchored input:
// output row r needs input rows r .. r+K-1, so a band needs the K-1 rows BELOW it, which live in the next device's band (or
// in several, when bands are shorter than K-1 rows).  They are copied over NVLink into the tail of this band's buffer with
// cudaMemcpyPeerAsync (P2P, no host staging) on the band's stream, then the band's convolution runs on its device: one
// exchange step, no collective.  Output row r stays on the device that holds input row r.
static int bands_conv_enqueue(ompimg_bands *b, int i, int K, const int32_t *taps, int32_t norm) {
    BCK(cudaSetDevice(b->dev[i]));
    const int rows = b->row0[i + 1] - b->row0[i];
    const int out_h = b->h - K + 1, out_w = b->w - K + 1;
    const int nout = (b->row0[i + 1] < out_h ? b->row0[i + 1] : out_h) - b->row0[i];
    BCK(cudaEventRecord(b->e0[i], b->st[i]));
    if (nout > 0) {
        const int need_end = b->row0[i] + nout + K - 1;                 // one past the last input row of this band's outputs
        for (int j = i + 1; j < b->ndev && b->row0[j] < need_end; ++j) {
            const int r0 = b->row0[j], r1 = b->row0[j + 1] < need_end ? b->row0[j + 1] : need_end;
            uint8_t *dst = b->rgb[i] + (size_t)(r0 - b->row0[i]) * b->rpitch;
            if (b->dev[j] == b->dev[i]) BCK(cudaMemcpyAsync(dst, b->rgb[j], (size_t)(r1 - r0) * b->rpitch, cudaMemcpyDeviceToDevice, b->st[i]));
            else BCK(cudaMemcpyPeerAsync(dst, b->dev[i], b->rgb[j], b->dev[j], (size_t)(r1 - r0) * b->rpitch, b->st[i]));
        }
        BCK(launch_conv(b->rgb[i], b->rpitch, nout, out_w, K, taps, norm, b->conv[i], b->cpitch, 1, b->sm_count[i], b->st[i]));
    }
    (void)rows;
    BCK(cudaEventRecord(b->e1[i], b->st[i]));
    return OMPIMG_OK;
}

static int bands_collect_ms(ompimg_bands *b, float *ms_max) {
    float mx = 0.f;
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        BCK(cudaEventSynchronize(b->e1[i]));
        float ms = 0.f;
        BCK(cudaEventElapsedTime(&ms, b->e0[i], b->e1[i]));
        if (ms > mx) mx = ms;
    }
    BCK(cudaSetDevice(b->home));
    if (ms_max) *ms_max = mx;
    return OMPIMG_OK;
}

extern "C" int ompimg_bands_conv_u8(ompimg_bands *b, int K, const int32_t *taps, int32_t norm, float *ms_max) {
    if (!b || !taps || b->ch != 3 || K < 1 || K > BANDS_TAIL_ROWS + 1 || norm == 0 || b->w < K || b->h < K) return OMPIMG_ERR_ARG;
    const size_t cpitch = ((size_t)(b->w - K + 1) * 3 + 127) / 128 * 128;
    if (!b->conv[0] || b->conv_k != K) {
        for (int i = 0; i < b->ndev; ++i) {
            BCK(cudaSetDevice(b->dev[i]));
            if (b->conv[i]) { BCK(cudaFree(b->conv[i])); b->conv[i] = nullptr; }
            BCK(cudaMalloc(&b->conv[i], cpitch * (size_t)(b->row0[i + 1] - b->row0[i])));
        }
        b->cpitch = cpitch; b->conv_k = K;
    }
    int rcs[16] = {};
    if (b->ndev > 1) {
        std::vector<std::thread> th;
        for (int i = 0; i < b->ndev; ++i) th.emplace_back([&, i] { rcs[i] = bands_conv_enqueue(b, i, K, taps, norm); });
        for (auto &t : th) t.join();
    } else {
        rcs[0] = bands_conv_enqueue(b, 0, K, taps, norm);
    }
    for (int i = 0; i < b->ndev; ++i) if (rcs[i]) return rcs[i];
    return bands_collect_ms(b, ms_max);
}

extern "C" int ompimg_bands_download_conv(ompimg_bands *b, uint8_t *out_host) {
    if (!b || !out_host || !b->conv[0]) return OMPIMG_ERR_ARG;
    const int K = b->conv_k, out_h = b->h - K + 1;
    const size_t row = (size_t)(b->w - K + 1) * 3;
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        const int y0 = b->row0[i], y1 = b->row0[i + 1] < out_h ? b->row0[i + 1] : out_h;
        if (y1 <= y0) continue;
        BCK(cudaMemcpy2DAsync(out_host + row * y0, row, b->conv[i], b->cpitch, row, y1 - y0, cudaMemcpyDeviceToHost, b->st[i]));
    }
    return bands_sync(b);
}

// sum of all bytes of the banded convolution output (the known-answer check of a full-size run)
extern "C" int ompimg_bands_conv_checksum(ompimg_bands *b, unsigned long long *sum) {
    if (!b || !sum || !b->conv[0]) return OMPIMG_ERR_ARG;
    const int K = b->conv_k, out_h = b->h - K + 1, row = (b->w - K + 1) * 3;
    unsigned long long total = 0;
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        const int y0 = b->row0[i], y1 = b->row0[i + 1] < out_h ? b->row0[i + 1] : out_h;
        BCK(cudaMemsetAsync(b->sum[i], 0, 8, b->st[i]));
        if (y1 > y0) {
            count_launch();
            byte_sum_kernel<<<b->sm_count[i] * 8, 256, 0, b->st[i]>>>(b->conv[i], b->cpitch, 0, row, 0, y1 - y0, b->sum[i]);
            BCK(cudaGetLastError());
        }
    }
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        unsigned long long s = 0;
        BCK(cudaMemcpyAsync(&s, b->sum[i], 8, cudaMemcpyDeviceToHost, b->st[i]));
        BCK(cudaStreamSynchronize(b->st[i]));
        total += s;
    }
    BCK(cudaSetDevice(b->home));
    *sum = total;
    return OMPIMG_OK;
}

// a7 (old/parallel_avg_pixel.c:17-30) over bands: every device sums its band (partial sums), the host adds the ndev triples --
// 24 bytes per device, the reduction the reference does with its `reduction(+:avgs)` clause.
extern "C" int ompimg_bands_channel_sums(ompimg_bands *b, unsigned long long *sums3, float *ms_max) {
    if (!b || !sums3 || b->ch != 3) return OMPIMG_ERR_ARG;
    unsigned long long *acc[16] = {};
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        BCK(cudaMalloc(&acc[i], 32));
        BCK(cudaMemsetAsync(acc[i], 0, 32, b->st[i]));
        BCK(cudaEventRecord(b->e0[i], b->st[i]));
        BCK(launch_channel_sums(b->rgb[i], b->rpitch, b->row0[i + 1] - b->row0[i], b->w, acc[i], b->sm_count[i], b->st[i]));
        BCK(cudaEventRecord(b->e1[i], b->st[i]));
    }
    sums3[0] = sums3[1] = sums3[2] = 0;
    for (int i = 0; i < b->ndev; ++i) {
        BCK(cudaSetDevice(b->dev[i]));
        unsigned long long part[3] = {};
        BCK(cudaMemcpyAsync(part, acc[i], 24, cudaMemcpyDeviceToHost, b->st[i]));
        BCK(cudaStreamSynchronize(b->st[i]));
        for (int c = 0; c < 3; ++c) sums3[c] += part[c];
        BCK(cudaFree(acc[i]));
    }
    return bands_collect_ms(b, ms_max);
}

// ---- one process per GPU: the neighbour's band is another PROCESS's allocation --------------------------------
// CUDA IPC maps it into this process (cudaIpcOpenMemHandle enables peer access lazily), after which the address is
// used exactly like a same-process peer pointer: the stencil kernel's bulk copies fetch the halo row over NVLink.
// The 64-byte handle travels by whatever side channel the launcher has (bench.py: torch.distributed all_gather).
static_assert(sizeof(cudaIpcMemHandle_t) == OMPIMG_IPC_HANDLE_BYTES, "IPC handle size");
extern "C" int ompimg_ipc_export(const void *dev_ptr, unsigned char *handle_out) {
    if (!dev_ptr || !handle_out) return OMPIMG_ERR_ARG;
    cudaIpcMemHandle_t h;
    BCK(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle_out, &h, sizeof h);
    return OMPIMG_OK;
}
extern "C" void *ompimg_ipc_open(const unsigned char *handle) {
    if (!handle) return nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        fprintf(stderr, "libompimg_b200: cudaIpcOpenMemHandle failed: %s\n", cudaGetErrorString(e));
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" int ompimg_ipc_close(void *mapped) {
    if (!mapped) return OMPIMG_OK;
    BCK(cudaIpcCloseMemHandle(mapped));
    return OMPIMG_OK;
}
