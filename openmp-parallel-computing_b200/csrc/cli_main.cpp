// cli_main.cpp -- `ompimg_grayscale`: the reference's CLI (monolithic/src/main.c:14-49; the binary its Flask and RabbitMQ
// workers exec per request, microservices/grayscale/app.py:24-33, event-driven/grayscale_service/app.py:59-63) with the
// codec steps either side of the kernel replaced as well:
//
//     ompimg_grayscale <input_img> <output_img.png> [passes]        same argv, same stdout line, same exit codes
//
//   decode   main.c:21 stbi_load        ->  ompimg_jpeg_gray_host: entropy decoding on the host, IDCT / upsampling / colour
//   kernel   main.c:33-35 passes loop   ->  ... and the `passes` grayscale passes on the device, one download
//   encode   main.c:41 stbi_write_png   ->  PNG written here: Sub filter, deflate split over the host cores (pigz-style
//                                           independent blocks, Adler-32 combined) -- stb's single-threaded compressor is
//                                           most of the reference's wall time on `test.jpg`
// With OMPIMG_WORKER=<socket> the process creates no CUDA context: the JPEG file goes to the persistent worker and the
// pixels come back (worker.cu).  The PNG decodes to exactly the pixels of the reference's PNG (it is not the same byte
// stream: the reference's is whatever stb's compressor emits).  Inputs other than the JPEGs the decoder handles are refused
// with the reference's own error line and exit code 1.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <thread>
#include <vector>

#include <zlib.h>

#include "../../include/ompimg_b200.h"

namespace {

void put32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::vector<uint8_t> &png, const char *type, const uint8_t *data, size_t n) {
    put32(png, (uint32_t)n);
    const size_t at = png.size();
    png.insert(png.end(), type, type + 4);
    if (n) png.insert(png.end(), data, data + n);
    put32(png, (uint32_t)crc32(0, png.data() + at, (uInt)(n + 4)));
}

// rows -> zlib stream, compressed as independent raw-deflate pieces on `nthreads` threads
bool deflate_parallel(const std::vector<uint8_t> &raw, int nthreads, std::vector<uint8_t> &out) {
    const size_t piece = 1u << 20;
    const size_t npieces = raw.empty() ? 1 : (raw.size() + piece - 1) / piece;
    std::vector<std::vector<uint8_t>> parts(npieces);
    std::vector<uLong> adlers(npieces);
    std::vector<int> ok(npieces, 0);
    auto work = [&](size_t first, size_t step) {
        for (size_t i = first; i < npieces; i += step) {
            const size_t lo = i * piece, n = raw.size() - lo < piece ? raw.size() - lo : piece;
            z_stream zs{};
            if (deflateInit2(&zs, 1, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return;
            parts[i].resize(deflateBound(&zs, (uLong)n) + 16);
            zs.next_in = const_cast<Bytef *>(raw.data() + lo); zs.avail_in = (uInt)n;
            zs.next_out = parts[i].data(); zs.avail_out = (uInt)parts[i].size();
            const int rc = deflate(&zs, i + 1 == npieces ? Z_FINISH : Z_FULL_FLUSH);     // a full flush ends on a byte boundary with no back-references across it
            const bool good = (i + 1 == npieces ? rc == Z_STREAM_END : rc == Z_OK) && zs.avail_in == 0;
            parts[i].resize(parts[i].size() - zs.avail_out);
            deflateEnd(&zs);
            adlers[i] = adler32(adler32(0, nullptr, 0), raw.data() + lo, (uInt)n);
            ok[i] = good;
        }
    };
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > npieces) nthreads = (int)npieces;
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, (size_t)t, (size_t)nthreads);
    work(0, (size_t)nthreads);
    for (auto &t : th) t.join();
    out.clear();
    out.push_back(0x78); out.push_back(0x01);
    uLong ad = adler32(0, nullptr, 0);
    for (size_t i = 0; i < npieces; ++i) {
        if (!ok[i]) return false;
        out.insert(out.end(), parts[i].begin(), parts[i].end());
        const size_t lo = i * piece, n = raw.size() - lo < piece ? raw.size() - lo : piece;
        ad = adler32_combine(ad, adlers[i], (z_off_t)n);
    }
    put32(out, (uint32_t)ad);
    return true;
}

bool write_png(const char *path, const uint8_t *px, int w, int h, int ch) {
    const size_t row = (size_t)w * ch;
    std::vector<uint8_t> raw((row + 1) * h);
    for (int y = 0; y < h; ++y) {                     // filter type 1 (Sub): each byte minus the same channel of the pixel before
        uint8_t *o = raw.data() + (size_t)y * (row + 1);
        const uint8_t *s = px + (size_t)y * row;
        *o++ = 1;
        for (int i = 0; i < ch && (size_t)i < row; ++i) o[i] = s[i];
        for (size_t i = ch; i < row; ++i) o[i] = (uint8_t)(s[i] - s[i - ch]);
    }
    std::vector<uint8_t> z;
    unsigned hc = std::thread::hardware_concurrency();
    if (const char *e = getenv("OMPIMG_PNG_THREADS")) hc = (unsigned)atoi(e);
    if (!deflate_parallel(raw, (int)(hc ? (hc > 16 ? 16 : hc) : 4), z)) return false;
    std::vector<uint8_t> png = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w, (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h,
                        8, (uint8_t)(ch == 1 ? 0 : ch == 2 ? 4 : ch == 3 ? 2 : 6), 0, 0, 0};
    chunk(png, "IHDR", ihdr, 13);
    chunk(png, "IDAT", z.data(), z.size());
    chunk(png, "IEND", nullptr, 0);
    FILE *f = fopen(path, "wb");
    if (!f) return false;
    const bool good = fwrite(png.data(), 1, png.size(), f) == png.size();
    return (fclose(f) == 0) && good;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "Uso: %s <input_img> <output_img.png> [passaggi_kernel]\n", argv[0]);      // main.c:16
        return 1;
    }
    std::vector<uint8_t> file;
    if (FILE *f = fopen(argv[1], "rb")) {
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (n > 0) { file.resize((size_t)n); if (fread(file.data(), 1, file.size(), f) != file.size()) file.clear(); }
        fclose(f);
    }
    int w = 0, h = 0, ch = 0;
    if (file.empty() || ompimg_jpeg_probe(file.data(), file.size(), &w, &h, &ch) != OMPIMG_OK) {
        fprintf(stderr, "Errore caricando immagine\n");                                             // main.c:23
        if (!file.empty()) fprintf(stderr, "ompimg_grayscale: %s\n", ompimg_last_error());
        return 1;
    }
    int passes = argc >= 4 ? atoi(argv[3]) : 1;
    if (passes < 1) passes = 1;
    std::vector<uint8_t> px((size_t)w * h * ch);
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (ompimg_jpeg_gray_host(file.data(), file.size(), passes, px.data(), px.size(), &w, &h, &ch) != OMPIMG_OK) {
        fprintf(stderr, "libompimg_b200: decode + grayscale failed: %s\n", ompimg_last_error());
        return 1;                                                                                   // non-zero exit -> HTTP 500 / CalledProcessError
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    printf("Compute kernel \xC3\x97%d: %.4f s\n", passes, (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) / 1e9);   // main.c:39
    if (!write_png(argv[2], px.data(), w, h, ch)) {
        fprintf(stderr, "Errore nel salvataggio\n");                                                // main.c:42
        return 1;
    }
    return 0;
}
