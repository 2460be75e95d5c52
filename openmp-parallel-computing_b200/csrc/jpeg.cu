// jpeg.cu -- the JPEG decode step in front of the kernel (monolithic/src/main.c:21 decodes with the
// vendored stb_image v2.30; the decoded pixels are what every kernel of the path starts from).
//
// Split: the entropy-coded segments are inherently serial (Huffman codes of unknown length, DC prediction, progressive
// refinement passes) and stay on the host -- a from-scratch ITU T.81 baseline + progressive decoder that produces the
// quantised coefficients of every 8x8 block.  Everything after that is per-block / per-pixel arithmetic and runs on the
// GPU: dequantisation, the inverse DCT, chroma upsampling and YCbCr -> RGB.  Those three steps are where decoders
// differ in their low bits, so they restate stb_image's integer arithmetic exactly (its SSE2 paths are bit-identical to its
// scalar code by construction, stb_image.h:2527-2529): the decoded image must equal the reference's byte for byte,
// because the grayscale golden PNG (images/test_gray.png) was produced from stb-decoded pixels.
//   dequantise     coefficient * table entry, truncated to 16 bits           stb_image.h:2227,2245,3073-3078
//   IDCT           12-bit fixed-point LLM, columns >>10 (+512), rows >>17    stb_image.h:2430-2524
//   upsampling     h2v2: (3*near+far) triangle, ends (t+2)>>2                stb_image.h:3505-3527 (h2v1 :3475-3501, h1v2 :3464-3473)
//   near/far rows  output row j: near = j/2, far = j/2 -+ 1 clamped          stb_image.h:3907-3941
//   YCbCr -> RGB   20-bit fixed point, the Cb->G term masked to 16 bits      stb_image.h:3659-3683
// The checker for all of it is the reference's own stb_image.h, compiled as it lies, in the test infrastructure.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "jpeg.h"
#include "kernels.h"

namespace ompimg {

// =================================================================================================
// host: markers, Huffman tables, entropy decoding
// =================================================================================================
namespace {

const uint8_t ZIGZAG[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                            41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                            30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool present = false;
    uint8_t bits[17] = {};        // codes of each length 1..16
    uint8_t vals[256] = {};
    int mincode[17], maxcode[18], valptr[17];
    int16_t fast[512];            // 9-bit prefix -> (length << 8) | symbol, or -1
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        for (int i = 0; i < 512; ++i) fast[i] = -1;
        code = 0; k = 0;
        for (int l = 1; l <= 9; ++l) {
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                const int first = code << (9 - l);
                for (int j = 0; j < (1 << (9 - l)); ++j) fast[first + j] = (int16_t)((l << 8) | vals[k]);
            }
            code <<= 1;
        }
    }
};

struct BitReader {
    const uint8_t *p, *end;
    uint64_t acc = 0;             // bits are consumed from the top
    int nbits = 0;
    int marker = 0;               // a marker met inside the entropy-coded data (0 = none): from then on zeros are fed
    const uint8_t *marker_at = nullptr;   // its 0xFF byte (the end of the data for the synthetic EOI)
    void reset_bits() { acc = 0; nbits = 0; }
    void fill() {
        while (nbits <= 56) {
            int b = 0;
            if (!marker && p < end) {
                b = *p++;
                if (b == 0xFF) {
                    int c = p < end ? *p++ : 0xD9;
                    while (c == 0xFF && p < end) c = *p++;          // fill bytes
                    if (c != 0) { marker = c; marker_at = p - 2; b = 0; }
                }
            } else if (!marker) {
                marker = 0xD9; marker_at = end;
            }
            acc |= (uint64_t)b << (56 - nbits);
            nbits += 8;
        }
    }
    inline uint32_t peek(int n) { if (nbits < n) fill(); return (uint32_t)(acc >> (64 - n)); }
    inline void skip(int n) { acc <<= n; nbits -= n; }
    inline uint32_t get(int n) { if (n == 0) return 0; const uint32_t v = peek(n); skip(n); return v; }
    inline int bit() { return (int)get(1); }
    inline int decode(const Huff &h) {
        if (nbits < 16) fill();
        const int f = h.fast[acc >> 55];
        if (f >= 0) { skip(f >> 8); return f & 255; }
        for (int l = 10; l <= 16; ++l) {                  // T.81 F.2.2.3: the first length whose largest code is not below the prefix
            const int code = (int)(acc >> (64 - l));
            if (code <= h.maxcode[l]) { skip(l); return h.vals[h.valptr[l] + code - h.mincode[l]]; }
        }
        return -1;
    }
    // T.81 F.2.2.1 EXTEND(RECEIVE(n), n)
    inline int receive_extend(int n) {
        if (n == 0) return 0;
        const int v = (int)get(n);
        return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
    }
};

struct Scan { int ncomp; int comp[4]; int ss, se, ah, al; };

struct Parser {
    JpegCoeffs &out;
    std::string &err;
    const uint8_t *data;
    size_t size, pos = 2;
    Huff dc[4], ac[4];
    int restart_interval = 0;
    int dc_pred[4] = {};
    int eob_run = 0;
    bool have_frame = false;
    Parser(JpegCoeffs &o, std::string &e, const uint8_t *d, size_t n) : out(o), err(e), data(d), size(n) {}

    bool fail(const char *m) { err = m; return false; }
    int be16(size_t at) const { return (data[at] << 8) | data[at + 1]; }

    bool frame(int marker, size_t at, int len) {
        if (len < 8 || data[at] != 8) return fail("only 8-bit precision JPEG is supported");
        out.progressive = marker == 0xC2;
        out.height = be16(at + 1); out.width = be16(at + 3);
        out.ncomp = data[at + 5];
        if (!out.width || !out.height) return fail("JPEG without dimensions");
        if (out.ncomp != 1 && out.ncomp != 3) return fail("JPEG with 1 or 3 components expected (CMYK/YCCK is not supported)");
        if (len < 6 + 3 * out.ncomp) return fail("bad SOF length");
        out.hmax = out.vmax = 1;
        int rgb = 0;
        for (int i = 0; i < out.ncomp; ++i) {
            JpegComp &c = out.comp[i];
            c.id = data[at + 6 + 3 * i];
            if (out.ncomp == 3 && c.id == "RGB"[i]) ++rgb;
            c.h = data[at + 7 + 3 * i] >> 4; c.v = data[at + 7 + 3 * i] & 15; c.tq = data[at + 8 + 3 * i];
            if (c.h < 1 || c.h > 2 || c.v < 1 || c.v > 2 || c.tq > 3) return fail("sampling factors other than 1 and 2 are not supported");
            out.hmax = c.h > out.hmax ? c.h : out.hmax;
            out.vmax = c.v > out.vmax ? c.v : out.vmax;
        }
        out.ids_are_rgb = rgb == 3;
        out.mcux = (out.width + 8 * out.hmax - 1) / (8 * out.hmax);
        out.mcuy = (out.height + 8 * out.vmax - 1) / (8 * out.vmax);
        for (int i = 0; i < out.ncomp; ++i) {
            JpegComp &c = out.comp[i];
            c.x = (out.width * c.h + out.hmax - 1) / out.hmax;            // effective samples (stb_image.h:3323-3324)
            c.y = (out.height * c.v + out.vmax - 1) / out.vmax;
            c.bw = out.mcux * c.h; c.bh = out.mcuy * c.v;                  // blocks incl. the MCU padding (:3332-3333)
            c.coeff.assign((size_t)c.bw * c.bh * 64, 0);
        }
        have_frame = true;
        return true;
    }

    // ---- one block of a baseline scan (T.81 F.2.2): raw coefficient values, de-zigzagged; dequantisation happens on the device
    bool block_baseline(BitReader &br, int16_t *blk, int ci, const Huff &hd, const Huff &ha) {
        const int t = br.decode(hd);
        if (t < 0 || t > 15) return fail("bad Huffman code");
        dc_pred[ci] += br.receive_extend(t);
        blk[0] = (int16_t)dc_pred[ci];
        for (int k = 1; k < 64;) {
            const int rs = br.decode(ha);
            if (rs < 0) return fail("bad Huffman code");
            const int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (rs != 0xF0) break;
                k += 16;
            } else {
                k += r;
                if (k > 63) return fail("coefficient index out of range");
                blk[ZIGZAG[k++]] = (int16_t)br.receive_extend(s);
            }
        }
        return true;
    }
    // ---- progressive (T.81 G.1.2)
    bool block_prog_dc(BitReader &br, int16_t *blk, int ci, const Huff &hd, const Scan &sc) {
        if (sc.ah == 0) {
            const int t = br.decode(hd);
            if (t < 0 || t > 15) return fail("bad Huffman code");
            dc_pred[ci] += br.receive_extend(t);
            blk[0] = (int16_t)(dc_pred[ci] * (1 << sc.al));
        } else if (br.bit()) {
            blk[0] = (int16_t)(blk[0] + (1 << sc.al));
        }
        return true;
    }
    bool block_prog_ac(BitReader &br, int16_t *blk, const Huff &ha, const Scan &sc) {
        if (sc.ah == 0) {
            if (eob_run) { --eob_run; return true; }
            for (int k = sc.ss; k <= sc.se;) {
                const int rs = br.decode(ha);
                if (rs < 0) return fail("bad Huffman code");
                const int s = rs & 15, r = rs >> 4;
                if (s == 0) {
                    if (r < 15) {
                        eob_run = (1 << r) - 1;
                        if (r) eob_run += (int)br.get(r);
                        break;
                    }
                    k += 16;
                } else {
                    k += r;
                    if (k > 63) return fail("coefficient index out of range");
                    blk[ZIGZAG[k++]] = (int16_t)(br.receive_extend(s) * (1 << sc.al));
                }
            }
            return true;
        }
        // refinement of coefficients that already have a value (G.1.2.3)
        const int16_t bit = (int16_t)(1 << sc.al);
        auto refine = [&](int16_t &p) {
            if (br.bit() && (p & bit) == 0) p = (int16_t)(p > 0 ? p + bit : p - bit);
        };
        if (eob_run) {
            --eob_run;
            for (int k = sc.ss; k <= sc.se; ++k) { int16_t &p = blk[ZIGZAG[k]]; if (p) refine(p); }
            return true;
        }
        for (int k = sc.ss; k <= sc.se;) {
            const int rs = br.decode(ha);
            if (rs < 0) return fail("bad Huffman code");
            int s = rs & 15, r = rs >> 4;
            if (s == 0) {
                if (r < 15) {
                    eob_run = (1 << r) - 1;
                    if (r) eob_run += (int)br.get(r);
                    r = 64;                                // to the end of the band: only refinements follow
                }
            } else {
                if (s != 1) return fail("bad Huffman code");
                s = br.bit() ? bit : -bit;
            }
            while (k <= sc.se) {
                int16_t &p = blk[ZIGZAG[k++]];
                if (p) refine(p);
                else {
                    if (r == 0) { p = (int16_t)s; break; }
                    --r;
                }
            }
        }
        return true;
    }

    bool scan(size_t at, int len, size_t *next) {
        if (!have_frame) return fail("SOS before SOF");
        Scan sc{};
        sc.ncomp = data[at];
        if (sc.ncomp < 1 || sc.ncomp > out.ncomp || len < 4 + 2 * sc.ncomp) return fail("bad SOS");
        int hd[4], ha[4];
        for (int i = 0; i < sc.ncomp; ++i) {
            const int id = data[at + 1 + 2 * i], q = data[at + 2 + 2 * i];
            int which = -1;
            for (int c = 0; c < out.ncomp; ++c) if (out.comp[c].id == id) which = c;
            if (which < 0) return fail("bad SOS component");
            sc.comp[i] = which; hd[i] = q >> 4; ha[i] = q & 15;
            if (hd[i] > 3 || ha[i] > 3) return fail("bad Huffman table index");
        }
        sc.ss = data[at + 1 + 2 * sc.ncomp]; sc.se = data[at + 2 + 2 * sc.ncomp];
        sc.ah = data[at + 3 + 2 * sc.ncomp] >> 4; sc.al = data[at + 3 + 2 * sc.ncomp] & 15;
        if (out.progressive) {
            if (sc.ss > 63 || sc.se > 63 || sc.ss > sc.se || sc.ah > 13 || sc.al > 13) return fail("bad SOS");
            if (sc.ss == 0 && sc.se != 0) return fail("progressive scan mixes DC and AC");
            if (sc.ss != 0 && sc.ncomp != 1) return fail("interleaved progressive AC scan");
        } else {
            sc.ss = 0; sc.se = 63; sc.ah = sc.al = 0;
        }
        BitReader br{data + at + len - 2, data + size};
        auto restart = [&]() { br.reset_bits(); for (int &d : dc_pred) d = 0; eob_run = 0; };
        restart();
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        // returns false when the scan has to stop (no restart marker where one is due: stb keeps what it has)
        auto mcu_done = [&]() -> bool {
            if (--todo > 0) return true;
            br.fill();
            if (br.marker < 0xD0 || br.marker > 0xD7) return false;
            br.marker = 0;
            restart();
            todo = restart_interval ? restart_interval : 0x7fffffff;
            return true;
        };
        auto one = [&](int i, int bx, int by) -> bool {
            JpegComp &c = out.comp[sc.comp[i]];
            int16_t *blk = c.coeff.data() + ((size_t)by * c.bw + bx) * 64;
            if (!out.progressive) {
                if (!dc[hd[i]].present || !ac[ha[i]].present) return fail("missing Huffman table");
                return block_baseline(br, blk, sc.comp[i], dc[hd[i]], ac[ha[i]]);
            }
            if (sc.ss == 0) { if (!dc[hd[i]].present && sc.ah == 0) return fail("missing Huffman table"); return block_prog_dc(br, blk, sc.comp[i], dc[hd[i]], sc); }
            if (!ac[ha[i]].present) return fail("missing Huffman table");
            return block_prog_ac(br, blk, ac[ha[i]], sc);
        };
        bool go = true;
        if (sc.ncomp == 1) {
            // a non-interleaved scan covers the component's own blocks only (T.81 A.2.2; stb_image.h:2961-2962)
            JpegComp &c = out.comp[sc.comp[0]];
            const int w = (c.x + 7) >> 3, h = (c.y + 7) >> 3;
            for (int j = 0; j < h && go; ++j)
                for (int i = 0; i < w && go; ++i) {
                    if (!one(0, i, j)) return false;
                    go = mcu_done();
                }
        } else {
            for (int my = 0; my < out.mcuy && go; ++my)
                for (int mx = 0; mx < out.mcux && go; ++mx) {
                    for (int i = 0; i < sc.ncomp; ++i) {
                        const JpegComp &c = out.comp[sc.comp[i]];
                        for (int y = 0; y < c.v; ++y)
                            for (int x = 0; x < c.h; ++x)
                                if (!one(i, mx * c.h + x, my * c.v + y)) return false;
                    }
                    go = mcu_done();
                }
        }
        // continue at the marker that ended the entropy-coded data
        br.fill();
        if (br.marker) {
            *next = (size_t)(br.marker_at - data);
        } else {
            size_t p = (size_t)(br.p - data);
            while (p + 1 < size && !(data[p] == 0xFF && data[p + 1] != 0 && data[p + 1] != 0xFF)) ++p;
            *next = p;
        }
        return true;
    }

    bool run() {
        if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) return fail("not a JPEG (no SOI)");
        out.jfif = false; out.app14_transform = -1;
        for (int t = 0; t < 4; ++t) for (int i = 0; i < 64; ++i) out.dequant[t][i] = 0;
        while (pos + 1 < size) {
            if (data[pos] != 0xFF) { ++pos; continue; }
            const int m = data[pos + 1];
            if (m == 0xFF) { ++pos; continue; }
            if (m == 0x00 || m == 0x01 || (m >= 0xD0 && m <= 0xD7)) { pos += 2; continue; }
            if (m == 0xD9) break;
            if (pos + 3 >= size) return fail("truncated JPEG");
            const int len = be16(pos + 2);
            if (len < 2 || pos + 2 + len > size) return fail("truncated JPEG");
            const size_t at = pos + 4;
            switch (m) {
                case 0xC0: case 0xC1: case 0xC2:
                    if (!frame(m, at, len - 2)) return false;
                    break;
                case 0xC3: case 0xC5: case 0xC6: case 0xC7: case 0xC9: case 0xCA: case 0xCB: case 0xCD: case 0xCE: case 0xCF:
                    return fail("lossless / hierarchical / arithmetic-coded JPEG is not supported");
                case 0xC4: {
                    size_t q = at;
                    const size_t e = pos + 2 + len;
                    while (q < e) {
                        if (q + 17 > e) return fail("bad DHT");
                        const int tc = data[q] >> 4, th = data[q] & 15;
                        if (tc > 1 || th > 3) return fail("bad DHT");
                        Huff &h = tc ? ac[th] : dc[th];
                        int n = 0;
                        for (int l = 1; l <= 16; ++l) { h.bits[l] = data[q + l]; n += h.bits[l]; }
                        if (n > 256 || q + 17 + n > e) return fail("bad DHT");
                        memcpy(h.vals, data + q + 17, n);
                        h.present = true;
                        h.build();
                        q += 17 + n;
                    }
                    break;
                }
                case 0xDB: {
                    size_t q = at;
                    const size_t e = pos + 2 + len;
                    while (q < e) {
                        const int p16 = data[q] >> 4, t = data[q] & 15;
                        if (p16 > 1 || t > 3 || q + 1 + (p16 ? 128 : 64) > e) return fail("bad DQT");
                        ++q;
                        for (int i = 0; i < 64; ++i) { out.dequant[t][ZIGZAG[i]] = (uint16_t)(p16 ? be16(q) : data[q]); q += p16 ? 2 : 1; }
                    }
                    break;
                }
                case 0xDD:
                    if (len != 4) return fail("bad DRI");
                    restart_interval = be16(at);
                    break;
                case 0xE0:
                    if (len >= 7 && !memcmp(data + at, "JFIF\0", 5)) out.jfif = true;
                    break;
                case 0xEE:
                    if (len >= 14 && !memcmp(data + at, "Adobe\0", 6)) out.app14_transform = data[at + 11];
                    break;
                case 0xDA: {
                    size_t next = 0;
                    if (!scan(at, len, &next)) return false;
                    pos = next;
                    continue;
                }
                default:
                    break;
            }
            pos += 2 + (size_t)len;
        }
        if (!have_frame) return fail("JPEG without a frame header");
        return true;
    }
};

}  // namespace

bool jpeg_probe(const uint8_t *data, size_t size, int *w, int *h, int *channels) {
    if (size < 4 || data[0] != 0xFF || data[1] != 0xD8) return false;
    size_t pos = 2;
    while (pos + 3 < size) {
        if (data[pos] != 0xFF) { ++pos; continue; }
        const int m = data[pos + 1];
        if (m == 0xFF) { ++pos; continue; }
        if (m == 0x00 || m == 0x01 || (m >= 0xD0 && m <= 0xD8)) { pos += 2; continue; }
        if (m == 0xD9 || m == 0xDA) return false;
        const int len = (data[pos + 2] << 8) | data[pos + 3];
        if (m == 0xC0 || m == 0xC1 || m == 0xC2) {
            if (pos + 9 >= size || data[pos + 4] != 8) return false;
            const int nc = data[pos + 9];
            if (nc != 1 && nc != 3) return false;
            if (h) *h = (data[pos + 5] << 8) | data[pos + 6];
            if (w) *w = (data[pos + 7] << 8) | data[pos + 8];
            if (channels) *channels = nc == 3 ? 3 : 1;
            return true;
        }
        pos += 2 + (size_t)len;
    }
    return false;
}

bool jpeg_entropy_decode(const uint8_t *data, size_t size, JpegCoeffs &out, std::string &err) {
    Parser p(out, err, data, size);
    return p.run();
}

// =================================================================================================
// device: dequantise + IDCT, upsample + colour conversion
// =================================================================================================
struct IdctParams {
    const int16_t *coeff;       // [blocks][64], natural (de-zigzagged) order
    uint8_t *plane;             // [bh*8][bw*8]
    int bw, nblocks;
    uint16_t dq[64];
};

#define F2F(x) ((int)((x) * 4096 + 0.5))
// one 8-point inverse DCT in stb's fixed-point form (stb_image.h:2430-2465); results in x0..x3 (even part) and t0..t3 (odd part)
#define IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                             \
    int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;                                  \
    p2 = s2; p3 = s6;                                                                        \
    p1 = (p2 + p3) * F2F(0.5411961f);                                                        \
    t2 = p1 + p3 * F2F(-1.847759065f);                                                       \
    t3 = p1 + p2 * F2F(0.765366865f);                                                        \
    p2 = s0; p3 = s4;                                                                        \
    t0 = (p2 + p3) * 4096; t1 = (p2 - p3) * 4096;                                            \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                                  \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                                      \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;                                  \
    p5 = (p3 + p4) * F2F(1.175875602f);                                                      \
    t0 = t0 * F2F(0.298631336f); t1 = t1 * F2F(2.053119869f);                                \
    t2 = t2 * F2F(3.072711026f); t3 = t3 * F2F(1.501321110f);                                \
    p1 = p5 + p1 * F2F(-0.899976223f); p2 = p5 + p2 * F2F(-2.562915447f);                    \
    p3 = p3 * F2F(-1.961570560f); p4 = p4 * F2F(-0.390180644f);                              \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

__device__ __forceinline__ int clamp255(int x) { return x < 0 ? 0 : x > 255 ? 255 : x; }

// one thread per 8x8 block: 64 coefficients in, 64 samples out
__global__ void __launch_bounds__(128) jpeg_idct_kernel(const IdctParams p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.nblocks) return;
    int d[64];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.coeff + (size_t)b * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 v = src[i];
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // coefficient * quantiser, truncated to 16 bits like stb's `short` arithmetic
                d[8 * i + 2 * j] = (int)(short)((int)(short)(w[j] & 0xffff) * (int)p.dq[8 * i + 2 * j]);
                d[8 * i + 2 * j + 1] = (int)(short)((int)(short)(w[j] >> 16) * (int)p.dq[8 * i + 2 * j + 1]);
            }
        }
    }
    int v[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) {          // columns (stb's all-zero-AC shortcut gives the same values as the full form)
        IDCT_1D(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i])
        x0 += 512; x1 += 512; x2 += 512; x3 += 512;
        v[i] = (x0 + t3) >> 10; v[56 + i] = (x0 - t3) >> 10;
        v[8 + i] = (x1 + t2) >> 10; v[48 + i] = (x1 - t2) >> 10;
        v[16 + i] = (x2 + t1) >> 10; v[40 + i] = (x2 - t1) >> 10;
        v[24 + i] = (x3 + t0) >> 10; v[32 + i] = (x3 - t0) >> 10;
    }
    const int bx = b % p.bw, by = b / p.bw;
    uint8_t *o = p.plane + ((size_t)by * 8) * ((size_t)p.bw * 8) + (size_t)bx * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {          // rows
        IDCT_1D(v[8 * i], v[8 * i + 1], v[8 * i + 2], v[8 * i + 3], v[8 * i + 4], v[8 * i + 5], v[8 * i + 6], v[8 * i + 7])
        x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
        const uint32_t lo = (uint32_t)clamp255((x0 + t3) >> 17) | ((uint32_t)clamp255((x1 + t2) >> 17) << 8) |
                            ((uint32_t)clamp255((x2 + t1) >> 17) << 16) | ((uint32_t)clamp255((x3 + t0) >> 17) << 24);
        const uint32_t hi = (uint32_t)clamp255((x3 - t0) >> 17) | ((uint32_t)clamp255((x2 - t1) >> 17) << 8) |
                            ((uint32_t)clamp255((x1 - t2) >> 17) << 16) | ((uint32_t)clamp255((x0 - t3) >> 17) << 24);
        *reinterpret_cast<uint2 *>(o + (size_t)i * p.bw * 8) = make_uint2(lo, hi);
    }
}

struct ColorParams {
    const uint8_t *plane[3];
    int pw[3];                  // plane row pitch (bw*8)
    int cy[3];                  // valid rows of the plane (comp.y)
    int hs[3], vs[3];           // upsampling factors hmax/h, vmax/v
    int w_lores[3];             // (width + hs - 1) / hs
    uint8_t *out;
    size_t out_pitch;
    int width, height, ncomp, out_ch;
    int ycbcr;                  // 1: YCbCr -> RGB, 0: the three components are R, G, B
};

// sample of component k at output pixel (x, j), upsampled exactly like stb_image's row resamplers
__device__ __forceinline__ int comp_sample(const ColorParams &p, int k, int x, int j) {
    const int hs = p.hs[k], vs = p.vs[k];
    const uint8_t *near, *far;
    if (vs == 1) {
        near = far = p.plane[k] + (size_t)j * p.pw[k];
    } else {
        // stb_image.h:3907-3941: odd rows pair chroma row j/2 with the one below, even rows with the one above, clamped
        const int m = j >> 1;
        const int f = (j & 1) ? min(m + 1, p.cy[k] - 1) : max(m - 1, 0);
        near = p.plane[k] + (size_t)m * p.pw[k];
        far = p.plane[k] + (size_t)f * p.pw[k];
    }
    if (hs == 1) return vs == 1 ? near[x] : (3 * near[x] + far[x] + 2) >> 2;
    const int w = p.w_lores[k];
    if (vs == 1) {
        // stbi__resample_row_h_2 (stb_image.h:3475-3501), including its asymmetric last pair
        if (w == 1) return near[0];
        const int i = x >> 1;
        if (x == 0) return near[0];
        if (x == 2 * w - 1) return near[w - 1];
        if (x == 2 * w - 2) return (near[w - 2] * 3 + near[w - 1] + 2) >> 2;
        return (x & 1) ? (3 * near[i] + 2 + near[i + 1]) >> 2 : (3 * near[i] + 2 + near[i - 1]) >> 2;
    }
    // stbi__resample_row_hv_2 (stb_image.h:3505-3527)
    if (x == 0) return (3 * near[0] + far[0] + 2) >> 2;
    if (x == 2 * w - 1) return (3 * near[w - 1] + far[w - 1] + 2) >> 2;
    const int i = (x + 1) >> 1;                         // x = 2i-1 or 2i
    const int ta = 3 * near[i - 1] + far[i - 1], tb = 3 * near[i] + far[i];
    return (x & 1) ? (3 * ta + tb + 8) >> 4 : (3 * tb + ta + 8) >> 4;
}

#define FLOAT2FIXED(x) (((int)((x) * 4096.0f + 0.5f)) << 8)
__global__ void __launch_bounds__(256) jpeg_color_kernel(const ColorParams p) {
    const long long total = (long long)p.width * p.height;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx / p.width), x = (int)(idx - (long long)j * p.width);
        uint8_t *o = p.out + (size_t)j * p.out_pitch + (size_t)x * p.out_ch;
        const int y = comp_sample(p, 0, x, j);
        if (p.ncomp == 1) { o[0] = (uint8_t)y; continue; }
        const int c1 = comp_sample(p, 1, x, j) & 255, c2 = comp_sample(p, 2, x, j) & 255;
        if (!p.ycbcr) { o[0] = (uint8_t)y; o[1] = (uint8_t)c1; o[2] = (uint8_t)c2; continue; }
        // stbi__YCbCr_to_RGB_row (stb_image.h:3660-3683)
        const int y_fixed = ((y & 255) << 20) + (1 << 19);
        const int cb = c1 - 128, cr = c2 - 128;
        int r = y_fixed + cr * FLOAT2FIXED(1.40200f);
        int g = y_fixed + (cr * -FLOAT2FIXED(0.71414f)) + ((cb * -FLOAT2FIXED(0.34414f)) & 0xffff0000);
        int b = y_fixed + cb * FLOAT2FIXED(1.77200f);
        o[0] = (uint8_t)clamp255(r >> 20);
        o[1] = (uint8_t)clamp255(g >> 20);
        o[2] = (uint8_t)clamp255(b >> 20);
    }
}

size_t jpeg_device_scratch_bytes(const JpegCoeffs &jc) {
    size_t n = 0;
    for (int i = 0; i < jc.ncomp; ++i) n += (size_t)jc.comp[i].bw * jc.comp[i].bh * 64 * (sizeof(int16_t) + 1) + 512;
    return n;
}

// coefficients (host, pinned or pageable) -> RGB (or the single plane) at `out` on the device.  `scratch` holds the
// uploaded coefficients and the component planes (jpeg_device_scratch_bytes).
cudaError_t launch_jpeg_backend(const JpegCoeffs &jc, uint8_t *scratch, uint8_t *out, size_t out_pitch, int sm_count, cudaStream_t st) {
    ColorParams cp{};
    uint8_t *q = scratch;
    for (int i = 0; i < jc.ncomp; ++i) {
        const JpegComp &c = jc.comp[i];
        const size_t nblocks = (size_t)c.bw * c.bh, cbytes = nblocks * 64 * sizeof(int16_t);
        int16_t *dcoef = reinterpret_cast<int16_t *>(q);
        q += (cbytes + 255) & ~(size_t)255;
        uint8_t *plane = q;
        q += (nblocks * 64 + 255) & ~(size_t)255;
        cudaError_t e = cudaMemcpyAsync(dcoef, c.coeff.data(), cbytes, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        IdctParams ip{};
        ip.coeff = dcoef; ip.plane = plane; ip.bw = c.bw; ip.nblocks = (int)nblocks;
        for (int k = 0; k < 64; ++k) ip.dq[k] = jc.dequant[c.tq][k];
        count_launch();
        jpeg_idct_kernel<<<(unsigned)((nblocks + 127) / 128), 128, 0, st>>>(ip);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        cp.plane[i] = plane; cp.pw[i] = c.bw * 8; cp.cy[i] = c.y;
        cp.hs[i] = jc.hmax / c.h; cp.vs[i] = jc.vmax / c.v;
        cp.w_lores[i] = (jc.width + cp.hs[i] - 1) / cp.hs[i];
    }
    cp.out = out; cp.out_pitch = out_pitch; cp.width = jc.width; cp.height = jc.height; cp.ncomp = jc.ncomp;
    cp.out_ch = jc.ncomp == 3 ? 3 : 1;
    // stb_image.h:3879: components are RGB already if their ids spell "RGB", or Adobe says "no transform" and there is no JFIF marker
    cp.ycbcr = jc.ncomp == 3 && !(jc.ids_are_rgb || (jc.app14_transform == 0 && !jc.jfif));
    const long long total = (long long)jc.width * jc.height;
    const int grid = (int)std::max(1LL, std::min((total + 255) / 256, (long long)sm_count * 32));
    count_launch();
    jpeg_color_kernel<<<grid, 256, 0, st>>>(cp);
    return cudaGetLastError();
}

}  // namespace ompimg
