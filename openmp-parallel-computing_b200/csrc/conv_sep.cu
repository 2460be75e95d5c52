// conv_sep.cu -- host side of the rank-1 convolution path: range checks, tap words, launch geometry.
// Two kernels with the same arithmetic: conv_tma.cuh (16-byte aligned base and pitch: warp-autonomous, tensor-map copies;
// instantiated per K in conv_tma_k{3,5,7,9}.cu) and conv_sep.cuh (any 4-byte aligned layout, e.g. the reference's dense
// (cols+K-1)*3 pitch: producer warp + 1-D bulk copies; conv_sep_k{3,5,7,9}.cu).
#include <cstdlib>

#include "conv_sep.cuh"
#include "conv_tma.cuh"
#include "tma.h"

namespace ompimg {

constexpr int CONV_SEP_MAX_K = 9;

// Returns false when the taps / layout are outside this path's ranges (caller uses conv.cu's kernels).
bool launch_conv_sep(const uint8_t *in, size_t in_pitch, int rows, int cols, int K, const int32_t *u, const int32_t *v, int32_t norm,
                     void *out, size_t out_pitch, int out_is_u8, int sm_count, cudaStream_t st, cudaError_t *err) {
    if (!(K == 3 || K == 5 || K == 7 || K == 9)) return false;
    if ((((uintptr_t)in | in_pitch) & 3) != 0) return false;
    const unsigned long long in_row = (unsigned long long)(cols + K - 1) * 3ull;
    if (in_row > 0x7fffffffull || in_pitch < in_row) return false;
    long long su = 0, sv = 0, mu = 0, mv = 0;
    bool nonneg = norm > 0;
    for (int i = 0; i < K; ++i) {
        const long long au = u[i] < 0 ? -(long long)u[i] : u[i], av = v[i] < 0 ? -(long long)v[i] : v[i];
        su += au; sv += av; mu = au > mu ? au : mu; mv = av > mv ? av : mv;
        nonneg = nonneg && u[i] >= 0 && v[i] >= 0;
    }
    const long long acc_max = 255ll * su * sv;
    if (acc_max >= (1ll << 31)) return false;
    int div = -1, shift = 0;
    uint32_t magic = 0;
    if (nonneg && mu <= 255 && mv <= 255 && 255ll * sv <= 65535) {
        if ((norm & (norm - 1)) == 0) {
            while ((1 << shift) < norm) ++shift;
            div = out_is_u8 && shift == 8 ? 3 : out_is_u8 && shift == 16 ? 4 : 0;
        } else {
            // floor(a / norm) == umulhi(a, m), m = ceil(2^32 / norm), for every a with a * (m*norm - 2^32) < 2^32
            const unsigned long long m = ((1ull << 32) + (unsigned long long)norm - 1) / (unsigned long long)norm;
            const unsigned long long e = m * (unsigned long long)norm - (1ull << 32);
            if (m < (1ull << 32) && (unsigned long long)acc_max * e < (1ull << 32)) { div = 1; magic = (uint32_t)m; }
        }
    }
    if (div < 0) {
        if (mu <= 127 && mv <= 127 && 255ll * sv <= 32767) div = 2;
        else return false;
    }
    // constant vertical taps (box blur): fold the constant into v and let the kernel keep running window sums
    int32_t vv[CONV_SEP_MAX_K];
    for (int i = 0; i < K; ++i) vv[i] = v[i];
    int vbox = 0;
    if (div != 2 && nonneg && u[0] >= 1) {
        bool flat = true;
        for (int i = 1; i < K; ++i) flat = flat && u[i] == u[0];
        const long long c = u[0], wmax = 255ll * sv * c * K;                // largest window sum
        if (flat && mv * c <= 255 && wmax <= 65535) {
            if ((norm & (norm - 1)) == 0) { vbox = 1; div = 0; }
            else {
                const unsigned long long m = ((1ull << 32) + (unsigned long long)norm - 1) / (unsigned long long)norm;
                // floor(s/norm) == umulhi(s << 16 | t, m) >> 16 for all s, t <= wmax (see the kernel)
                if (m < (1ull << 32) && wmax + (((wmax + 1) * norm) >> 16) + 1 < 65536) { vbox = 1; div = 1; magic = (uint32_t)m; }
            }
            if (vbox) for (int i = 0; i < K; ++i) vv[i] = (int32_t)(v[i] * c);
        }
    }

    uint32_t th[4][3], ut[5];
    for (int i = 0; i < 4; ++i)
        for (int m = 0; m < 3; ++m) {
            uint32_t wv = 0;
            for (int b = 0; b < 4; ++b) {
                const int k = 4 * m + b - i;
                if (k >= 0 && k < K) wv |= (uint32_t)(vv[k] & 255) << (8 * b);
            }
            th[i][m] = wv;
        }
    auto ub = [&](int k) -> uint32_t { return k >= 0 && k < K ? (uint32_t)(u[k] & 255) : 0u; };
    for (int q = 0; q < 5; ++q) ut[q] = ub(2 * q) | ub(2 * q + 1) << 8 | ub(2 * q - 1) << 16 | ub(2 * q) << 24;

    // unit-ended unsigned taps (every binomial, the plain box): the kernel variants with fewer IDP (ConvLane's E1)
    static int e1_env = -1;
    if (e1_env < 0) { const char *e = getenv("OMPIMG_CONV_E1"); e1_env = e ? atoi(e) : 1; }
    const int ends1 = e1_env && div != 2 && nonneg && vv[0] == 1 && vv[K - 1] == 1 && (vbox || (u[0] == 1 && u[K - 1] == 1));

    // ---- 16-byte aligned base and pitch, unsigned taps: the warp-autonomous kernel fed by tensor-map copies (conv_tma.cuh)
    static int tma_env = -1;
    if (tma_env < 0) { const char *e = getenv("OMPIMG_CONV_TMA"); tma_env = e ? atoi(e) : 1; }
    const unsigned long long in_row8 = (in_row + 7ull) & ~7ull;
    if (tma_env && div != 2 && (((uintptr_t)in | in_pitch) & 15) == 0 && in_pitch >= in_row8) {
        if (EncodeTiledFn enc = encode_tiled_fn()) {
            ConvTmaParams q{};
            const cuuint64_t dims[2] = {in_row8 / 8, (cuuint64_t)(rows + K - 1)};
            const cuuint64_t strides[1] = {in_pitch};
            const cuuint32_t box[2] = {ct_box_bytes(K) / 8, 2}, estr[2] = {1, 1};
            if (enc(&q.map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<uint8_t *>(in), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                q.out = reinterpret_cast<uint8_t *>(out); q.out_pitch = out_pitch;
                q.rows = rows; q.cols = cols;
                q.nstrips = (cols + CT_WPX - 1) / CT_WPX;
                q.out_wide = (((uintptr_t)out | out_pitch) & 15) == 0;
                for (int i = 0; i < 4; ++i) for (int m = 0; m < 3; ++m) q.th[i][m] = th[i][m];
                for (int k = 0; k < 5; ++k) q.ut[k] = ut[k];
                q.magic = magic; q.shift = shift;
                switch (K) {
                    case 3: *err = launch_conv_tma_k<3>(q, out_is_u8, div, vbox, ends1, sm_count, st); break;
                    case 5: *err = launch_conv_tma_k<5>(q, out_is_u8, div, vbox, ends1, sm_count, st); break;
                    case 7: *err = launch_conv_tma_k<7>(q, out_is_u8, div, vbox, ends1, sm_count, st); break;
                    default: *err = launch_conv_tma_k<9>(q, out_is_u8, div, vbox, ends1, sm_count, st); break;
                }
                return true;
            }
        }
    }

    ConvSepParams p{};
    p.in = in; p.out = reinterpret_cast<uint8_t *>(out); p.in_pitch = in_pitch; p.out_pitch = out_pitch;
    p.in_end = in + (size_t)(rows + K - 2) * in_pitch + (size_t)in_row;
    p.rows = rows; p.cols = cols;
    p.in_row_bytes = (uint32_t)in_row;
    p.magic = magic; p.shift = shift; p.norm = norm;
    p.out_wide = (((uintptr_t)out | out_pitch) & 15) == 0;
    for (int i = 0; i < 4; ++i) for (int m = 0; m < 3; ++m) p.th[i][m] = th[i][m];
    for (int q = 0; q < 5; ++q) p.ut[q] = ut[q];

    p.nw = min(CS_NW, (cols + CS_WPX - 1) / CS_WPX);
    const int SW = p.nw * CS_WPX;
    const int nstrips = (cols + SW - 1) / SW;
    if (nstrips > 65535) return false;
    p.row_slot = (uint32_t)((16 + (SW + K - 1) * 3 + 16 + 127) & ~127);
    p.stage_bytes = 2 * p.row_slot;
    const int np = (K + 1) / 2;
    p.stages = np >= 4 ? np : 2 * np;                      // a multiple of (K+1)/2, see the kernel
    if (const char *e = getenv("OMPIMG_CONV_STAGES")) p.stages = np * max(1, min(4, atoi(e)));      // tuning knobs (tools/prof.py)
    p.relaxed = 1;
    if (const char *e = getenv("OMPIMG_CONV_RELAX")) p.relaxed = atoi(e);
    const int ob = out_is_u8 ? 1 : 4;
    const int smem = p.stages * (int)p.stage_bytes + 16 * p.stages + 128 + p.nw * 2 * CS_WPX * 3 * ob;
    // bands: an even number of rows, sized so that the CTAs fill a whole number of waves of resident CTAs
    const int occ = K <= 5 ? CS_MINB5 : CS_MINB9;
    const long long slots = (long long)sm_count * occ;
    int best_rows = rows, best_score = 1 << 30;
    for (int wv = 1; wv <= 64; ++wv) {
        const long long nb = (wv * slots) / nstrips;
        if (nb < 1) continue;
        int br = (int)((rows + nb - 1) / nb);
        br += br & 1;
        if (br < 2 * K) break;
        const int score = br > 96 ? br - 96 : 96 - br;
        if (score < best_score) { best_score = score; best_rows = br; }
    }
    if (const char *e = getenv("OMPIMG_CONV_BAND")) best_rows = max(2, atoi(e));
    best_rows += best_rows & 1;
    p.band_rows = best_rows;
    const int nbands = (rows + p.band_rows - 1) / p.band_rows;
    const dim3 grid((unsigned)nbands, (unsigned)nstrips);
    switch (K) {
        case 3: *err = launch_conv_sep_k<3>(p, out_is_u8, div, vbox, ends1, grid, smem, st); break;
        case 5: *err = launch_conv_sep_k<5>(p, out_is_u8, div, vbox, ends1, grid, smem, st); break;
        case 7: *err = launch_conv_sep_k<7>(p, out_is_u8, div, vbox, ends1, grid, smem, st); break;
        default: *err = launch_conv_sep_k<9>(p, out_is_u8, div, vbox, ends1, grid, smem, st); break;
    }
    return true;
}

}  // namespace ompimg
