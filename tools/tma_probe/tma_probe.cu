// tools/tma_probe/tma_probe.cu -- which small tensor-map boxes does the copy engine accept?  (development probe, not product)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, uint8_t *out, int c0, int c1, int c2, uint32_t bytes) {
    __shared__ __align__(128) uint8_t buf[1024];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(d), "l"(reinterpret_cast<uint64_t>(&map)), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
    }
    __syncthreads();
    uint32_t ok = 0;
    for (int i = 0; i < 100000 && !ok; ++i)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b) : "memory");
    if (threadIdx.x < 256) out[threadIdx.x] = ok ? buf[threadIdx.x] : 0xEE;
}

int main(int argc, char **argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int idx = -1;
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) { printf("no encoder\n"); return 1; }
    EncodeTiledFn enc = (EncodeTiledFn)f;
    const int H = 64, F = 2;
    struct Cfg { const char *name; CUtensorMapDataType t; int es; int dim0; size_t pitch; int box0; CUtensorMapL2promotion l2; int c0; };
    const Cfg cfgs[] = {
        {"u64 box2 dim5 pitch48 l2none c0=1", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 5, 48, 2, CU_TENSOR_MAP_L2_PROMOTION_NONE, 1},
        {"u64 box2 dim5 pitch48 l2none c0=0", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 5, 48, 2, CU_TENSOR_MAP_L2_PROMOTION_NONE, 0},
        {"u64 box2 dim6 pitch48 l2none c0=2", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 6, 48, 2, CU_TENSOR_MAP_L2_PROMOTION_NONE, 2},
        {"u64 box2 dim5 pitch48 l2_128 c0=1", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 5, 48, 2, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, 1},
        {"u64 box4 dim9 pitch80 l2none c0=1", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 9, 80, 4, CU_TENSOR_MAP_L2_PROMOTION_NONE, 1},
        {"u32 box4 dim10 pitch48 l2none c0=2", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 10, 48, 4, CU_TENSOR_MAP_L2_PROMOTION_NONE, 2},
        {"u8 box16 dim40 pitch48 l2none c0=8", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 40, 48, 16, CU_TENSOR_MAP_L2_PROMOTION_NONE, 8},
        {"u64 box2 dim5 pitch48 l2none c0=3 (half out of range)", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 5, 48, 2, CU_TENSOR_MAP_L2_PROMOTION_NONE, 4},
        {"u64 box4 dim6 pitch48 l2none c0=1 (32-byte box)", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 6, 48, 4, CU_TENSOR_MAP_L2_PROMOTION_NONE, 1},
        {"u64 box4 dim6 pitch64 l2none c0=2 (32-byte box, 32B-aligned)", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 6, 64, 4, CU_TENSOR_MAP_L2_PROMOTION_NONE, 2},
        {"u64 box2 dim6 pitch64 l2none c0=2 (16-byte box at 16B-aligned start)", CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, 6, 64, 2, CU_TENSOR_MAP_L2_PROMOTION_NONE, 2},
        {"u32 box4 dim12 pitch48 c0=4 (16-byte box, u32, aligned)", CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 12, 48, 4, CU_TENSOR_MAP_L2_PROMOTION_NONE, 4},
        {"u8 box32 dim48 pitch48 c0=8 (32-byte box, u8, 8B start)", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 48, 48, 32, CU_TENSOR_MAP_L2_PROMOTION_NONE, 8},
        {"u8 box32 dim48 pitch48 c0=16 (32-byte box, u8, 16B start)", CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 48, 48, 32, CU_TENSOR_MAP_L2_PROMOTION_NONE, 16},
    };
    for (const Cfg &c : cfgs) {
        if (++idx != only && only >= 0) continue;
        uint8_t *g, *o;
        cudaMalloc(&g, c.pitch * H * F);
        cudaMalloc(&o, 256);
        std::vector<uint8_t> h(c.pitch * H * F);
        for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 8));
        cudaMemcpy(g, h.data(), h.size(), cudaMemcpyHostToDevice);
        CUtensorMap m;
        const cuuint64_t dims[3] = {(cuuint64_t)c.dim0, H, F};
        const cuuint64_t strides[2] = {c.pitch, c.pitch * H};
        const cuuint32_t box[3] = {(cuuint32_t)c.box0, 3, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&m, c.t, 3, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, c.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-60s encode failed %d\n", c.name, (int)r); continue; }
        const uint32_t bytes = (uint32_t)(c.box0 * c.es * 3);
        probe<<<1, 256>>>(m, o, c.c0, 5, 1, bytes);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-60s KERNEL ERROR: %s\n", c.name, cudaGetErrorString(e)); return 2; }
        uint8_t res[256];
        cudaMemcpy(res, o, 256, cudaMemcpyDeviceToHost);
        int bad = 0;
        const int rowb = c.box0 * c.es;
        for (int r3 = 0; r3 < 3; ++r3)
            for (int j = 0; j < rowb; ++j) {
                const size_t col = (size_t)c.c0 * c.es + j;
                const uint8_t want = col < (size_t)c.dim0 * c.es ? h[(size_t)1 * c.pitch * H + (size_t)(5 + r3) * c.pitch + col] : 0;
                if (res[r3 * rowb + j] != want) ++bad;
            }
        printf("%-60s ok, %d bytes, mismatches %d (first bytes %02x %02x)\n", c.name, bytes, bad, res[0], res[1]);
        cudaFree(g); cudaFree(o);
    }
    return 0;
}
