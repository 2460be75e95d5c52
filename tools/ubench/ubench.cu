// Instruction-throughput microbenchmark for sm_100a (B200).
// Purpose: measure the per-SM issue rate of the handful of SASS ops the
// Sobel / grayscale kernels are built from, so kernel design (SWAR half2 vs
// fp32, PRMT packing, F2IP saturating pack, MUFU.SQRT) rests on measured
// numbers rather than on guesses.  Not part of the product library.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 4
#define ITERS 1024

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

// Each bench: CHAINS independent dependency chains per thread, UNROLL ops per
// chain per iteration.  OPS = number of SASS instructions of interest per BODY.
#define BENCH_KERNEL(NAME, DECL, BODY, SINK)                                   \
__global__ void __launch_bounds__(1024, 1) NAME(uint32_t* out, long long* cyc, \
                                                 uint32_t seed) {              \
  DECL;                                                                        \
  __syncthreads();                                                             \
  long long t0 = clock64();                                                    \
  for (int it = 0; it < ITERS; ++it) {                                         \
    _Pragma("unroll")                                                          \
    for (int u = 0; u < UNROLL; ++u) { BODY; }                                 \
  }                                                                            \
  long long t1 = clock64();                                                    \
  __syncthreads();                                                             \
  uint32_t s = 0; SINK;                                                        \
  if (s == 0x12345678u) out[threadIdx.x] = s;                                  \
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                             \
}

#define R8(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7)

// ---- fp32 ----
#define D_F(i) float f##i = __uint_as_float(0x3f800000u + seed + threadIdx.x + i); 
#define DECL_F float ka = __uint_as_float(0x3f800001u + seed), kb = __uint_as_float(0x33800000u + seed); R8(D_F)
#define S_F(i) s += __float_as_uint(f##i);
#define B_FFMA(i) asm("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f##i) : "f"(ka), "f"(kb));
BENCH_KERNEL(k_ffma, DECL_F, R8(B_FFMA), R8(S_F))
#define B_FADD(i) asm("add.rn.f32 %0, %0, %1;" : "+f"(f##i) : "f"(kb));
BENCH_KERNEL(k_fadd, DECL_F, R8(B_FADD), R8(S_F))
#define B_FMUL(i) asm("mul.rn.f32 %0, %0, %1;" : "+f"(f##i) : "f"(ka));
BENCH_KERNEL(k_fmul, DECL_F, R8(B_FMUL), R8(S_F))
#define B_FADDRZ(i) asm("add.rz.f32 %0, %0, %1;" : "+f"(f##i) : "f"(kb));
BENCH_KERNEL(k_fadd_rz, DECL_F, R8(B_FADDRZ), R8(S_F))
#define B_FMNMX(i) asm("min.f32 %0, %0, %1;" : "+f"(f##i) : "f"(ka));
BENCH_KERNEL(k_fmnmx, DECL_F, R8(B_FMNMX), R8(S_F))
#define B_MUFU(i) asm("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f##i));
BENCH_KERNEL(k_mufu_sqrt, DECL_F, R8(B_MUFU), R8(S_F))
// fp32x2
#define D_F2(i) unsigned long long g##i = ((unsigned long long)(0x3f800000u + seed + i) << 32) | (0x3f800000u + threadIdx.x);
#define DECL_F2 unsigned long long ka2 = 0x3f8000013f800001ull + seed, kb2 = 0x3380000033800000ull + seed; R8(D_F2)
#define S_F2(i) s += (uint32_t)g##i + (uint32_t)(g##i >> 32);
#define B_FFMA2(i) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(g##i) : "l"(ka2), "l"(kb2));
BENCH_KERNEL(k_ffma2, DECL_F2, R8(B_FFMA2), R8(S_F2))
// ---- half2 ----
#define D_H(i) uint32_t h##i = 0x3c003c00u + seed + threadIdx.x + i;
#define DECL_H uint32_t ha = 0x3c013c01u + seed, hb = 0x14001400u + seed; R8(D_H)
#define S_H(i) s += h##i;
#define B_HFMA2(i) asm("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(h##i) : "r"(ha), "r"(hb));
BENCH_KERNEL(k_hfma2, DECL_H, R8(B_HFMA2), R8(S_H))
#define B_HADD2(i) asm("add.rn.f16x2 %0, %0, %1;" : "+r"(h##i) : "r"(hb));
BENCH_KERNEL(k_hadd2, DECL_H, R8(B_HADD2), R8(S_H))
// mixed f32 += f16*f16
#define DECL_FH unsigned short hx = (unsigned short)(0x3c00u + seed + threadIdx.x); DECL_F
#define B_FHFMA(i) asm("fma.rn.f32.f16 %0, %1, %1, %0;" : "+f"(f##i) : "h"(hx));
BENCH_KERNEL(k_fhfma, DECL_FH, R8(B_FHFMA), R8(S_F))
// ---- integer / logic ----
#define D_I(i) uint32_t x##i = seed * 2654435761u + threadIdx.x * 40503u + i;
#define DECL_I uint32_t ia = seed + 0x00ff00ffu, ib = seed + 0x64006400u; R8(D_I)
#define S_I(i) s += x##i;
#define B_PRMT(i) asm("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(x##i) : "r"(ib));
BENCH_KERNEL(k_prmt, DECL_I, R8(B_PRMT), R8(S_I))
#define B_LOP3(i) asm("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(x##i) : "r"(ia), "r"(ib));
BENCH_KERNEL(k_lop3, DECL_I, R8(B_LOP3), R8(S_I))
#define B_IADD3(i) asm("add.u32 %0, %0, %1;" : "+r"(x##i) : "r"(ia));
BENCH_KERNEL(k_iadd, DECL_I, R8(B_IADD3), R8(S_I))
#define B_IMAD(i) asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x##i) : "r"(ia), "r"(ib));
BENCH_KERNEL(k_imad, DECL_I, R8(B_IMAD), R8(S_I))
#define B_SHF(i) asm("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x##i) : "r"(ia));
BENCH_KERNEL(k_shf, DECL_I, R8(B_SHF), R8(S_I))
#define B_IDP4A(i) asm("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x##i) : "r"(ia), "r"(ib));
BENCH_KERNEL(k_idp4a, DECL_I, R8(B_IDP4A), R8(S_I))
#define B_VIMNMX(i) asm("max.s16x2 %0, %0, %1;" : "+r"(x##i) : "r"(ia));
BENCH_KERNEL(k_vimnmx_s16x2, DECL_I, R8(B_VIMNMX), R8(S_I))
#define B_SHFL(i) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(x##i) : "r"(ia & 31));
BENCH_KERNEL(k_shfl, DECL_I, R8(B_SHFL), R8(S_I))
// ---- conversions ----
// float -> 2 x u8 saturating pack (F2IP.U8.F32.TRUNC)
#define DECL_CV DECL_F uint32_t acc = seed;
#define B_F2IP(i) { uint32_t t_; asm("{ .reg .s32 a_, b_; cvt.rzi.s32.f32 a_, %1; cvt.rzi.s32.f32 b_, %2; cvt.pack.sat.u8.s32.b32 %0, a_, b_, %3; }" : "=r"(t_) : "f"(f##i), "f"(ka), "r"(acc)); f##i = __uint_as_float(t_ | 0x3f000000u); }
BENCH_KERNEL(k_f2ip_plus_lop, DECL_CV, R8(B_F2IP), R8(S_F))
#define B_F2I(i) { int t_; asm("cvt.rzi.s32.f32 %0, %1;" : "=r"(t_) : "f"(f##i)); f##i = __int_as_float(t_ | 0x3f000000); }
BENCH_KERNEL(k_f2i_plus_lop, DECL_F, R8(B_F2I), R8(S_F))
#define B_I2F(i) { float t_; asm("cvt.rn.f32.u32 %0, %1;" : "=f"(t_) : "r"(x##i)); x##i = __float_as_uint(t_); }
BENCH_KERNEL(k_i2f_u32, DECL_I, R8(B_I2F), R8(S_I))
#define B_I2F8(i) { float t_; asm("{ .reg .u16 q_; .reg .u8 b_; cvt.u16.u32 q_, %1; cvt.u8.u16 b_, q_; cvt.rn.f32.u8 %0, b_; }" : "=f"(t_) : "r"(x##i)); x##i = __float_as_uint(t_); }
// ---- shared memory ----
#define DECL_S __shared__ uint4 sm[1024]; sm[threadIdx.x] = make_uint4(0, seed & 0, 0, 0); uint32_t sa = (uint32_t)__cvta_generic_to_shared(&sm[0]) + (threadIdx.x & 31) * 16 + (threadIdx.x >> 5) * 512; R8(D_I)
#define B_LDS32(i) { uint32_t t_; asm("ld.shared.u32 %0, [%1];" : "=r"(t_) : "r"(sa ^ (x##i & 0x10))); x##i += t_; }
BENCH_KERNEL(k_lds32_plus_2alu, DECL_S, R8(B_LDS32), R8(S_I))
#define B_LDS128(i) { uint32_t a_, b_, c_, d_; asm("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a_), "=r"(b_), "=r"(c_), "=r"(d_) : "r"(sa ^ (x##i & 0x10))); x##i += a_ ^ d_; }
BENCH_KERNEL(k_lds128_plus_alu, DECL_S, R8(B_LDS128), R8(S_I))
// ---- dual-pipe mixes ----
#define DECL_MIX DECL_H DECL_F uint32_t ia = seed + 0x00ff00ffu, ib = seed + 0x64006400u; R8(D_I) unsigned short hx = (unsigned short)(0x3c00u + seed);
#define S_MIX(i) s += h##i + x##i + __float_as_uint(f##i);
#define B_MIX_HP(i) B_HFMA2(i) B_PRMT(i)
BENCH_KERNEL(k_mix_hfma2_prmt, DECL_MIX, R8(B_MIX_HP), R8(S_MIX))
#define B_MIX_HHP(i) B_HFMA2(i) B_HADD2(i) B_PRMT(i)
BENCH_KERNEL(k_mix_2half_1prmt, DECL_MIX, R8(B_MIX_HHP), R8(S_MIX))
#define B_MIX_HM(i) B_HFMA2(i) B_HADD2(i) B_HFMA2(i) B_MUFU(i)
BENCH_KERNEL(k_mix_3half_1mufu, DECL_MIX, R8(B_MIX_HM), R8(S_MIX))
#define B_MIX_FP(i) B_FFMA(i) B_PRMT(i)
BENCH_KERNEL(k_mix_ffma_prmt, DECL_MIX, R8(B_MIX_FP), R8(S_MIX))
#define B_MIX_FHM(i) B_FHFMA(i) B_FHFMA(i) B_MUFU(i)
BENCH_KERNEL(k_mix_2fhfma_1mufu, DECL_MIX, R8(B_MIX_FHM), R8(S_MIX))
#define B_MIX_II(i) B_IADD3(i) B_IMAD(i)
BENCH_KERNEL(k_mix_iadd_imad, DECL_MIX, R8(B_MIX_II), R8(S_MIX))
#define B_MIX_LP(i) B_LOP3(i) B_PRMT(i)
BENCH_KERNEL(k_mix_lop3_prmt, DECL_MIX, R8(B_MIX_LP), R8(S_MIX))
// Sobel-like blend: 5 half2-class + 2 FHFMA + 1 PRMT + 1 MUFU + 0.5 F2IP per "pixel"
#define B_MIX_SOBEL(i) B_HFMA2(i) B_HADD2(i) B_PRMT(i) B_HFMA2(i) B_FHFMA(i) B_HADD2(i) B_FHFMA(i) B_MUFU(i) B_LOP3(i) B_HADD2(i)
BENCH_KERNEL(k_mix_sobel_like10, DECL_MIX, R8(B_MIX_SOBEL), R8(S_MIX))


#define B_MIX_FH(i) B_FFMA(i) B_HFMA2(i)
BENCH_KERNEL(k_mix_ffma_hfma2, DECL_MIX, R8(B_MIX_FH), R8(S_MIX))
#define B_MIX_HI(i) B_HFMA2(i) B_LOP3(i)
BENCH_KERNEL(k_mix_hfma2_lop3, DECL_MIX, R8(B_MIX_HI), R8(S_MIX))
#define B_MIX_FHF(i) B_FHFMA(i) B_HFMA2(i)
BENCH_KERNEL(k_mix_fhfma_hfma2, DECL_MIX, R8(B_MIX_FHF), R8(S_MIX))
#define B_MIX_FM(i) B_FFMA(i) B_FFMA(i) B_FFMA(i) B_MUFU(i)
BENCH_KERNEL(k_mix_3ffma_1mufu, DECL_MIX, R8(B_MIX_FM), R8(S_MIX))
#define B_MIX_IMF(i) B_IMAD(i) B_FFMA(i)
BENCH_KERNEL(k_mix_imad_ffma, DECL_MIX, R8(B_MIX_IMF), R8(S_MIX))
#define B_MIX_IMH(i) B_IMAD(i) B_HFMA2(i)
BENCH_KERNEL(k_mix_imad_hfma2, DECL_MIX, R8(B_MIX_IMH), R8(S_MIX))
#define B_MIX_IML(i) B_IMAD(i) B_LOP3(i)
BENCH_KERNEL(k_mix_imad_lop3, DECL_MIX, R8(B_MIX_IML), R8(S_MIX))
// F2IP chained through integer->float reinterpretation only (no extra op)
#define B_F2IPX(i) { uint32_t t_; asm("{ .reg .s32 a_, b_; cvt.rzi.s32.f32 a_, %1; cvt.rzi.s32.f32 b_, %2; cvt.pack.sat.u8.s32.b32 %0, a_, b_, %3; }" : "=r"(t_) : "f"(f##i), "f"(ka), "r"(acc)); f##i = __uint_as_float(t_); }
BENCH_KERNEL(k_f2ip_alone, DECL_CV, R8(B_F2IPX), R8(S_F))
#define B_MIX_MF2(i) B_MUFU(i) B_F2IPX(i)
BENCH_KERNEL(k_mix_mufu_f2ip, DECL_CV, R8(B_MIX_MF2), R8(S_F))

typedef void (*kern_t)(uint32_t*, long long*, uint32_t);
struct Bench { const char* name; kern_t fn; int ops_per_body; };

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  int nsm = p.multiProcessorCount;
  printf("# device %s, %d SMs, clock %d kHz\n", p.name, nsm, p.clockRate);
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, 4096 * 4)); CK(cudaMalloc(&cyc, nsm * 8 * sizeof(long long)));
  Bench benches[] = {
    {"FFMA", k_ffma, 8}, {"FADD", k_fadd, 8}, {"FMUL", k_fmul, 8}, {"FADD.RZ", k_fadd_rz, 8},
    {"FMNMX", k_fmnmx, 8}, {"MUFU.SQRT", k_mufu_sqrt, 8}, {"FFMA2", k_ffma2, 8},
    {"HFMA2", k_hfma2, 8}, {"HADD2", k_hadd2, 8}, {"FHFMA", k_fhfma, 8},
    {"PRMT", k_prmt, 8}, {"LOP3", k_lop3, 8}, {"IADD", k_iadd, 8}, {"IMAD", k_imad, 8}, {"SHF", k_shf, 8},
    {"IDP4A", k_idp4a, 8}, {"VIMNMX.S16x2", k_vimnmx_s16x2, 8}, {"SHFL", k_shfl, 8},
    {"F2IP(+LOP)", k_f2ip_plus_lop, 8}, {"F2I(+LOP)", k_f2i_plus_lop, 8}, {"I2F.U32", k_i2f_u32, 8},
    {"LDS32(+2alu)", k_lds32_plus_2alu, 8}, {"LDS128(+alu)", k_lds128_plus_alu, 8},
    {"mix HFMA2+PRMT", k_mix_hfma2_prmt, 16}, {"mix 2half+PRMT", k_mix_2half_1prmt, 24},
    {"mix 3half+MUFU", k_mix_3half_1mufu, 32}, {"mix FFMA+PRMT", k_mix_ffma_prmt, 16},
    {"mix 2FHFMA+MUFU", k_mix_2fhfma_1mufu, 24}, {"mix IADD+IMAD", k_mix_iadd_imad, 16},
    {"mix LOP3+PRMT", k_mix_lop3_prmt, 16}, {"mix sobel-like10", k_mix_sobel_like10, 80},
    {"mix FFMA+HFMA2", k_mix_ffma_hfma2, 16}, {"mix HFMA2+LOP3", k_mix_hfma2_lop3, 16}, {"mix FHFMA+HFMA2", k_mix_fhfma_hfma2, 16},
    {"mix 3FFMA+MUFU", k_mix_3ffma_1mufu, 32}, {"mix IMAD+FFMA", k_mix_imad_ffma, 16}, {"mix IMAD+HFMA2", k_mix_imad_hfma2, 16},
    {"mix IMAD+LOP3", k_mix_imad_lop3, 16}, {"F2IP alone", k_f2ip_alone, 8}, {"mix MUFU+F2IP", k_mix_mufu_f2ip, 16},
  };
  int threads_list[] = {256, 512, 1024};
  printf("%-20s %8s %14s %16s\n", "op", "threads", "cycles(med)", "thread-ops/clk/SM");
  for (auto& b : benches) {
    for (int threads : threads_list) {
      std::vector<long long> h(nsm);
      for (int rep = 0; rep < 2; ++rep) {
        b.fn<<<nsm, threads>>>(out, cyc, 0);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h.data(), cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost));
      std::sort(h.begin(), h.end());
      long long med = h[nsm / 2];
      double ops = (double)b.ops_per_body * UNROLL * ITERS * threads;
      printf("%-20s %8d %14lld %16.1f\n", b.name, threads, med, ops / (double)med);
    }
  }
  return 0;
}
