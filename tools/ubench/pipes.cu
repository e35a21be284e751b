// tools/ubench/pipes.cu -- which issue pipe does each SASS op of the stitch kernels use, and at what rate?
// (diagnostic only; not part of the product).  Each test runs 8 independent dependency chains per thread,
// 32 warps per SM, and reports warp-instructions per cycle per SM sub-partition (SMSP).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
#define CH 8
#define DEF_KERNEL(name, INIT, BODY)                                                            \
__global__ void __launch_bounds__(1024) name(uint32_t* out, long long* cyc, uint32_t seed) {    \
    uint32_t x[CH]; float f[CH]; uint64_t w[CH];                                                 \
    _Pragma("unroll") for (int i = 0; i < CH; i++) { x[i] = seed + threadIdx.x * 7 + i; f[i] = (float)(x[i] & 255); w[i] = x[i]; } \
    uint32_t c1 = seed * 3 + 1, c2 = seed + 5; float g1 = 1.0001f + seed, g2 = 0.5f; (void)c1; (void)c2; (void)g1; (void)g2; \
    INIT                                                                                         \
    __shared__ unsigned long long s_t0, s_t1; if (threadIdx.x == 0) { s_t0 = ~0ull; s_t1 = 0; } __syncthreads(); \
    long long t0 = clock64();                                                                    \
    _Pragma("unroll 8") for (int it = 0; it < ITERS; it++) {                                     \
        _Pragma("unroll") for (int i = 0; i < CH; i++) { BODY }                                  \
    }                                                                                            \
    long long t1 = clock64();                                                                    \
    uint32_t acc = 0;                                                                            \
    _Pragma("unroll") for (int i = 0; i < CH; i++) acc += x[i] + __float_as_uint(f[i]) + (uint32_t)w[i] + (uint32_t)(w[i] >> 32); \
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                            \
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_t0, (unsigned long long)t0); atomicMax(&s_t1, (unsigned long long)t1); } __syncthreads(); \
    if (threadIdx.x == 0) cyc[blockIdx.x] = (long long)(s_t1 - s_t0);                            \
}
#define A_IMAD   asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
#define A_IMADW  asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(c1));
#define A_IMADHI asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
#define A_DP2A   asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
#define A_DP4A   asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
#define A_PRMT   asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(x[i]) : "r"(c1));
#define A_LOP    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(c1), "r"(c2));
#define A_SHR    asm volatile("shr.u32 %0, %0, 3;" : "+r"(x[i]));
#define A_SHF    asm volatile("shf.r.wrap.b32 %0, %0, %1, 5;" : "+r"(x[i]) : "r"(c1));
#define A_IADD   asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c1));
#define A_IADD3  asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[i]) : "r"(c1), "r"(c2));
#define A_MNMX   asm volatile("min.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(c1));
#define A_RELU   asm volatile("min.s32.relu %0, %0, %1;" : "+r"(x[i]) : "r"(c1));
#define A_I2F    asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(x[i])); x[i] ^= __float_as_uint(f[i]);
#define A_I2FP   asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(x[i]));
#define A_F2I    asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(x[i]) : "f"(f[i]));
#define A_FADD   asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g1));
#define A_FADDRM asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g1));
#define A_FMUL   asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g1));
#define A_FFMA   asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g1), "f"(g2));
#define A_FMNMX  asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g1));
#define A_LDS    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x[i]) : "r"((x[i] & 0xFFCu) + sbase));
#define A_LDS64  asm volatile("{.reg .u32 t; ld.shared.v2.u32 {%0, t}, [%1]; xor.b32 %0, %0, t;}" : "=r"(x[i]) : "r"((x[i] & 0xFF8u) + sbase));
#define A_BFE    asm volatile("bfe.u32 %0, %0, 5, 9;" : "+r"(x[i]));
#define A_BFI    asm volatile("bfi.b32 %0, %1, %0, 8, 8;" : "+r"(x[i]) : "r"(c1));
#define A_SHLADD asm volatile("mad.lo.u32 %0, %0, 256, %1;" : "+r"(x[i]) : "r"(c1));
#define A_HFMA2  asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c1), "r"(c2));
#define A_FFMA2  asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[i]) : "l"(pk1), "l"(pk2));
#define A_FADD2RM asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(pk1));
#define A_FMUL2  asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(w[i]) : "l"(pk1));
#define PKINIT unsigned long long pk1, pk2; asm("mov.b64 %0, {%1, %2};" : "=l"(pk1) : "f"(g1), "f"(g2)); asm("mov.b64 %0, {%1, %2};" : "=l"(pk2) : "f"(g2), "f"(g1));
#define A_VIMAX3 x[i] = (uint32_t)__vimax3_s32((int)x[i], (int)c1, (int)(c2 ^ x[i]));
#define SMEM __shared__ uint32_t sm[1024]; sm[threadIdx.x] = threadIdx.x * 4; __syncthreads(); uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
#define NOINIT
DEF_KERNEL(k_imad, NOINIT, A_IMAD)
DEF_KERNEL(k_imadw, NOINIT, A_IMADW)
DEF_KERNEL(k_imadhi, NOINIT, A_IMADHI)
DEF_KERNEL(k_dp2a, NOINIT, A_DP2A)
DEF_KERNEL(k_dp4a, NOINIT, A_DP4A)
DEF_KERNEL(k_prmt, NOINIT, A_PRMT)
DEF_KERNEL(k_lop, NOINIT, A_LOP)
DEF_KERNEL(k_shr, NOINIT, A_SHR)
DEF_KERNEL(k_shf, NOINIT, A_SHF)
DEF_KERNEL(k_iadd, NOINIT, A_IADD)
DEF_KERNEL(k_iadd3, NOINIT, A_IADD3)
DEF_KERNEL(k_mnmx, NOINIT, A_MNMX)
DEF_KERNEL(k_relu, NOINIT, A_RELU)
DEF_KERNEL(k_i2f, NOINIT, A_I2F)
DEF_KERNEL(k_f2i, NOINIT, A_F2I A_I2FP)
DEF_KERNEL(k_fadd, NOINIT, A_FADD)
DEF_KERNEL(k_faddrm, NOINIT, A_FADDRM)
DEF_KERNEL(k_fmul, NOINIT, A_FMUL)
DEF_KERNEL(k_ffma, NOINIT, A_FFMA)
DEF_KERNEL(k_fmnmx, NOINIT, A_FMNMX)
DEF_KERNEL(k_lds, SMEM, A_LDS)
DEF_KERNEL(k_lds64, SMEM, A_LDS64)
DEF_KERNEL(k_bfe, NOINIT, A_BFE)
DEF_KERNEL(k_bfi, NOINIT, A_BFI)
DEF_KERNEL(k_shladd, NOINIT, A_SHLADD)
DEF_KERNEL(k_hfma2, NOINIT, A_HFMA2)
DEF_KERNEL(k_vimax3, NOINIT, A_VIMAX3)
DEF_KERNEL(k_ffma2, PKINIT, A_FFMA2)
DEF_KERNEL(k_fadd2rm, PKINIT, A_FADD2RM)
DEF_KERNEL(k_fmul2, PKINIT, A_FMUL2)
DEF_KERNEL(k_ffma2_imad, PKINIT, A_FFMA2 A_IMAD)
DEF_KERNEL(k_ffma2_lop, PKINIT, A_FFMA2 A_LOP)
DEF_KERNEL(k_ffma2_ffma, PKINIT, A_FFMA2 A_FFMA)
DEF_KERNEL(k_ffma2_imad_lop, PKINIT, A_FFMA2 A_IMAD A_LOP)
// mixes: same pipe -> rates add; different pipes -> overlap
DEF_KERNEL(k_imad_lop, NOINIT, A_IMAD A_LOP)
DEF_KERNEL(k_imad_ffma, NOINIT, A_IMAD A_FFMA)
DEF_KERNEL(k_imad_fadd, NOINIT, A_IMAD A_FADD)
DEF_KERNEL(k_lop_fadd, NOINIT, A_LOP A_FADD)
DEF_KERNEL(k_imad_dp2a, NOINIT, A_IMAD A_DP2A)
DEF_KERNEL(k_lop_dp2a, NOINIT, A_LOP A_DP2A)
DEF_KERNEL(k_imad_prmt, NOINIT, A_IMAD A_PRMT)
DEF_KERNEL(k_lop_prmt, NOINIT, A_LOP A_PRMT)
DEF_KERNEL(k_imad_mnmx, NOINIT, A_IMAD A_MNMX)
DEF_KERNEL(k_lop_mnmx, NOINIT, A_LOP A_MNMX)
DEF_KERNEL(k_imad_i2f, NOINIT, A_IMAD A_I2F)
DEF_KERNEL(k_lop_i2fp, NOINIT, A_LOP A_I2FP)
DEF_KERNEL(k_imad_i2fp, NOINIT, A_IMAD A_I2FP)
DEF_KERNEL(k_imad_imadw, NOINIT, A_IMAD A_IMADW)
DEF_KERNEL(k_lop_imadw, NOINIT, A_LOP A_IMADW)
DEF_KERNEL(k_lop_fmnmx, NOINIT, A_LOP A_FMNMX)
DEF_KERNEL(k_imad_fmnmx, NOINIT, A_IMAD A_FMNMX)
DEF_KERNEL(k_lop_faddrm, NOINIT, A_LOP A_FADDRM)
DEF_KERNEL(k_imad_faddrm, NOINIT, A_IMAD A_FADDRM)
DEF_KERNEL(k_lop_fmul, NOINIT, A_LOP A_FMUL)
DEF_KERNEL(k_imad_lds, SMEM, A_IMAD A_LDS)
DEF_KERNEL(k_imad_lop_lds, SMEM, A_IMAD A_LOP A_LDS)
DEF_KERNEL(k_imad_lop_lop, NOINIT, A_IMAD A_LOP A_LOP)
DEF_KERNEL(k_imad_imad_lop, NOINIT, A_IMAD A_IMAD A_LOP)
DEF_KERNEL(k_lop_hfma2, NOINIT, A_LOP A_HFMA2)
DEF_KERNEL(k_imad_hfma2, NOINIT, A_IMAD A_HFMA2)
DEF_KERNEL(k_lop_shladd, NOINIT, A_LOP A_SHLADD)
DEF_KERNEL(k_lop_shf, NOINIT, A_LOP A_SHF)

typedef void (*kfn)(uint32_t*, long long*, uint32_t);
struct T { const char* name; kfn f; int n; };
int main() {
    T tests[] = {
        {"IMAD", k_imad, 1}, {"IMAD.WIDE", k_imadw, 1}, {"IMAD.HI", k_imadhi, 1}, {"IDP.2A", k_dp2a, 1}, {"IDP.4A", k_dp4a, 1}, {"PRMT", k_prmt, 1},
        {"LOP3", k_lop, 1}, {"SHR", k_shr, 1}, {"SHF", k_shf, 1}, {"IADD", k_iadd, 1}, {"IADD3(2 adds)", k_iadd3, 1}, {"VIMNMX", k_mnmx, 1}, {"VIMNMX.RELU", k_relu, 1},
        {"I2F+xor", k_i2f, 2}, {"F2I+I2F", k_f2i, 2}, {"FADD", k_fadd, 1}, {"FADD.RM", k_faddrm, 1}, {"FMUL", k_fmul, 1}, {"FFMA", k_ffma, 1}, {"FMNMX", k_fmnmx, 1},
        {"LDS(+lop+add)", k_lds, 1}, {"LDS64(+..)", k_lds64, 1}, {"BFE", k_bfe, 1}, {"BFI", k_bfi, 1}, {"IMAD x256+c", k_shladd, 1}, {"HFMA2", k_hfma2, 1}, {"VIMNMX3", k_vimax3, 1},
        {"IMAD+LOP3", k_imad_lop, 2}, {"IMAD+FFMA", k_imad_ffma, 2}, {"IMAD+FADD", k_imad_fadd, 2}, {"LOP3+FADD", k_lop_fadd, 2}, {"IMAD+IDP.2A", k_imad_dp2a, 2}, {"LOP3+IDP.2A", k_lop_dp2a, 2},
        {"IMAD+PRMT", k_imad_prmt, 2}, {"LOP3+PRMT", k_lop_prmt, 2}, {"IMAD+VIMNMX", k_imad_mnmx, 2}, {"LOP3+VIMNMX", k_lop_mnmx, 2}, {"IMAD+I2F+xor", k_imad_i2f, 3}, {"LOP3+I2FP", k_lop_i2fp, 2}, {"IMAD+I2FP", k_imad_i2fp, 2},
        {"IMAD+IMAD.WIDE", k_imad_imadw, 2}, {"LOP3+IMAD.WIDE", k_lop_imadw, 2}, {"LOP3+FMNMX", k_lop_fmnmx, 2}, {"IMAD+FMNMX", k_imad_fmnmx, 2}, {"LOP3+FADD.RM", k_lop_faddrm, 2}, {"IMAD+FADD.RM", k_imad_faddrm, 2},
        {"LOP3+FMUL", k_lop_fmul, 2}, {"IMAD+LDS", k_imad_lds, 2}, {"IMAD+LOP3+LDS", k_imad_lop_lds, 3}, {"IMAD+LOP3+LOP3", k_imad_lop_lop, 3}, {"IMAD+IMAD+LOP3", k_imad_imad_lop, 3},
        {"FFMA2", k_ffma2, 1}, {"FADD2.RM", k_fadd2rm, 1}, {"FMUL2", k_fmul2, 1}, {"FFMA2+IMAD", k_ffma2_imad, 2}, {"FFMA2+LOP3", k_ffma2_lop, 2},
        {"FFMA2+FFMA", k_ffma2_ffma, 2}, {"FFMA2+IMAD+LOP3", k_ffma2_imad_lop, 3},
        {"LOP3+HFMA2", k_lop_hfma2, 2}, {"IMAD+HFMA2", k_imad_hfma2, 2}, {"LOP3+IMADx256", k_lop_shladd, 2}, {"LOP3+SHF", k_lop_shf, 2},
    };
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* cyc; cudaMalloc(&out, nsm * 1024 * 4); cudaMalloc(&cyc, nsm * 8);
    long long* h = new long long[nsm];
    printf("%-18s %10s %14s  (asm ops per iteration as written; see SASS for the real count)\n", "test", "cycles", "ops/clk/SMSP");
    for (auto& t : tests) {
        t.f<<<nsm, 1024>>>(out, cyc, 0); t.f<<<nsm, 1024>>>(out, cyc, 0);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, nsm * 8, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < nsm; i++) s += h[i]; s /= nsm;
        printf("%-18s %10.0f %14.3f\n", t.name, s, (double)ITERS * CH * t.n * 8 / s);   // 8 warps per SMSP
    }
    cudaError_t e = cudaDeviceSynchronize(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
