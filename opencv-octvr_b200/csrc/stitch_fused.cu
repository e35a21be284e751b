// csrc/stitch_fused.cu -- K_stitch_fused: colour conversion + remap + gain + feather blend + 4:2:0 store of a whole
// frame in one persistent kernel (sm_100a).  Nothing but the input planes, the 8 B/pair tables and the output frame
// touches HBM.  Interface and data layout: kernels.cuh (FusedParams); table construction: mapper.cpp (build_fused).
//
// Arithmetic contracts are those of kernels.cu (bit-exact against the reference's CPU functions, see oracle/):
//   colour   : imgproc/src/color.cpp:6087-6169 (YUV->RGB), :6430-6481 (RGB->YUV 4:2:0)
//   bilinear : imgproc/src/imgwarp.cpp:4383-4442 + :3812-4020 -> (sum S_k a_k b_k + 512) >> 10
//   gain     : core/src/arithm.cpp multiply-by-scalar in f64 -> sat_u8(rint(v*g))
//   feather  : stitching/src/cuda/blender.cu:73-98 (short)(v*W) truncation, blenders.cpp:581 (x 1/N, rint)
//
// Issue-slot budget (measured pipe model of tools/ubench/pipes.cu: ALU ops LOP3/SHF/PRMT/IADD3/VIMNMX and
// "heavy" ops IMAD/IDP each issue at 0.5 / clk / SM sub-partition, FP32 ops at ~1 / clk): the gather loop is
// written so that neither pipe exceeds ~60 % of the issue slots of a pair.
#include "kernels.cuh"
#include "device_common.cuh"

namespace ob {

// ---- shared-memory map of a CTA (dynamic) ----
constexpr int FS_RGBX = 0;                                   // FUSED_CAP u32
constexpr int FS_ENT = FS_RGBX + FUSED_CAP * 4;              // 2 x FT_PX x 8 B (TMA bulk destination, 16-B aligned)
constexpr int FS_META = FS_ENT + 2 * FT_PX * 8;              // 3 x FTileBlock (padded to 1056 B)
constexpr int FS_META_STRIDE = 1056;
constexpr int FS_Y = FS_META + 3 * FS_META_STRIDE;           // 32 x 32 luma bytes
constexpr int FS_U = FS_Y + FT_PX;                           // 16 x 16
constexpr int FS_V = FS_U + FT_PX / 4;
constexpr int FS_GAIN = FS_V + FT_PX / 4;                    // MAX_CAMS x {g32, bias, flag, pad}
constexpr int FS_MBAR = FS_GAIN + MAX_CAMS * 16;             // 2 x u64
constexpr int FS_TOTAL = FS_MBAR + 16;
static_assert(FS_ENT % 128 == 0 && FS_META % 16 == 0 && FS_Y % 16 == 0 && FS_MBAR % 8 == 0, "alignment");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void f_mbar_init(uint32_t mbar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void f_mbar_expect_tx(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_mbar_wait(uint32_t mbar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "FWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra FDONE_%=;\n\t"
        "bra FWAIT_%=;\n\t"
        "FDONE_%=:\n\t}" ::"r"(mbar), "r"(phase) : "memory");
}
// TMA bulk copy (no tensor map): global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// ---- the prefetched input bytes of one conversion item: 8 px x 2 rows of Y, 4 U, 4 V ----
struct ItemRegs { uint2 y0, y1; uint32_t u, v; };

// item -> (row pair, 8-px group) of the job's source box
__device__ __forceinline__ void item_pos(const FJob& J, int item, int& rp, int& gx)
{
    rp = (int)(((uint32_t)item * J.rcp) >> 20);
    gx = item - rp * J.groups;
}

// Item load.  Fast path: the item lies inside the source and the planes allow 64/32-bit loads.  An item entirely
// outside the source needs no bytes (mask 0 -> zeros = BORDER_CONSTANT).  Anything else (image border, unaligned
// planes) is flagged ITEM_SLOW and converted pixel by pixel at conversion time.
constexpr uint32_t ITEM_SLOW = 0x100u;
__device__ __forceinline__ void load_item(const CamSrc& c, const FJob& J, int item, ItemRegs& r, uint32_t& mask)
{
    int rp, gx;
    item_pos(J, item, rp, gx);
    const int x0 = J.bx0 + (gx << 3), y0 = J.by0 + (rp << 1);
    if (x0 + 8 <= 0 || x0 >= c.w || y0 < 0 || y0 >= c.h) { mask = 0u; return; }
    if (c.aligned4 && x0 >= 0 && x0 + 8 <= c.w) {
        const uint8_t* yp = c.y + (size_t)y0 * c.y_pitch + x0;
        r.y0 = __ldg(reinterpret_cast<const uint2*>(yp));
        r.y1 = __ldg(reinterpret_cast<const uint2*>(yp + c.y_pitch));
        if (c.uv_step == 1) {
            r.u = __ldg(reinterpret_cast<const uint32_t*>(c.u + (size_t)(y0 >> 1) * c.u_pitch + (x0 >> 1)));
            r.v = __ldg(reinterpret_cast<const uint32_t*>(c.v + (size_t)(y0 >> 1) * c.v_pitch + (x0 >> 1)));
        } else {                                           // NV12: U0 V0 U1 V1 U2 V2 U3 V3
            const uint2 uv = __ldg(reinterpret_cast<const uint2*>(c.u + (size_t)(y0 >> 1) * c.u_pitch + x0));
            r.u = __byte_perm(uv.x, uv.y, 0x6420);
            r.v = __byte_perm(uv.x, uv.y, 0x7531);
        }
        mask = 0xFFu;
    } else
        mask = ITEM_SLOW;
}

// BT.601 limited-range integer conversion of one pixel (imgproc/src/color.cpp:6087-6169), Y already clamped to >= 16
// and the -16 folded into the chroma terms: R | G << 8 | B << 16
__device__ __forceinline__ uint32_t yuv_px16(uint32_t Yc, int ruv, int guv, int buv)
{
    const int yy = (int)Yc * 1220542;
    const uint32_t r = (uint32_t)__vimin_s32_relu((yy + ruv) >> 20, 255);
    const uint32_t g = (uint32_t)__vimin_s32_relu((yy + guv) >> 20, 255);
    const uint32_t b = (uint32_t)__vimin_s32_relu((yy + buv) >> 20, 255);
    return r + g * 256u + b * 65536u;
}

__device__ __forceinline__ uint32_t vignette_px(uint32_t p, float k)
{
    // cudaarithm mul_mat.cu:198-213 : saturate_cast<uchar>(u8 * f32), round-to-nearest-even
    const int r = clamp255(__float2int_rn((float)(p & 255u) * k));
    const int g = clamp255(__float2int_rn((float)((p >> 8) & 255u) * k));
    const int b = clamp255(__float2int_rn((float)((p >> 16) & 255u) * k));
    return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

// border / unaligned item: byte loads, one pixel at a time (rare)
__device__ __noinline__ void convert_item_slow(const CamSrc& c, int x0, int y0, uint32_t* d0, uint32_t* d1)
{
    for (int k = 0; k < 8; k++) {
        const int x = x0 + k;
        uint32_t a = 0u, b = 0u;
        if (x >= 0 && x < c.w) {
            const int u = (int)__ldg(c.u + (size_t)(y0 >> 1) * c.u_pitch + (size_t)(x >> 1) * c.uv_step) - 128;
            const int v = (int)__ldg(c.v + (size_t)(y0 >> 1) * c.v_pitch + (size_t)(x >> 1) * c.uv_step) - 128;
            constexpr int K16 = 16 * 1220542;
            const int ruv = (1 << 19) - K16 + 1673527 * v, guv = (1 << 19) - K16 - 852492 * v - 409993 * u, buv = (1 << 19) - K16 + 2116026 * u;
            a = yuv_px16(max((uint32_t)__ldg(c.y + (size_t)y0 * c.y_pitch + x), 16u), ruv, guv, buv);
            b = yuv_px16(max((uint32_t)__ldg(c.y + (size_t)(y0 + 1) * c.y_pitch + x), 16u), ruv, guv, buv);
            if (c.vignette) {
                a = vignette_px(a, __ldg(c.vignette + (size_t)y0 * c.w + x));
                b = vignette_px(b, __ldg(c.vignette + (size_t)(y0 + 1) * c.w + x));
            }
        }
        d0[k] = a; d1[k] = b;
    }
}

// convert one item and store its 2 x 8 RGBX pixels into the stage (pixels outside the source become 0 = BORDER_CONSTANT)
__device__ __forceinline__ void convert_item(const CamSrc& c, const FJob& J, int item, const ItemRegs& r, uint32_t mask, uint32_t* s_rgbx)
{
    int rp, gx;
    item_pos(J, item, rp, gx);
    uint4* d0 = reinterpret_cast<uint4*>(s_rgbx + (rp << 1) * J.bw + (gx << 3));
    uint4* d1 = reinterpret_cast<uint4*>(s_rgbx + ((rp << 1) + 1) * J.bw + (gx << 3));
    if (mask == 0u) {
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        d0[0] = z; d0[1] = z; d1[0] = z; d1[1] = z;
        return;
    }
    if (mask & ITEM_SLOW) {
        convert_item_slow(c, J.bx0 + (gx << 3), J.by0 + (rp << 1), reinterpret_cast<uint32_t*>(d0), reinterpret_cast<uint32_t*>(d1));
        return;
    }
    uint32_t a[8], b[8];
    constexpr int K16 = 16 * 1220542;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        const int u = (int)((r.u >> (8 * k)) & 255u) - 128, v = (int)((r.v >> (8 * k)) & 255u) - 128;
        const int ruv = (1 << 19) - K16 + 1673527 * v;
        const int guv = (1 << 19) - K16 - 852492 * v - 409993 * u;
        const int buv = (1 << 19) - K16 + 2116026 * u;
        const uint32_t wa = k < 2 ? r.y0.x : r.y0.y, wb = k < 2 ? r.y1.x : r.y1.y;
        const int sh = 16 * (k & 1);
        a[2 * k] = yuv_px16(max((wa >> sh) & 255u, 16u), ruv, guv, buv);
        a[2 * k + 1] = yuv_px16(max((wa >> (sh + 8)) & 255u, 16u), ruv, guv, buv);
        b[2 * k] = yuv_px16(max((wb >> sh) & 255u, 16u), ruv, guv, buv);
        b[2 * k + 1] = yuv_px16(max((wb >> (sh + 8)) & 255u, 16u), ruv, guv, buv);
    }
    if (c.vignette) {
        const int x0 = J.bx0 + (gx << 3), y0 = J.by0 + (rp << 1);
        #pragma unroll
        for (int k = 0; k < 8; k++)
            {
                a[k] = vignette_px(a[k], __ldg(c.vignette + (size_t)y0 * c.w + x0 + k));
                b[k] = vignette_px(b[k], __ldg(c.vignette + (size_t)(y0 + 1) * c.w + x0 + k));
            }
    }
    d0[0] = make_uint4(a[0], a[1], a[2], a[3]); d0[1] = make_uint4(a[4], a[5], a[6], a[7]);
    d1[0] = make_uint4(b[0], b[1], b[2], b[3]); d1[1] = make_uint4(b[4], b[5], b[6], b[7]);
}

// one table entry: four taps from the stage, 1/32-px bilinear, gain, weight, accumulate.
// ex = byte offset of the top-left tap in the stage | fy << 16 | fx << 24 ; ew = f32 weight bits
template <int GAIN>
__device__ __forceinline__ void fused_pair(uint32_t ex, uint32_t ew, uint32_t stage, uint32_t bw4, float g32, float gbias,
                                           const uint8_t* __restrict__ lut, uint32_t& ar, uint32_t& ag, uint32_t& ab)
{
    const uint32_t a0 = stage + (ex & 0xFFFFu), a1 = a0 + bw4;
    const uint32_t t00 = lds_u32(a0), t01 = lds_u32(a0 + 4), t10 = lds_u32(a1), t11 = lds_u32(a1 + 4);
    const uint32_t fx = ex >> 24, fy = __byte_perm(ex, 0u, 0x4442);
    const uint32_t wx = fx * 65535u + 32u;                 // (32-fx) | fx << 16
    const uint32_t wb = wx * fy, wt = wx * 32u - wb;       // {(32-fx) fy, fx fy}, {(32-fx)(32-fy), fx (32-fy)} as 16-bit pairs
    const uint32_t rg0 = __byte_perm(t00, t01, 0x5140);    // R00 R01 G00 G01
    const uint32_t bb0 = __byte_perm(t00, t01, 0x6262);    // B00 B01 .. ..
    const uint32_t rg1 = __byte_perm(t10, t11, 0x5140);
    const uint32_t bb1 = __byte_perm(t10, t11, 0x6262);
    const uint32_t r = __dp2a_lo(wb, rg1, __dp2a_lo(wt, rg0, 512u));     // sum + 512 (< 2^18): exact in f32
    const uint32_t g = __dp2a_hi(wb, rg1, __dp2a_hi(wt, rg0, 512u));
    const uint32_t b = __dp2a_lo(wb, bb1, __dp2a_lo(wt, bb0, 512u));
    // floor(x / 1024) + 2^23 by a round-down fma: the integer lands in the mantissa (no shift, no F2I)
    float rf = __fmaf_rd(__uint2float_rn(r), 0.0009765625f, MAGIC_RD);
    float gf = __fmaf_rd(__uint2float_rn(g), 0.0009765625f, MAGIC_RD);
    float bf = __fmaf_rd(__uint2float_rn(b), 0.0009765625f, MAGIC_RD);
    if (GAIN) {
        if (lut == nullptr) {
            rf = gain_apply_biased(rf, g32, gbias); gf = gain_apply_biased(gf, g32, gbias); bf = gain_apply_biased(bf, g32, gbias);
        } else {
            rf = (float)__ldg(lut + (__float_as_uint(rf) & 255u)); gf = (float)__ldg(lut + (__float_as_uint(gf) & 255u));
            bf = (float)__ldg(lut + (__float_as_uint(bf) & 255u));
        }
    } else {
        rf = __fadd_rn(rf, -MAGIC_RD); gf = __fadd_rn(gf, -MAGIC_RD); bf = __fadd_rn(bf, -MAGIC_RD);
    }
    // (short)(v * W): f32 product, truncated (v*W >= 0 so floor == trunc); 2^23 bias removed in the same add.
    const float w = __uint_as_float(ew);
    ar += __float_as_uint(__fadd_rd(__fmul_rn(rf, w), MAGIC_RD)) - 0x4B000000u;
    ag += __float_as_uint(__fadd_rd(__fmul_rn(gf, w), MAGIC_RD)) - 0x4B000000u;
    ab += __float_as_uint(__fadd_rd(__fmul_rn(bf, w), MAGIC_RD)) - 0x4B000000u;
}

template <int GAIN>
__global__ void __launch_bounds__(FT_THREADS, 3) k_stitch_fused(const __grid_constant__ FusedParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* s_rgbx = reinterpret_cast<uint32_t*>(smem + FS_RGBX);
    uint8_t* s_y = smem + FS_Y; uint8_t* s_u = smem + FS_U; uint8_t* s_v = smem + FS_V;
    float4* s_gain = reinterpret_cast<float4*>(smem + FS_GAIN);
    const uint32_t stage = smem_u32(s_rgbx), ent_base = smem_u32(smem + FS_ENT), mbar_base = smem_u32(smem + FS_MBAR);
    const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;

    int ti = __ldg(p.bin_start + blockIdx.x);
    const int tend = __ldg(p.bin_start + blockIdx.x + 1);
    if (ti >= tend) return;

    // ---- prologue: metadata of the first two tiles, gain constants, barriers ----
    auto meta_slot = [&](int seq) { return smem + FS_META + (seq % 3) * FS_META_STRIDE; };
    if (tid < 65) {
        reinterpret_cast<uint4*>(meta_slot(ti))[tid] = __ldg(reinterpret_cast<const uint4*>(p.blocks + ti) + tid);
        if (ti + 1 < tend) reinterpret_cast<uint4*>(meta_slot(ti + 1))[tid] = __ldg(reinterpret_cast<const uint4*>(p.blocks + ti + 1) + tid);
    }
    if (GAIN && tid < p.n) {
        const float g32 = __ldcg(p.gain_f32 + tid);
        s_gain[tid] = make_float4(g32, gain_bias_f32(g32), __int_as_float(__ldcg(p.gain_flag + tid)), 0.f);
    }
    if (tid == 0) {
        f_mbar_init(mbar_base, 1); f_mbar_init(mbar_base + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    ItemRegs it0, it1;
    uint32_t imask = 0u;                                    // bits 0-15 item 0, bits 16-31 item 1
    auto prefetch_job = [&](const FJob& N, int j, int buf) {
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            f_mbar_expect_tx(mbar_base + 8 * buf, FT_PX * 8);
            bulk_g2s(ent_base + buf * (FT_PX * 8), p.entries + (size_t)j * FT_PX, FT_PX * 8, mbar_base + 8 * buf);
        }
        const CamSrc& c = p.cam[N.cam];
        uint32_t m0 = 0u, m1 = 0u;
        if (tid < N.nitems) load_item(c, N, tid, it0, m0);
        if (tid + FT_THREADS < N.nitems) load_item(c, N, tid + FT_THREADS, it1, m1);
        imask = m0 | (m1 << 16);
    };
    {
        const FTileBlock* B = reinterpret_cast<const FTileBlock*>(meta_slot(ti));
        prefetch_job(B->job[0], B->tile.j0, 0);
    }

    uint32_t phase0 = 0u, phase1 = 0u;
    int buf = 0;
    for (; ti < tend; ti++) {
        const FTileBlock* B = reinterpret_cast<const FTileBlock*>(meta_slot(ti));
        const FTile T = B->tile;
        uint32_t acc[FT_PPT][3];
        #pragma unroll
        for (int q = 0; q < FT_PPT; q++) acc[q][0] = acc[q][1] = acc[q][2] = 0u;
        uint4 meta_next = make_uint4(0u, 0u, 0u, 0u);

        #pragma unroll 1
        for (int k = 0; k < T.nj; k++) {
            const FJob J = B->job[k];
            // ---- (1) convert this job's source box (input bytes were prefetched during the previous job) ----
            {
                const CamSrc& c = p.cam[J.cam];
                if (tid < J.nitems) convert_item(c, J, tid, it0, imask & 0xFFFFu, s_rgbx);
                if (tid + FT_THREADS < J.nitems) convert_item(c, J, tid + FT_THREADS, it1, imask >> 16, s_rgbx);
            }
            __syncthreads();
            // ---- (2) prefetch the next job (entries by TMA bulk copy, input bytes into registers) ----
            if (k + 1 < T.nj) prefetch_job(B->job[k + 1], T.j0 + k + 1, buf ^ 1);
            else if (ti + 1 < tend) {
                const FTileBlock* Bn = reinterpret_cast<const FTileBlock*>(meta_slot(ti + 1));
                prefetch_job(Bn->job[0], Bn->tile.j0, buf ^ 1);
            }
            if (k == 0 && ti + 2 < tend && tid < 65) meta_next = __ldg(reinterpret_cast<const uint4*>(p.blocks + ti + 2) + tid);
            // ---- (3) gather ----
            if (buf == 0) { f_mbar_wait(mbar_base, phase0); phase0 ^= 1u; } else { f_mbar_wait(mbar_base + 8, phase1); phase1 ^= 1u; }
            {
                const uint32_t eaddr = ent_base + buf * (FT_PX * 8) + tid * 32;
                uint4 e0, e1;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e0.x), "=r"(e0.y), "=r"(e0.z), "=r"(e0.w) : "r"(eaddr));
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e1.x), "=r"(e1.y), "=r"(e1.z), "=r"(e1.w) : "r"(eaddr + 16));
                const uint32_t bw4 = (uint32_t)J.bw * 4u;
                float g32 = 0.f, gb = 0.f;
                const uint8_t* lut = nullptr;
                if (GAIN) {
                    const float4 gc = s_gain[J.cam];
                    g32 = gc.x; gb = gc.y;
                    if (__float_as_int(gc.z) != 0) lut = p.gain_lut + J.cam * 256;
                }
                fused_pair<GAIN>(e0.x, e0.y, stage, bw4, g32, gb, lut, acc[0][0], acc[0][1], acc[0][2]);
                fused_pair<GAIN>(e0.z, e0.w, stage, bw4, g32, gb, lut, acc[1][0], acc[1][1], acc[1][2]);
                fused_pair<GAIN>(e1.x, e1.y, stage, bw4, g32, gb, lut, acc[2][0], acc[2][1], acc[2][2]);
                fused_pair<GAIN>(e1.z, e1.w, stage, bw4, g32, gb, lut, acc[3][0], acc[3][1], acc[3][2]);
            }
            if (k == 0 && ti + 2 < tend && tid < 65) reinterpret_cast<uint4*>(meta_slot(ti + 2))[tid] = meta_next;
            buf ^= 1;
            __syncthreads();                                // stage and entry buffer are free again
        }

        // ---- tile epilogue: normalise, RGB -> YUV 4:2:0 into shared memory, then 128-bit row stores ----
        const int tx0 = T.tx * FT_W, ty0 = T.ty * FT_H;
        #pragma unroll
        for (int q = 0; q < FT_PPT; q++) {
            const int row = ly + 8 * q;
            const uint32_t px = normalise_px(acc[q][0], acc[q][1], acc[q][2], p.inv_n);
            const int R = px & 255u, G = (px >> 8) & 255u, Bc = (px >> 16) & 255u;
            s_y[row * FT_W + lx] = (uint8_t)rgb_luma(R, G, Bc);
            if (((lx | ly) & 1) == 0) {                      // top-left pixel of a 2x2 block carries the chroma (color.cpp:6456-6481)
                s_u[(row >> 1) * (FT_W / 2) + (lx >> 1)] = (uint8_t)rgb_cb(R, G, Bc);
                s_v[(row >> 1) * (FT_W / 2) + (lx >> 1)] = (uint8_t)rgb_cr(R, G, Bc);
            }
            if (p.rgb_out) {
                const int x = tx0 + lx, y = ty0 + row;
                if (x < p.out_w && y < p.out_h) {
                    uint8_t* o = p.rgb_out + (size_t)y * p.rgb_pitch + 3 * x;
                    o[0] = (uint8_t)R; o[1] = (uint8_t)G; o[2] = (uint8_t)Bc;
                }
            }
        }
        __syncthreads();
        if (p.oy) {
            if (tid < 64) {                                 // luma: 32 rows x 2 halves of 16 px
                const int row = tid >> 1, hx = (tid & 1) << 4;
                const int x = tx0 + hx, y = ty0 + row;
                if (y < p.out_h && x < p.out_w) {
                    const uint4 v = *reinterpret_cast<const uint4*>(s_y + row * FT_W + hx);
                    uint8_t* o = p.oy + (size_t)y * p.oy_pitch + x;
                    if (x + 16 <= p.out_w && (((uintptr_t)o) & 15) == 0) *reinterpret_cast<uint4*>(o) = v;
                    else { const uint8_t* sv = s_y + row * FT_W + hx; for (int k = 0; k < 16 && x + k < p.out_w; k++) o[k] = sv[k]; }
                }
            } else if (tid < 96) {                          // chroma: 16 rows x 16 samples per plane
                const int t = tid - 64;
                if (p.uv_step == 1) {
                    const int row = t & 15;
                    const uint8_t* sv = (t < 16 ? s_u : s_v) + row * (FT_W / 2);
                    const int x = tx0 >> 1, y = (ty0 >> 1) + row;
                    if (y < (p.out_h >> 1) && x < (p.out_w >> 1)) {
                        uint8_t* o = (t < 16 ? p.ou + (size_t)y * p.ou_pitch : p.ov + (size_t)y * p.ov_pitch) + x;
                        if (x + 16 <= (p.out_w >> 1) && (((uintptr_t)o) & 15) == 0) *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(sv);
                        else for (int k = 0; k < 16 && x + k < (p.out_w >> 1); k++) o[k] = sv[k];
                    }
                } else {                                    // interleaved (NV12-style) chroma: 8 samples of each plane per thread
                    const int row = t >> 1, hx = (t & 1) << 3;
                    const int x = (tx0 >> 1) + hx, y = (ty0 >> 1) + row;
                    if (y < (p.out_h >> 1))
                        for (int k = 0; k < 8 && x + k < (p.out_w >> 1); k++) {
                            p.ou[(size_t)y * p.ou_pitch + (size_t)(x + k) * p.uv_step] = s_u[row * (FT_W / 2) + hx + k];
                            p.ov[(size_t)y * p.ov_pitch + (size_t)(x + k) * p.uv_step] = s_v[row * (FT_W / 2) + hx + k];
                        }
                }
            }
        }
        // the next epilogue writes s_y/s_u/s_v only after at least two more CTA barriers (one job), so no barrier here
    }
}

int fused_ctas_per_sm()
{
    int nb = 0;
    cudaFuncSetAttribute(k_stitch_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_TOTAL);
    cudaFuncSetAttribute(k_stitch_fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS_TOTAL);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_stitch_fused<1>, FT_THREADS, FS_TOTAL) != cudaSuccess) nb = 0;
    return nb;
}

void launch_stitch_fused(const FusedParams& p, int grid, cudaStream_t s)
{
    if (p.use_gain) k_stitch_fused<1><<<grid, FT_THREADS, FS_TOTAL, s>>>(p);
    else k_stitch_fused<0><<<grid, FT_THREADS, FS_TOTAL, s>>>(p);
}

}  // namespace ob
