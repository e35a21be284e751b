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
constexpr int FS_RGBX = 0;                                   // FUSED_CAP u32: the converted source box of the current job
constexpr int FS_SLOT = FS_RGBX + FUSED_CAP * 4;             // thread-private 16-byte slots for the input bytes of the NEXT job's items
                                                             //   (cp.async destinations; slot s of thread t at [s][t]): luma row 0,
                                                             //   luma row 1 (4 px each), U chunk, V chunk (or NV12 UVUV, unused)
constexpr int FS_ENT = FS_SLOT + 2 * FT_THREADS * 16;        // 2 x FT_PX x 8 B table entries (TMA bulk destinations)
constexpr int FS_OUT = FS_ENT + 2 * FT_PX * 8;               // 2 x {32 x 32 luma, 16 x 16 U, 16 x 16 V} of finished tiles
constexpr int FS_OUT_STRIDE = FT_PX + FT_PX / 2;
constexpr int FS_GAIN = FS_OUT + 2 * FS_OUT_STRIDE;          // MAX_CAMS x {g32, bias, flag, pad}
constexpr int FS_MBAR = FS_GAIN + MAX_CAMS * 16;             // ent[2]
constexpr int FS_TOTAL = FS_MBAR + 16;
static_assert(FS_SLOT % 16 == 0 && FS_ENT % 128 == 0 && FS_OUT % 16 == 0 && FS_OUT_STRIDE % 16 == 0 && FS_MBAR % 8 == 0, "alignment");
static_assert(4 * (FS_TOTAL + 1024) <= 228 * 1024, "four CTAs per SM");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
// per-thread asynchronous copies (SASS: LDGSTS): the input bytes of the next job go global -> shared without
// passing through registers
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- conversion items: 8 px x 2 rows of the job's source box (four chroma samples), listed by the host ----
// what a thread keeps of a job record
struct JobRegs { int cam, bx0, by0, bw, nitems, stage_off; uint32_t tile_xy; };
// issue the asynchronous copies of one item's input bytes (class FAST on 4-byte aligned planes) into the thread's slot.
// Planar chroma is copied as the aligned 4-sample chunk that holds the item's two samples.
__device__ __forceinline__ void fetch_item(const CamSrc& c, const JobRegs& J, uint32_t desc, uint32_t slot)
{
    const int rp = desc & 127u, gx = (desc >> 7) & 127u;
    const int x0 = J.bx0 + (gx << 2), y0 = J.by0 + (rp << 1);
    const uint8_t* yp = c.y + (size_t)y0 * c.y_pitch + x0;
    cp_async4(slot, yp);
    cp_async4(slot + 4, yp + c.y_pitch);
    if (c.uv_step == 1) {
        const size_t co = (size_t)((x0 >> 1) & ~3);
        cp_async4(slot + 8, c.u + (size_t)(y0 >> 1) * c.u_pitch + co);
        cp_async4(slot + 12, c.v + (size_t)(y0 >> 1) * c.v_pitch + co);
    } else                                                   // NV12: U0 V0 U1 V1
        cp_async4(slot + 8, c.u + (size_t)(y0 >> 1) * c.u_pitch + x0);
}

// BT.601 limited-range integer conversion (imgproc/src/color.cpp:6087-6169):
//   R = sat((CY (max(Y,16) - 16) + 1673527 (V-128) + 2^19) >> 20), ...
// evaluated as the HIGH word of a 64-bit multiply-add: (Yc << 8)(CY << 4) + (chroma term << 12) = (sum) << 12, so
// bits 63..32 are sum >> 20 -- one IMAD.HI instead of IMAD + shift.  The byte operands arrive pre-shifted by 8 from
// the same PRMT that extracts them; the chroma terms are 64-bit IMAD.WIDE results shared by a 2x2 block.
constexpr int CY = 1220542, CK = (1 << 19) - 16 * CY;
__device__ __forceinline__ int mad_hi64(uint32_t a8, int b16, long long c) { return (int)(((long long)(int)a8 * b16 + c) >> 32); }
struct Chroma { long long r, g, b; };
__device__ __forceinline__ Chroma chroma_terms(uint32_t U8, uint32_t V8)   // U8 = U << 8, V8 = V << 8
{
    Chroma c;
    c.r = (long long)(int)V8 * (1673527 * 16) + (((long long)(CK - 128 * 1673527)) << 12);
    c.g = (long long)(int)V8 * (-852492 * 16) + ((long long)(int)U8 * (-409993 * 16) + (((long long)(CK + 128 * (852492 + 409993))) << 12));
    c.b = (long long)(int)U8 * (2116026 * 16) + (((long long)(CK - 128 * 2116026)) << 12);
    return c;
}
__device__ __forceinline__ uint32_t yuv_px(uint32_t Y8, const Chroma& c)    // Y8 = Y << 8 ; returns R | G << 8 | B << 16
{
    const uint32_t yc = max(Y8, 16u << 8);
    const uint32_t r = (uint32_t)__vimin_s32_relu(mad_hi64(yc, CY * 16, c.r), 255);
    const uint32_t g = (uint32_t)__vimin_s32_relu(mad_hi64(yc, CY * 16, c.g), 255);
    const uint32_t b = (uint32_t)__vimin_s32_relu(mad_hi64(yc, CY * 16, c.b), 255);
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

// border / unaligned item: byte loads from global memory, one pixel at a time (rare)
__device__ __noinline__ void convert_item_slow(const CamSrc& c, int x0, int y0, uint32_t* d0, uint32_t* d1)
{
    for (int k = 0; k < 4; k++) {
        const int x = x0 + k;
        uint32_t a = 0u, b = 0u;
        if (x >= 0 && x < c.w && y0 >= 0 && y0 < c.h) {
            const Chroma ch = chroma_terms((uint32_t)__ldg(c.u + (size_t)(y0 >> 1) * c.u_pitch + (size_t)(x >> 1) * c.uv_step) << 8,
                                           (uint32_t)__ldg(c.v + (size_t)(y0 >> 1) * c.v_pitch + (size_t)(x >> 1) * c.uv_step) << 8);
            a = yuv_px((uint32_t)__ldg(c.y + (size_t)y0 * c.y_pitch + x) << 8, ch);
            b = yuv_px((uint32_t)__ldg(c.y + (size_t)(y0 + 1) * c.y_pitch + x) << 8, ch);
            if (c.vignette) {
                a = vignette_px(a, __ldg(c.vignette + (size_t)y0 * c.w + x));
                b = vignette_px(b, __ldg(c.vignette + (size_t)(y0 + 1) * c.w + x));
            }
        }
        d0[k] = a; d1[k] = b;
    }
}

// convert one item (input bytes in the thread's slot) and store its 2 x 4 RGBX pixels into the RGBX stage
__device__ __forceinline__ void convert_item(const CamSrc& c, const JobRegs& J, uint32_t desc, bool fast, const uint8_t* slot, uint32_t* s_rgbx)
{
    const int rp = desc & 127u, gx = (desc >> 7) & 127u;
    const int o0 = J.stage_off + (rp << 1) * J.bw + (gx << 2);
    uint4* d0 = reinterpret_cast<uint4*>(s_rgbx + o0);
    uint4* d1 = reinterpret_cast<uint4*>(s_rgbx + o0 + J.bw);
    if (!fast) {
        if ((desc >> 14) == FITEM_ZERO) {                    // BORDER_CONSTANT
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            *d0 = z; *d1 = z;
        } else
            convert_item_slow(c, J.bx0 + (gx << 2), J.by0 + (rp << 1), reinterpret_cast<uint32_t*>(d0), reinterpret_cast<uint32_t*>(d1));
        return;
    }
    const uint4 in = *reinterpret_cast<const uint4*>(slot);  // luma row 0, luma row 1, chroma
    uint32_t ub, vb;                                         // the item's two U and two V samples in bytes 0, 1
    if (c.uv_step == 1) { const int sh = (gx & 1) << 4; ub = in.z >> sh; vb = in.w >> sh; }
    else { ub = __byte_perm(in.z, 0u, 0x4420); vb = __byte_perm(in.z, 0u, 0x4431); }
    const Chroma c0 = chroma_terms(__byte_perm(ub, 0u, 0x4404), __byte_perm(vb, 0u, 0x4404));
    const Chroma c1 = chroma_terms(__byte_perm(ub, 0u, 0x4414), __byte_perm(vb, 0u, 0x4414));
    uint4 a, b;
    a.x = yuv_px(__byte_perm(in.x, 0u, 0x4404), c0); a.y = yuv_px(__byte_perm(in.x, 0u, 0x4414), c0);
    a.z = yuv_px(__byte_perm(in.x, 0u, 0x4424), c1); a.w = yuv_px(__byte_perm(in.x, 0u, 0x4434), c1);
    b.x = yuv_px(__byte_perm(in.y, 0u, 0x4404), c0); b.y = yuv_px(__byte_perm(in.y, 0u, 0x4414), c0);
    b.z = yuv_px(__byte_perm(in.y, 0u, 0x4424), c1); b.w = yuv_px(__byte_perm(in.y, 0u, 0x4434), c1);
    if (c.vignette) {
        const float* vg = c.vignette + (size_t)(J.by0 + (rp << 1)) * c.w + J.bx0 + (gx << 2);
        a.x = vignette_px(a.x, __ldg(vg)); a.y = vignette_px(a.y, __ldg(vg + 1));
        a.z = vignette_px(a.z, __ldg(vg + 2)); a.w = vignette_px(a.w, __ldg(vg + 3));
        b.x = vignette_px(b.x, __ldg(vg + c.w)); b.y = vignette_px(b.y, __ldg(vg + c.w + 1));
        b.z = vignette_px(b.z, __ldg(vg + c.w + 2)); b.w = vignette_px(b.w, __ldg(vg + c.w + 3));
    }
    *d0 = a; *d1 = b;
}

template <int GAIN, bool LUT>
__device__ __forceinline__ void gather_job(const uint4 e0, const uint4 e1, const uint8_t* s0, const uint8_t* s1, float g32, float gb,
                                           const uint8_t* lut, uint32_t (&arg)[FT_PPT], uint32_t (&ab)[FT_PPT / 2])
{
    fused_pair<GAIN, LUT, false>(e0.x, e0.y, s0, s1, g32, gb, lut, arg[0], ab[0]);
    fused_pair<GAIN, LUT, true>(e0.z, e0.w, s0, s1, g32, gb, lut, arg[1], ab[0]);
    fused_pair<GAIN, LUT, false>(e1.x, e1.y, s0, s1, g32, gb, lut, arg[2], ab[1]);
    fused_pair<GAIN, LUT, true>(e1.z, e1.w, s0, s1, g32, gb, lut, arg[3], ab[1]);
}

// dst_16s.convertTo(CV_8UC3, 1.0/N): sat_u8(rint((float)acc * (float)(1/N))) without F2I: the product is rounded to
// f32 first (as the reference does), the magic add rounds it to the nearest-even integer in the mantissa
__device__ __forceinline__ int normalise_ch(uint32_t acc, float inv_n)
{
    const uint32_t bits = __float_as_uint(__fadd_rn(__fmul_rn(__uint2float_rn(acc), inv_n), MAGIC_RN));
    return (int)(min(bits, 0x4B4000FFu) & 255u);
}

template <int GAIN>
__global__ void __launch_bounds__(FT_THREADS, 4) k_stitch_fused(const __grid_constant__ FusedParams p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* s_rgbx = reinterpret_cast<uint32_t*>(smem + FS_RGBX);
    float4* s_gain = reinterpret_cast<float4*>(smem + FS_GAIN);
    const uint32_t ent_base = smem_u32(smem + FS_ENT), mbar = smem_u32(smem + FS_MBAR);
    const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
    const uint8_t* slot0 = smem + FS_SLOT + tid * 16;         // the thread's two item slots
    const uint8_t* slot1 = slot0 + FT_THREADS * 16;
    const uint32_t slot0_a = smem_u32(slot0), slot1_a = slot0_a + FT_THREADS * 16;

    const FBin bin = p.bins[blockIdx.x];
    const int js = bin.start, je = bin.end;
    if (js >= je) return;
    if (GAIN && tid < p.n) {
        const float g32 = __ldcg(p.gain_f32 + tid);
        s_gain[tid] = make_float4(g32, gain_bias_f32(g32), __int_as_float(__ldcg(p.gain_flag + tid)), 0.f);
    }
    if (tid == 0) {
        f_mbar_init(mbar, 1); f_mbar_init(mbar + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const CamSrc* cams = p.cam;
    const uint16_t* items = p.items + tid;                  // job j's descriptors at items[j * FUSED_MAXITEMS + slot * FT_THREADS + thread]
    auto load_job = [&](int j) {
        const uint4* q = reinterpret_cast<const uint4*>(p.jobs + j);
        const uint4 a = __ldg(q), b = __ldg(q + 1);
        JobRegs J;
        J.cam = (int)a.x; J.bx0 = (int)a.y; J.by0 = (int)a.z; J.bw = (int)a.w; J.nitems = (int)b.x; J.tile_xy = b.z; J.stage_off = (int)b.w;
        return J;
    };
    auto is_fast = [&](const JobRegs& Q, uint32_t d, int item) {
        return item < Q.nitems && (d >> 14) == FITEM_FAST && cams[Q.cam & 0xFF].aligned4;
    };
    // input bytes of the thread's items of job Q -> its slots (asynchronous; consumed after the next job boundary)
    auto fetch_job = [&](const JobRegs& Q, uint32_t da, uint32_t db) {
        const CamSrc& c = cams[Q.cam & 0xFF];
        if (is_fast(Q, da, tid)) fetch_item(c, Q, da, slot0_a);
        if (is_fast(Q, db, tid + FT_THREADS)) fetch_item(c, Q, db, slot1_a);
        cp_async_commit();
    };
    auto request_entries = [&](int j, int buf) {             // one thread: TMA bulk copy of a job's 8 KB of table entries
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        f_mbar_expect_tx(mbar + 8 * buf, FT_PX * 8);
        bulk_g2s(ent_base + buf * (FT_PX * 8), p.entries + (size_t)j * FT_PX, FT_PX * 8, mbar + 8 * buf);
    };

    // ---- prologue: job js fully fetched, job js + 1's record and descriptors in registers ----
    JobRegs J = load_job(js), N = J;
    uint32_t d0 = __ldg(items + (size_t)js * FUSED_MAXITEMS), d1 = __ldg(items + (size_t)js * FUSED_MAXITEMS + FT_THREADS);
    uint32_t e0 = 0u, e1 = 0u;
    if (tid == 0) request_entries(js, 0);
    fetch_job(J, d0, d1);
    if (js + 1 < je) {
        N = load_job(js + 1);
        e0 = __ldg(items + (size_t)(js + 1) * FUSED_MAXITEMS); e1 = __ldg(items + (size_t)(js + 1) * FUSED_MAXITEMS + FT_THREADS);
    }

    uint32_t arg[FT_PPT], ab[FT_PPT / 2];
    #pragma unroll
    for (int q = 0; q < FT_PPT; q++) arg[q] = 0u;
    #pragma unroll
    for (int q = 0; q < FT_PPT / 2; q++) ab[q] = 0u;
    uint32_t phase = 0u, tiles_done = 0u, pending_xy = 0xFFFFFFFFu;
    int buf = 0;
    // 128-bit row stores of a finished tile from its shared-memory buffer (threads 0..95)
    auto store_tile = [&](uint32_t tile_xy, uint32_t which) {
        const uint8_t* s_y = smem + FS_OUT + (which & 1u) * FS_OUT_STRIDE;
        const uint8_t* s_u = s_y + FT_PX; const uint8_t* s_v = s_u + FT_PX / 4;
        const int tx0 = (int)(tile_xy & 0xFFFFu) * FT_W, ty0 = (int)(tile_xy >> 16) * FT_H;
        if (tid < 64) {                                     // luma: 32 rows x 2 halves of 16 px
            const int row = tid >> 1, hx = (tid & 1) << 4;
            const int x = tx0 + hx, y = ty0 + row;
            if (y < p.out_h && x < p.out_w) {
                const uint4 v = *reinterpret_cast<const uint4*>(s_y + row * FT_W + hx);
                uint8_t* o = p.oy + (size_t)y * p.oy_pitch + x;
                if (x + 16 <= p.out_w && (((uintptr_t)o) & 15) == 0) *reinterpret_cast<uint4*>(o) = v;
                else { const uint8_t* sv = s_y + row * FT_W + hx; for (int k = 0; k < 16 && x + k < p.out_w; k++) o[k] = sv[k]; }
            }
        } else if (tid < 96) {                              // chroma: 16 rows x 16 samples per plane
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
            } else {                                        // interleaved (NV12-style) chroma: 8 samples of each plane per thread
                const int row = t >> 1, hx = (t & 1) << 3;
                const int x = (tx0 >> 1) + hx, y = (ty0 >> 1) + row;
                if (y < (p.out_h >> 1))
                    for (int k = 0; k < 8 && x + k < (p.out_w >> 1); k++) {
                        p.ou[(size_t)y * p.ou_pitch + (size_t)(x + k) * p.uv_step] = s_u[row * (FT_W / 2) + hx + k];
                        p.ov[(size_t)y * p.ov_pitch + (size_t)(x + k) * p.uv_step] = s_v[row * (FT_W / 2) + hx + k];
                    }
            }
        }
    };
    #pragma unroll 1
    for (int j = js; j < je; j++) {
        // ---- (1) convert this job's items: the thread's own bytes have landed once its copy group completes ----
        cp_async_wait_all();
        {
            const CamSrc& c = cams[J.cam & 0xFF];
            if (tid < J.nitems) convert_item(c, J, d0, is_fast(J, d0, tid), slot0, s_rgbx);
            if (tid + FT_THREADS < J.nitems) convert_item(c, J, d1, is_fast(J, d1, tid + FT_THREADS), slot1, s_rgbx);
        }
        __syncthreads();       // the ONE barrier of a job: conversion done; every thread has also left the previous gather / epilogue
        if (pending_xy != 0xFFFFFFFFu) {                    // the tile finished by the previous job: its bytes are in place now
            if (p.oy) store_tile(pending_xy, tiles_done - 1u);
            pending_xy = 0xFFFFFFFFu;
        }
        // ---- (2) put the next jobs in flight: entries of j + 1, input bytes of j + 1, record and descriptors of j + 2 ----
        if (j + 1 < je) {
            if (tid == 0) request_entries(j + 1, buf ^ 1);
            fetch_job(N, e0, e1);
        }
        JobRegs NN = N;
        uint32_t g0 = 0u, g1 = 0u;
        if (j + 2 < je) {
            NN = load_job(j + 2);
            g0 = __ldg(items + (size_t)(j + 2) * FUSED_MAXITEMS); g1 = __ldg(items + (size_t)(j + 2) * FUSED_MAXITEMS + FT_THREADS);
        }
        // ---- (3) gather ----
        f_mbar_wait(mbar + 8 * buf, (phase >> buf) & 1u);
        phase ^= 1u << buf;
        {
            const uint4* ep = reinterpret_cast<const uint4*>(smem + FS_ENT + buf * (FT_PX * 8)) + tid * 2;
            const uint4 en0 = ep[0], en1 = ep[1];
            const uint8_t* s0 = smem + FS_RGBX;
            const uint8_t* s1 = s0 + J.bw * 4;
            const int cam = J.cam & 0xFF;
            if (GAIN) {
                const float4 gc = s_gain[cam];
                if (__float_as_int(gc.z) == 0) gather_job<GAIN, false>(en0, en1, s0, s1, gc.x, gc.y, nullptr, arg, ab);
                else gather_job<GAIN, true>(en0, en1, s0, s1, gc.x, gc.y, p.gain_lut + cam * 256, arg, ab);
            } else
                gather_job<0, false>(en0, en1, s0, s1, 0.f, 0.f, nullptr, arg, ab);
        }
        buf ^= 1;
        if (J.cam < 0) {
            // ---- tile epilogue: normalise, RGB -> YUV 4:2:0 into the tile's shared-memory buffer (stored after the next barrier) ----
            uint8_t* s_y = smem + FS_OUT + (tiles_done & 1u) * FS_OUT_STRIDE;
            uint8_t* s_u = s_y + FT_PX; uint8_t* s_v = s_u + FT_PX / 4;
            tiles_done++;
            pending_xy = J.tile_xy;
            const int tx0 = (int)(J.tile_xy & 0xFFFFu) * FT_W, ty0 = (int)(J.tile_xy >> 16) * FT_H;
            #pragma unroll
            for (int q = 0; q < FT_PPT; q++) {
                const int row = ly + 8 * q;
                const int R = normalise_ch(arg[q] & 0xFFFFu, p.inv_n), G = normalise_ch(arg[q] >> 16, p.inv_n);
                const int Bc = normalise_ch((q & 1) ? ab[q >> 1] >> 16 : ab[q >> 1] & 0xFFFFu, p.inv_n);
                s_y[row * FT_W + lx] = (uint8_t)rgb_luma(R, G, Bc);
                if (((lx | ly) & 1) == 0) {                  // top-left pixel of a 2x2 block carries the chroma (color.cpp:6456-6481)
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
            #pragma unroll
            for (int q = 0; q < FT_PPT; q++) arg[q] = 0u;
            #pragma unroll
            for (int q = 0; q < FT_PPT / 2; q++) ab[q] = 0u;
        }
        J = N; d0 = e0; d1 = e1;
        N = NN; e0 = g0; e1 = g1;
        // consecutive boxes normally sit side by side in the stage, so the next conversion may start while other warps
        // still gather; only when the host could not place them apart (FJOB_SYNC) a second barrier is needed
        if (j + 1 < je && (J.cam & FJOB_SYNC)) __syncthreads();
    }
    __syncthreads();
    if (pending_xy != 0xFFFFFFFFu && p.oy) store_tile(pending_xy, tiles_done - 1u);
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
