// csrc/blend_ring.cu -- K_blend_ring: the feather / no-blend frame kernel of the default path (sm_100a).
//
// Same arithmetic contracts as kernels.cu (bit-exact against the reference's CPU functions, see oracle/):
//   bilinear : imgproc/src/imgwarp.cpp:4383-4442 + :3812-4020 -> (sum S_k a_k b_k + 512) >> 10
//   gain     : core/src/arithm.cpp multiply-by-scalar in f64 -> sat_u8(rint(v*g))
//   feather  : stitching/src/cuda/blender.cu:73-98 (short)(v*W) truncation, blenders.cpp:581 (x 1/N, rint)
//   colour   : imgproc/src/color.cpp:6430-6481 (RGB -> YUV 4:2:0)
//
// Structure (what K_blend_staged's ncu profile asked for: its top stall was the CTA barrier behind a single-buffered
// TMA stage, and 63 % of its warp-instructions were per-tile overhead, not the gather):
//   * persistent CTAs (grid = SMs x resident CTAs) fetch 32x16 output tiles from a device-side ticket counter;
//   * a CTA is 4 consumer warps + 1 producer warp.  The producer walks the tile's jobs (one per covering camera),
//     allocates [entries | source box] in a byte ring of shared memory and issues two TMA operations per job: a bulk
//     copy of the job's 4 KB of table entries (UBLKCP) and a 2-D tensor copy of the job's source box out of the
//     camera's RGBX plane (UTMALDG, zero fill outside the image = BORDER_CONSTANT).  Up to RG_NST jobs are in flight;
//     one "full" mbarrier per job carries the byte count, one "empty" mbarrier per job takes one arrival per consumer
//     warp.  There is no CTA-wide barrier after start-up: warps drift freely inside the ring;
//   * a consumer thread owns 4 pixels of the tile (one column, rows 2w, 2w+1, 2w+8, 2w+9 for warp w), so per-job and
//     per-tile overhead is amortised over 4 table entries, the accumulators stay in registers across the jobs of a
//     tile, and every warp owns whole 4:2:0 chroma rows: the epilogue transposes through a warp-private 192-byte patch
//     (no CTA barrier) and leaves with 128-bit stores (luma: one full 32-byte sector per tile row);
//   * the FP32 part of an entry pair runs on packed FFMA2 / FMUL2 / FADD2 (sm_100 f32x2, with directed rounding);
//     the bilinear sum is produced directly as the mantissa of a float (IDP.2A accumulating onto bits(2^23)): no I2F.
#include "kernels.cuh"
#include <cstdlib>
#include "device_common.cuh"

namespace ob {

constexpr int RG_CONS_WARPS = 4;
constexpr int RG_CONS = RG_CONS_WARPS * 32;           // consumer threads = 128 = TILE_PX / 4
constexpr int RG_THREADS = RG_CONS + 32;              // + producer warp
constexpr int RG_MIN_CTAS = (227 * 1024) / (RING_BYTES + 2048) < 7 ? (227 * 1024) / (RING_BYTES + 2048) : 7;
constexpr int RG_NST = 8;                             // job descriptors in flight
constexpr uint32_t RG_EXIT = 1u, RG_EMPTY = 2u, RG_LAST = 4u, RG_CLAMP = 8u;
// shared-memory map (dynamic, 1024-byte aligned base)
constexpr int RS_RING = 0;
constexpr int RS_META = RS_RING + RING_BYTES;          // RG_NST x 32 B {box address, pitch | flags << 16 | cam << 24, gain bits, tile x | y << 16; next job}
constexpr int RS_FULL = RS_META + RG_NST * 32;         // RG_NST mbarriers
constexpr int RS_EMPTY = RS_FULL + RG_NST * 8;
constexpr int RS_OFF = RS_EMPTY + RG_NST * 8;          // RG_NST x u32: ring offset of each in-flight job (producer only)
constexpr int RS_PATCH = RS_OFF + RG_NST * 4;          // per consumer warp: 4 x 32 luma, 2 x 16 U, 2 x 16 V
constexpr int RS_TISSUE = RS_PATCH + RG_CONS_WARPS * 192;   // RG_NST x u64 issue time stamps (diagnostics)
constexpr int RS_TOTAL = RS_TISSUE + RG_NST * 8;
static_assert(RS_PATCH % 16 == 0 && RS_META % 16 == 0 && RS_FULL % 8 == 0, "alignment");
static_assert(TILE_PX == 4 * RG_CONS && TILE_W == 32 && TILE_H == 16, "thread -> pixel map");

__device__ __forceinline__ uint32_t r_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a)
{
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void r_mbar_init(uint32_t mbar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory"); }
__device__ __forceinline__ void r_mbar_expect_tx(uint32_t mbar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void r_mbar_arrive(uint32_t mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); }
// try_wait suspends the warp in hardware up to the hint; the loop only re-issues after a time-out
__device__ __forceinline__ void r_mbar_wait(uint32_t mbar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "RWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra RDONE_%=;\n\t"
        "nanosleep.u32 400;\n\t"
        "bra RWAIT_%=;\n\t"
        "RDONE_%=:\n\t}" ::"r"(mbar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void r_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "l"(pol) : "memory");
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* ptr, uint64_t pol)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ void r_tma_2d(uint32_t dst, const void* tmap, int x, int y, uint32_t mbar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)tmap), "r"(mbar), "r"(x), "r"(y) : "memory");
}

// ---- packed FP32 (sm_100 f32x2: SASS FFMA2 / FMUL2 / FADD2, rounding modes included) ----
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_pack(uint32_t lo, uint32_t hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
__device__ __forceinline__ f2 f2_dup(float v) { f2 r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v)); return r; }
__device__ __forceinline__ void f2_unpack(f2 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f2 f2_fma_rn(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 f2_fma_rm(f2 a, f2 b, f2 c) { f2 r; asm("fma.rm.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 f2_mul_rn(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_add_rn(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_add_rm(f2 a, f2 b) { f2 r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 f2_min_each(f2 a, float m)
{
    uint32_t lo, hi; f2_unpack(a, lo, hi);
    return f2_pack(__float_as_uint(fminf(__uint_as_float(lo), m)), __float_as_uint(fminf(__uint_as_float(hi), m)));
}

constexpr float TWO23 = 8388608.f;
constexpr uint32_t TWO23_BITS = 0x4B000000u;

// Integer half of one table entry: four taps from the ring, 1/32-px bilinear with IDP.2A accumulating onto
// bits(2^23) + 512, so r / g / b come out as the BIT PATTERNS of the floats 2^23 + (sum + 512)  (sum + 512 < 2^18).
// ex = byte offset of the top-left tap in the job's box << 16 | fy << 8 | fx.
__device__ __forceinline__ void ring_bilin(uint32_t ex, uint32_t box, uint32_t pitch, uint32_t& r, uint32_t& g, uint32_t& b)
{
    const uint32_t a0 = box + (ex >> 16), a1 = a0 + pitch;
    const uint32_t t00 = lds32(a0), t01 = lds32(a0 + 4), t10 = lds32(a1), t11 = lds32(a1 + 4);
    const uint32_t fx = ex & 0xFFu, fy = __byte_perm(ex, 0u, 0x4441);
    const uint32_t wx = fx * 65535u + 32u;                 // (32-fx) | fx << 16
    const uint32_t wb = wx * fy, wt = wx * 32u - wb;       // {(32-fx) fy, fx fy}, {(32-fx)(32-fy), fx (32-fy)} as 16-bit pairs
    const uint32_t rg0 = __byte_perm(t00, t01, 0x5140);    // R00 R01 G00 G01
    const uint32_t bb0 = __byte_perm(t00, t01, 0x6262);    // B00 B01 .. ..
    const uint32_t rg1 = __byte_perm(t10, t11, 0x5140);
    const uint32_t bb1 = __byte_perm(t10, t11, 0x6262);
    const uint32_t c = TWO23_BITS + 512u;
    r = __dp2a_lo(wb, rg1, __dp2a_lo(wt, rg0, c));
    g = __dp2a_hi(wb, rg1, __dp2a_hi(wt, rg0, c));
    b = __dp2a_lo(wb, bb1, __dp2a_lo(wt, bb0, c));
}

// FP half of TWO entries (a, b) of one channel.  in: bits of 2^23 + (sum + 512).  Steps, each exact in the sense the
// contract needs:
//   t = fma.rm(in, 2^-10, 2^23 - 2^13)        = 2^23 + floor((sum + 512) / 1024)          (the remapped 8-bit value v)
//   y = fma.rn(t, g, 2^23 - 2^23 g)           = 2^23 + rint(v g)                          (gain_tables() verifies g per camera)
//   y = min(y, 2^23 + 255)                                                                (saturate_cast<uchar>; only if g can exceed)
//   p = fma.rn(y, w, -2^23 w)                 = fl(u w)     (-2^23 w is exact, so the fma rounds the exact product u w once)
//   k = add.rm(p, 2^23)                       = 2^23 + floor(fl(u w))                     ((short)(v * W), blender.cu:82-97)
#ifndef RING_DEBUG
#define RING_DEBUG 0
#endif
#ifndef RING_SCALAR_FP
#define RING_SCALAR_FP 0
#endif
template <int MODE>   // 0: no gain, 1: gain without clamp, 2: gain with clamp
__device__ __forceinline__ f2 ring_chan(f2 in, f2 g2, f2 gb2, f2 w2, f2 c2)
{
#if RING_SCALAR_FP
    uint32_t i0, i1, g0, g1, b0, b1, w0, w1, c0, c1;
    f2_unpack(in, i0, i1); f2_unpack(g2, g0, g1); f2_unpack(gb2, b0, b1); f2_unpack(w2, w0, w1); f2_unpack(c2, c0, c1);
    auto one = [&](uint32_t i, uint32_t g, uint32_t b, uint32_t w, uint32_t c) {
        float t = __fmaf_rd(__uint_as_float(i), 0.0009765625f, 8380416.f);
        if (MODE >= 1) t = __fmaf_rn(t, __uint_as_float(g), __uint_as_float(b));
        if (MODE == 2) t = fminf(t, TWO23 + 255.f);
        return __float_as_uint(__fadd_rd(__fmaf_rn(t, __uint_as_float(w), __uint_as_float(c)), TWO23));
    };
    return f2_pack(one(i0, g0, b0, w0, c0), one(i1, g1, b1, w1, c1));
#else
    f2 t = f2_fma_rm(in, f2_dup(0.0009765625f), f2_dup(8380416.f));
    if (MODE >= 1) t = f2_fma_rn(t, g2, gb2);
    if (MODE == 2) t = f2_min_each(t, TWO23 + 255.f);
    const f2 p = f2_fma_rn(t, w2, c2);
    return f2_add_rm(p, f2_dup(TWO23));
#endif
}

// Two table entries (pixels a and b of this thread) of one job.  e = {ex_a, ex_b, weight_a, weight_b}.
template <int MODE>
__device__ __forceinline__ void ring_pair(const uint4 e, uint32_t box, uint32_t pitch, f2 g2, f2 gb2, const uint8_t* __restrict__ lut,
                                          uint32_t& arg_a, uint32_t& arg_b, uint32_t& abb)
{
    uint32_t ra, ga, ba, rb, gb, bb;
    ring_bilin(e.x, box, pitch, ra, ga, ba);
    ring_bilin(e.y, box, pitch, rb, gb, bb);
    uint32_t kra, krb, kga, kgb, kba, kbb;
    if (MODE == 3) {                                       // per-camera LUT (no f32 multiplier reproduces the f64 rule): scalar path
        const float wa = __uint_as_float(e.z), wb = __uint_as_float(e.w);
        auto one = [&](uint32_t in, float w) {
            const float t = __fmaf_rd(__uint_as_float(in), 0.0009765625f, 8380416.f);
            const float u = (float)__ldg(lut + (__float_as_uint(t) & 255u));
            return __float_as_uint(__fadd_rd(__fmul_rn(u, w), TWO23));
        };
        kra = one(ra, wa); kga = one(ga, wa); kba = one(ba, wa);
        krb = one(rb, wb); kgb = one(gb, wb); kbb = one(bb, wb);
    } else {
        const f2 w2 = f2_pack(e.z, e.w);
        const f2 c2 = f2_mul_rn(w2, f2_dup(-TWO23));
        f2_unpack(ring_chan<MODE>(f2_pack(ra, rb), g2, gb2, w2, c2), kra, krb);
        f2_unpack(ring_chan<MODE>(f2_pack(ga, gb), g2, gb2, w2, c2), kga, kgb);
        f2_unpack(ring_chan<MODE>(f2_pack(ba, bb), g2, gb2, w2, c2), kba, kbb);
    }
    // sums of floor(v W) <= 255 * MAX_CAMS < 2^16: R | G << 16 per pixel, B_a | B_b << 16 per pair.
    // bits(2^23 + k) * 65536 = k << 16 (mod 2^32), so the high halves need no bias removal
    arg_a += kra - TWO23_BITS; arg_a = kga * 65536u + arg_a;
    arg_b += krb - TWO23_BITS; arg_b = kgb * 65536u + arg_b;
    abb += kba - TWO23_BITS; abb = kbb * 65536u + abb;
}

// dst_16s.convertTo(CV_8UC3, 1.0/N) of TWO sums, each given as the bits of 2^23 + acc (acc < 2^16):
// fma(2^23 + acc, inv_n, -2^23 inv_n) = fl(acc * inv_n) because 2^23 inv_n is exact, then sat_u8(rint(.)) -> bits of 2^23 + value
__device__ __forceinline__ void ring_norm2(uint32_t b0, uint32_t b1, f2 inv2, f2 nbias2, uint32_t& o0, uint32_t& o1)
{
    // No saturation needed: every value was clamped to 255 before weighting and the weights of a pixel sum to at most
    // N (1 + n 2^-23) by construction (prep.cpp feather_weights / overwrite_weights), so acc * inv_n < 255.5.
    f2_unpack(f2_add_rn(f2_fma_rn(f2_pack(b0, b1), inv2, nbias2), f2_dup(TWO23)), o0, o1);
}
__device__ __forceinline__ uint32_t lo16_biased(uint32_t v) { return (v & 0xFFFFu) | TWO23_BITS; }          // one LOP3
__device__ __forceinline__ uint32_t hi16_biased(uint32_t v) { return __byte_perm(v, TWO23_BITS, 0x7432); }   // one PRMT: {v.b2, v.b3, 0x00, 0x4B}

template <int GAIN>
__global__ void __launch_bounds__(RG_THREADS, RG_MIN_CTAS) k_blend_ring(const __grid_constant__ RingParams p)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t sbase;                                        // opaque, so the compiler keeps it in a register instead of re-deriving it
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sbase) : "l"(smem));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < RG_NST; s++) { r_mbar_init(sbase + RS_FULL + 8 * s, 1); r_mbar_init(sbase + RS_EMPTY + 8 * s, RG_CONS_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == RG_CONS_WARPS) {
        // ------------------------------------------------------------------ producer warp
        // Programmatic dependent launch: this grid may become resident while the previous kernel of the stream (conversion + gain
        // chain) is still finishing; everything above touched only this CTA's shared memory.  From here on the producer reads what
        // that kernel wrote (gains, RGBX planes through TMA), so it waits for it to complete.  The consumers only ever read data
        // that the producer published, static tables, and write the output frame.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        uint32_t gword = 0u, gclamp = 0u;                  // lane c: gain word of camera c (0: none, f32 bits, all ones: LUT)
        if (GAIN && lane < MAX_CAMS) {
            if (__ldg(p.gain_flag + lane) == 0) {
                const float g = __ldg(p.gain_f32 + lane);
                gword = __float_as_uint(g);
                gclamp = (g * 255.f >= 255.4f) ? 1u : 0u;  // rint(255 g) can exceed 255
            } else
                gword = 0xFFFFFFFFu;
        }
        uint32_t head = 0, i_old = 0, i_new = 0;           // lane 0: ring state; jobs [i_old, i_new) are in flight
        // lane 0: find room for `size` bytes (a multiple of 128) and a descriptor slot; returns the ring offset
        auto acquire = [&](uint32_t size) -> uint32_t {
            for (;;) {
                if (i_new - i_old < (uint32_t)RG_NST) {
                    if (i_new == i_old) { head = size; return 0u; }
                    const uint32_t tail = lds32(sbase + RS_OFF + 4 * (i_old & (RG_NST - 1)));
                    if (head > tail) {
                        if (head + size <= (uint32_t)RING_BYTES) { const uint32_t pos = head; head += size; return pos; }
                        if (size <= tail) { head = size; return 0u; }
                    } else if (head + size <= tail) { const uint32_t pos = head; head += size; return pos; }
                }
                r_mbar_wait(sbase + RS_EMPTY + 8 * (i_old & (RG_NST - 1)), (i_old / RG_NST) & 1u);   // retire the oldest job
                i_old++;
            }
        };
        // A descriptor is published one step late, because it also tells the consumers which job FOLLOWS it (they fetch
        // that job's table entries from global memory while they work on this one).
        struct Desc { uint32_t w0, tmap_cam, box_bytes, pitch_job, gw, flags, txy; };
        auto publish = [&](const Desc& d, uint32_t next_job) {              // lane 0
            const uint32_t size = (d.box_bytes + 127u) & ~127u;
            const uint32_t pos = acquire(size), slot = i_new & (RG_NST - 1);
            const uint32_t full = sbase + RS_FULL + 8 * slot, box = sbase + RS_RING + pos;
            sts32(sbase + RS_OFF + 4 * slot, pos);
            sts128(sbase + RS_META + 32 * slot, make_uint4(box, (d.pitch_job & 0xFFFu) * 4u | (d.flags << 16) | ((d.tmap_cam >> 16) << 24), d.gw, d.txy));
            sts32(sbase + RS_META + 32 * slot + 16, next_job);
#if RING_DEBUG
            { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); asm volatile("st.shared.u64 [%0], %1;" ::"r"(sbase + RS_TISSUE + 8 * slot), "l"(t) : "memory"); }
#endif
            if (d.box_bytes) {
                // the job's table entries: pulled into L2 now, so the consumers' loads (one job ahead of use) are L2 hits
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.entries + (size_t)(d.pitch_job >> 12) * (RING_ENT_BYTES / 16)), "r"(RING_ENT_BYTES) : "memory");
                r_mbar_expect_tx(full, d.box_bytes);
                r_tma_2d(box, (const char*)p.tmaps + (size_t)(d.tmap_cam & 0xFFFFu) * 128, (int)(short)(d.w0 & 0xFFFFu), (int)(short)(d.w0 >> 16), full);
            } else
                r_mbar_arrive(full);
            i_new++;
        };
        Desc pend = { 0u, 0u, 0u, 0u, 0u, RG_EMPTY, 0u };                  // first descriptor: nothing to do but announce the first job
        // tickets: `tile` is being issued, `t1` is known and its records are on their way, `t2` is an atomic in flight
        // the first tile of a CTA is its block index (grid <= ntiles), the others come from the ticket counter
        const int G = (int)gridDim.x;
        int tile = (int)blockIdx.x, t1 = 0;
        if (lane == 0) t1 = G + (int)atomicAdd(p.counter, 1u);
        t1 = __shfl_sync(0xffffffffu, t1, 0);
        uint4 rec = make_uint4(0u, 0u, 0u, 0u);
        if (tile < p.ntiles && lane < p.nslot) rec = __ldg(p.jobs + (size_t)(p.tile0 + tile) * p.nslot + lane);
        while (tile < p.ntiles) {
            uint4 rec1 = make_uint4(0u, 0u, 0u, 0u);
            if (t1 < p.ntiles && lane < p.nslot) rec1 = __ldg(p.jobs + (size_t)(p.tile0 + t1) * p.nslot + lane);
            int t2 = t1;
            if (lane == 0 && t1 < p.ntiles) t2 = G + (int)atomicAdd(p.counter, 1u);
            const int tabs = p.tile0 + tile;
            const uint32_t txy = (uint32_t)(tabs % p.tiles_x) | ((uint32_t)(tabs / p.tiles_x) << 16);
            uint32_t mask = __ballot_sync(0xffffffffu, rec.z != 0u);
            if (mask == 0u) {                                                   // nobody covers this tile: black
                if (lane == 0) { publish(pend, 0xFFFFFFFFu); pend = Desc{ 0u, 0u, 0u, 0u, 0u, RG_EMPTY | RG_LAST, txy }; }
            }
            while (mask) {
                const int k = __ffs(mask) - 1;
                mask &= mask - 1u;
                const uint32_t w0 = __shfl_sync(0xffffffffu, rec.x, k), w1 = __shfl_sync(0xffffffffu, rec.y, k);
                const uint32_t w2 = __shfl_sync(0xffffffffu, rec.z, k), w3 = __shfl_sync(0xffffffffu, rec.w, k);
                const uint32_t cam = (w1 >> 16) & 31u;
                const uint32_t gw = __shfl_sync(0xffffffffu, gword, cam), gc = __shfl_sync(0xffffffffu, gclamp, cam);
                if (lane == 0) {
                    publish(pend, w3 >> 12);
                    pend = Desc{ w0, w1, w2, w3, gw, (mask == 0u ? RG_LAST : 0u) | (gc ? RG_CLAMP : 0u), txy };
                }
            }
            tile = t1; rec = rec1; t1 = __shfl_sync(0xffffffffu, t2, 0);
        }
        if (lane == 0) {
            // every CTA draws exactly one ticket >= ntiles; the CTA that draws the last one re-arms the counter
            // (ntiles - G valid tickets + one failing ticket per CTA = ntiles draws; the last one returns ntiles - 1, seen here as G + ntiles - 1)
            if (tile == p.ntiles + G - 1) { __threadfence(); *(volatile unsigned int*)p.counter = 0u; }
            publish(pend, 0xFFFFFFFFu);
            publish(Desc{ 0u, 0u, 0u, 0u, 0u, RG_EXIT, 0u }, 0xFFFFFFFFu);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const uint32_t patch = sbase + RS_PATCH + 192 * warp;
    const uint4* const ent_base = p.entries + tid;
    uint32_t arg[4] = { 0u, 0u, 0u, 0u }, abb[2] = { 0u, 0u };
    uint32_t soff = 0u, parity = 0u;                       // descriptor slot * 8, phase parity of its barriers
    // One job.  (ea, eb): its table entries, fetched during the previous job; (na, nb): the following job's, fetched now.
    auto stage = [&](const uint4& ea, const uint4& eb, uint4& na, uint4& nb) -> bool {
        r_mbar_wait(sbase + RS_FULL + soff, parity);
        const uint4 m = lds128(sbase + RS_META + 4 * soff);
        const uint32_t next_job = lds32(sbase + RS_META + 4 * soff + 16);
        const uint32_t flags = m.y >> 16;                  // low byte: flags, high byte: camera
        if (flags & RG_EXIT) return false;
        if (next_job != 0xFFFFFFFFu) {                     // in flight under this job's gather
            const uint4* ep = ent_base + (size_t)next_job * (RING_ENT_BYTES / 16);
            na = __ldcs(ep); nb = __ldcs(ep + RG_CONS);
        }
        if (!(flags & RG_EMPTY)) {
            const uint32_t box = m.x, pitch = m.y & 0xFFFFu;
            if (!GAIN) {
                ring_pair<0>(ea, box, pitch, 0ull, 0ull, nullptr, arg[0], arg[1], abb[0]);
                ring_pair<0>(eb, box, pitch, 0ull, 0ull, nullptr, arg[2], arg[3], abb[1]);
            } else if (m.z != 0xFFFFFFFFu) {
                const float g = __uint_as_float(m.z);
                const f2 g2 = f2_dup(g), gb2 = f2_dup(__fmaf_rn(-TWO23, g, TWO23));
                if (flags & RG_CLAMP) {
                    ring_pair<2>(ea, box, pitch, g2, gb2, nullptr, arg[0], arg[1], abb[0]);
                    ring_pair<2>(eb, box, pitch, g2, gb2, nullptr, arg[2], arg[3], abb[1]);
                } else {
                    ring_pair<1>(ea, box, pitch, g2, gb2, nullptr, arg[0], arg[1], abb[0]);
                    ring_pair<1>(eb, box, pitch, g2, gb2, nullptr, arg[2], arg[3], abb[1]);
                }
            } else {
                const uint8_t* lut = p.gain_lut + (flags >> 8) * 256;
                ring_pair<3>(ea, box, pitch, 0ull, 0ull, lut, arg[0], arg[1], abb[0]);
                ring_pair<3>(eb, box, pitch, 0ull, 0ull, lut, arg[2], arg[3], abb[1]);
            }
        }
        __syncwarp();
        if (lane == 0) r_mbar_arrive(sbase + RS_EMPTY + soff);          // this warp is done with the job's ring bytes and descriptor
        soff = (soff + 8u) & (8u * RG_NST - 8u);
        if (soff == 0u) parity ^= 1u;
        if (!(flags & RG_LAST)) return true;

        // ---- tile finished: normalise, RGB -> YUV 4:2:0, store ----
        const int tx0 = (int)(m.w & 0xFFFFu) * TILE_W, ty0 = (int)(m.w >> 16) * TILE_H;
        uint32_t nr[4], ng[4], nb_[4];                    // bits of 2^23 + channel value
        const f2 inv2 = f2_dup(p.inv_n), nbias2 = f2_dup(-TWO23 * p.inv_n);
        ring_norm2(lo16_biased(arg[0]), lo16_biased(arg[1]), inv2, nbias2, nr[0], nr[1]);
        ring_norm2(lo16_biased(arg[2]), lo16_biased(arg[3]), inv2, nbias2, nr[2], nr[3]);
        ring_norm2(hi16_biased(arg[0]), hi16_biased(arg[1]), inv2, nbias2, ng[0], ng[1]);
        ring_norm2(hi16_biased(arg[2]), hi16_biased(arg[3]), inv2, nbias2, ng[2], ng[3]);
        ring_norm2(lo16_biased(abb[0]), hi16_biased(abb[0]), inv2, nbias2, nb_[0], nb_[1]);
        ring_norm2(lo16_biased(abb[1]), hi16_biased(abb[1]), inv2, nbias2, nb_[2], nb_[3]);
        arg[0] = arg[1] = arg[2] = arg[3] = 0u; abb[0] = abb[1] = 0u;
        // colour on the biased integers: the bias terms c * bits(2^23) are constants mod 2^32 (the true sums are < 2^28)
        constexpr uint32_t KY = (1u << 19) + (16u << 20) - (269484u + 528482u + 102760u) * TWO23_BITS;
        constexpr uint32_t KC = (1u << 19) + (128u << 20) - TWO23_BITS;            // U and V coefficient sums are both 1
        if (p.fast_store && tx0 + TILE_W <= p.out_w && ty0 + TILE_H <= p.out_h) {
            #pragma unroll
            for (int q = 0; q < 4; q++)
                sts8(patch + 32 * q + lane, (269484u * nr[q] + 528482u * ng[q] + 102760u * nb_[q] + KY) >> 20);
            if (!(lane & 1)) {
                #pragma unroll
                for (int q = 0; q < 4; q += 2) {                               // rows 2w and 2w + 8 are even
                    sts8(patch + 128 + 8 * q + (lane >> 1), (0u - 155188u * nr[q] - 305135u * ng[q] + 460324u * nb_[q] + KC) >> 20);
                    sts8(patch + 160 + 8 * q + (lane >> 1), (460324u * nr[q] - 385875u * ng[q] - 74448u * nb_[q] + KC) >> 20);
                }
            }
            __syncwarp();
            if (lane < 8) {                                                      // luma: 4 rows x 2 x 16 bytes
                const int q = lane >> 1, row = ty0 + 2 * warp + (q & 1) + 8 * (q >> 1);
                const uint4 v = lds128(patch + 16 * lane);
                *reinterpret_cast<uint4*>(p.oy + (size_t)row * p.oy_pitch + tx0 + 16 * (lane & 1)) = v;
            } else if (lane < 12) {                                              // chroma: {U, V} x 2 rows x 16 bytes
                const int idx = lane - 8, crow = (ty0 >> 1) + warp + 4 * (idx & 1);
                const uint4 v = lds128(patch + 128 + 16 * idx);
                uint8_t* o = (idx >> 1) ? p.ov + (size_t)crow * p.ov_pitch : p.ou + (size_t)crow * p.ou_pitch;
                *reinterpret_cast<uint4*>(o + (tx0 >> 1)) = v;
            }
        } else {
            // generic stores: edge tiles, NV12 / unaligned outputs, RGB result for the after-blend stages
            #pragma unroll
            for (int q = 0; q < 4; q++) {
                const int x = tx0 + lane, y = ty0 + 2 * warp + (q & 1) + 8 * (q >> 1);
                if (x >= p.out_w || y >= p.out_h) continue;
                const int R = (int)(nr[q] & 255u), G = (int)(ng[q] & 255u), B = (int)(nb_[q] & 255u);
                if (p.oy) {
                    p.oy[(size_t)y * p.oy_pitch + x] = (uint8_t)rgb_luma(R, G, B);
                    if (!((x | y) & 1)) {
                        const size_t co = (size_t)(x >> 1) * p.uv_step;
                        p.ou[(size_t)(y >> 1) * p.ou_pitch + co] = (uint8_t)rgb_cb(R, G, B);
                        p.ov[(size_t)(y >> 1) * p.ov_pitch + co] = (uint8_t)rgb_cr(R, G, B);
                    }
                }
                if (p.rgb_out) {
                    uint8_t* o = p.rgb_out + (size_t)y * p.rgb_pitch + 3 * x;
                    o[0] = (uint8_t)R; o[1] = (uint8_t)G; o[2] = (uint8_t)B;
                }
            }
        }
        return true;
    };
    uint4 e0 = make_uint4(0u, 0u, 0u, 0u), e1 = e0, n0 = e0, n1 = e0;
    #pragma unroll 1
    for (;;) {                                             // ping-pong: no register moves between jobs
        if (!stage(e0, e1, n0, n1)) break;
        if (!stage(n0, n1, e0, e1)) break;
    }
}

int ring_ctas_per_sm()
{
    static int v = -1;
    if (v < 0) {
        int a = 0, b = 0;
        cudaFuncSetAttribute(k_blend_ring<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_TOTAL);
        cudaFuncSetAttribute(k_blend_ring<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_TOTAL);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_blend_ring<0>, RG_THREADS, RS_TOTAL) != cudaSuccess) a = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_blend_ring<1>, RG_THREADS, RS_TOTAL) != cudaSuccess) b = 0;
        v = a < b ? a : b;
    }
    return v;
}

void launch_blend_ring(const RingParams& p, int grid, cudaStream_t s)
{
    // Programmatic dependent launch behind k_convert_gain (which signals launch_dependents): the persistent CTAs set themselves
    // up under the tail of the gain chain (measured on C2: 103.3 -> 101.1 us per frame).  OCTVR_PDL=0 launches normally.
    static const bool pdl = [] { const char* e = getenv("OCTVR_PDL"); return !e || atoi(e) != 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(RG_THREADS); cfg.dynamicSmemBytes = RS_TOTAL; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    if (p.use_gain) cudaLaunchKernelEx(&cfg, k_blend_ring<1>, p);
    else cudaLaunchKernelEx(&cfg, k_blend_ring<0>, p);
}

}  // namespace ob
