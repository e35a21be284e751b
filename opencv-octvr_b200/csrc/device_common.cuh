// csrc/device_common.cuh -- device helpers shared by the per-frame kernels (kernels.cu, multiband.cu).
#pragma once
#include "common.h"

namespace ob {

// ------------------------------------------------------------------------------------------------
// shared device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

// fixed-point bilinear of three 8-bit channels from four RGBX taps.  fx, fy in [0, 32).
__device__ __forceinline__ void bilerp_rgbx(uint32_t t00, uint32_t t01, uint32_t t10, uint32_t t11,
                                            uint32_t fx, uint32_t fy, int& r, int& g, int& b)
{
    const uint32_t wx = fx * 65535u + 32u;                 // (32-fx) | fx << 16
    const uint32_t ay = 32u - fy, by = fy;
    const uint32_t rg0 = __byte_perm(t00, t01, 0x5140);    // R00 R01 G00 G01
    const uint32_t bb0 = __byte_perm(t00, t01, 0x6262);    // B00 B01 .. ..
    const uint32_t rg1 = __byte_perm(t10, t11, 0x5140);
    const uint32_t bb1 = __byte_perm(t10, t11, 0x6262);
    const uint32_t hr0 = __dp2a_lo(wx, rg0, 0u), hg0 = __dp2a_hi(wx, rg0, 0u), hb0 = __dp2a_lo(wx, bb0, 0u);
    const uint32_t hr1 = __dp2a_lo(wx, rg1, 0u), hg1 = __dp2a_hi(wx, rg1, 0u), hb1 = __dp2a_lo(wx, bb1, 0u);
    r = (int)((hr0 * ay + hr1 * by + 512u) >> 10);
    g = (int)((hg0 * ay + hg1 * by + 512u) >> 10);
    b = (int)((hb0 * ay + hb1 * by + 512u) >> 10);
}

// gather the four taps for a table entry; taps outside the source contribute 0 (BORDER_CONSTANT)
__device__ __forceinline__ void fetch_taps(const uint32_t* __restrict__ src, int pitch, uint2 c,
                                           uint32_t& t00, uint32_t& t01, uint32_t& t10, uint32_t& t11)
{
    const int off = (int)c.x;
    if (!(c.y & C_BORDER)) {
        t00 = __ldg(src + off); t01 = __ldg(src + off + 1);
        t10 = __ldg(src + off + pitch); t11 = __ldg(src + off + pitch + 1);
    } else {
        const uint32_t in = c.y >> C_TAP_SHIFT;
        t00 = (in & 1u) ? __ldg(src + off) : 0u;
        t01 = (in & 2u) ? __ldg(src + off + 1) : 0u;
        t10 = (in & 4u) ? __ldg(src + off + pitch) : 0u;
        t11 = (in & 8u) ? __ldg(src + off + pitch + 1) : 0u;
    }
}

constexpr float MAGIC_RN = 12582912.f;   // 1.5 * 2^23 : x + MAGIC rounds x to nearest-even integer
constexpr float MAGIC_RD = 8388608.f;    // 2^23 with round-down add : floor(x)

// sat_u8(rint(v * g)) for integer v in [0,255] as a float; g32 has been verified against the f64 rule
__device__ __forceinline__ float gain_apply_f32(float v, float g32)
{
    return fminf(__fadd_rn(__fmaf_rn(v, g32, MAGIC_RN), -MAGIC_RN), 255.f);
}

// The same value from t = 2^23 + v (what the fused kernel's floor-by-magic-add leaves in the register):
// fma(t, g, C) with C = MAGIC_RN - 2^23 g has the same real argument v g + MAGIC_RN as gain_apply_f32, hence the same
// rounding, provided C is exactly representable; gain_tables() verifies both forms against the f64 rule per camera.
__device__ __forceinline__ float gain_bias_f32(float g32) { return __fmaf_rn(-MAGIC_RD, g32, MAGIC_RN); }
__device__ __forceinline__ float gain_apply_biased(float t, float g32, float c)
{
    return fminf(__fadd_rn(__fmaf_rn(t, g32, c), -MAGIC_RN), 255.f);
}

// The form K_blend_ring uses (bias 2^23, valid because v g >= 0): fma(2^23 + v, g, 2^23 - 2^23 g) = 2^23 + rint(v g) when
// 2^23 - 2^23 g is exact in f32; verified per camera by gain_tables() like the two forms above.
__device__ __forceinline__ float gain_apply_two23(float v, float g32)
{
    const float y = __fmaf_rn(__fadd_rn(8388608.f, v), g32, __fmaf_rn(-8388608.f, g32, 8388608.f));
    return __fadd_rn(fminf(y, 8388608.f + 255.f), -8388608.f);
}

// RGB888 -> Y, U, V of cv::cvtColor(RGB2YUV_I420) (imgproc/src/color.cpp:6456-6481), no clamps needed:
// the coefficient sums keep Y in [16,235] and U,V in [16,240] for 8-bit inputs.
__device__ __forceinline__ uint32_t rgb_luma(int R, int G, int B)
{
    return (uint32_t)(269484 * R + 528482 * G + 102760 * B + (1 << 19) + (16 << 20)) >> 20;
}
__device__ __forceinline__ uint32_t rgb_cb(int R, int G, int B)
{
    return (uint32_t)(-155188 * R - 305135 * G + 460324 * B + (1 << 19) + (128 << 20)) >> 20;
}
__device__ __forceinline__ uint32_t rgb_cr(int R, int G, int B)
{
    return (uint32_t)(460324 * R - 385875 * G - 74448 * B + (1 << 19) + (128 << 20)) >> 20;
}

// dst_16s.convertTo(CV_8UC3, 1.0/N): sat_u8(rint((float)acc * (float)(1/N))), packed R | G<<8 | B<<16
__device__ __forceinline__ uint32_t normalise_px(uint32_t ar, uint32_t ag, uint32_t ab, float inv_n)
{
    const int R = min(__float2int_rn(__fmul_rn((float)(int)ar, inv_n)), 255);
    const int G = min(__float2int_rn(__fmul_rn((float)(int)ag, inv_n)), 255);
    const int B = min(__float2int_rn(__fmul_rn((float)(int)ab, inv_n)), 255);
    return (uint32_t)R | ((uint32_t)G << 8) | ((uint32_t)B << 16);
}

// one table entry: four taps from the stage, 1/32-px bilinear, gain, weight, accumulate.
// ex = byte offset of the top-left tap in the stage | fy << 16 | fx << 24 ; ew = f32 weight bits.
// The sums of floor(v * W) are at most 255 * MAX_CAMS < 2^16: R and G share one accumulator (G in the high half), the
// blue sums of two pixels share another.  bits(2^23 + k) * 65536 = k << 16 mod 2^32, so the high halves need no bias
// removal and take their add on the IMAD pipe.
template <int GAIN, bool LUT, bool BHI>
__device__ __forceinline__ void fused_pair(uint32_t ex, uint32_t ew, const uint8_t* __restrict__ s0, const uint8_t* __restrict__ s1,
                                           float g32, float gbias, const uint8_t* __restrict__ lut, uint32_t& arg, uint32_t& ab)
{
    const uint32_t off = ex & 0xFFFFu;
    const uint32_t t00 = *reinterpret_cast<const uint32_t*>(s0 + off), t01 = *reinterpret_cast<const uint32_t*>(s0 + off + 4);
    const uint32_t t10 = *reinterpret_cast<const uint32_t*>(s1 + off), t11 = *reinterpret_cast<const uint32_t*>(s1 + off + 4);
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
    if (GAIN && !LUT) {
        rf = gain_apply_biased(rf, g32, gbias); gf = gain_apply_biased(gf, g32, gbias); bf = gain_apply_biased(bf, g32, gbias);
    } else if (GAIN) {
        rf = (float)__ldg(lut + (__float_as_uint(rf) & 255u)); gf = (float)__ldg(lut + (__float_as_uint(gf) & 255u));
        bf = (float)__ldg(lut + (__float_as_uint(bf) & 255u));
    } else {
        rf = __fadd_rn(rf, -MAGIC_RD); gf = __fadd_rn(gf, -MAGIC_RD); bf = __fadd_rn(bf, -MAGIC_RD);
    }
    // (short)(v * W): f32 product, truncated (v*W >= 0 so floor == trunc)
    const float w = __uint_as_float(ew);
    arg += __float_as_uint(__fadd_rd(__fmul_rn(rf, w), MAGIC_RD)) - 0x4B000000u;
    arg = __float_as_uint(__fadd_rd(__fmul_rn(gf, w), MAGIC_RD)) * 65536u + arg;
    if (BHI) ab = __float_as_uint(__fadd_rd(__fmul_rn(bf, w), MAGIC_RD)) * 65536u + ab;
    else ab += __float_as_uint(__fadd_rd(__fmul_rn(bf, w), MAGIC_RD)) - 0x4B000000u;
}


}  // namespace ob
