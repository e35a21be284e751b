// csrc/multiband.cu -- multiband (Laplacian pyramid) blending with the arithmetic of the reference's CPU
// cv::detail::MultiBandBlender(false, bands, CV_32F) fed 16S images (modules/stitching/src/blenders.cpp:221-477,
// 764-825, 881-892, 923-933; pyramids: modules/imgproc/src/pyramids.cpp:849-1060), which is the contract for
// octvr's blend > 0 mode (Mapper passes seam masks as the blend masks, mapper.cpp:160-165).
//
// Static per mapper (init, host): per-camera bordered rectangles (gap 3*2^bands, aligned to 2^bands, BORDER_REFLECT
// baked into the remap table), f32 weight Gaussian pyramids of the seam masks, summed band weights.
// Per frame (device, all integer-exact; 14 launches for 5 bands):
//   k_mb_warp_staged  remap + gain of every camera into its bordered level-0 image (RGBX8888): one TMA box copy of the
//                     source footprint per 32 x 16 tile, taps from shared memory (k_mb_warp: direct-gather fallback)
//   k_mb_down         Gaussian pyramid level l -> l+1 per camera ([1 4 6 4 1]^2, (v+128)>>8, REFLECT_101), 16S storage,
//                     packed 16-bit-lane arithmetic (camera levels stay within [0, 255])
//   k_mb_band         ONE launch for the destination levels 1..bands: sum over cameras of (short)((G_l - pyrUp(G_{l+1})) *
//                     W_l), then (short)(sum / (sumW + 1e-5)) -- Laplacian, weighting, accumulation, normalisation fused
//   k_mb_collapse     dst_{l-1} += pyrUp(dst_l) (saturating), levels >= 1
//   k_mb_final        level-0 band computed in registers + pyrUp(dst_1), mask, 8-bit narrowing, RGB / YUV 4:2:0 store
// Row-band mappers (multi-GPU partition of one frame) keep a row window of every table and pyramid: see multiband_create.
#include "mapper.h"
#include "prep.h"
#include "device_common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <exception>
#include <climits>

namespace ob {

constexpr int MB_MAX_LEVELS = 9;     // bands <= 8

struct MbCam {
    int x0, y0;                       // top-left of the bordered rect relative to the padded dst roi (level 0)
    int bw, bh;                       // bordered size (level 0), multiples of 2^bands
    unsigned long long off_g[MB_MAX_LEVELS];   // pixel offset of level l in its pool (level 0: g0 pool, else g pool)
    unsigned long long off_w[MB_MAX_LEVELS];   // float offset of the level-l weights
};

// one job of the staged warp: the taps of the tile's valid pixels lie inside the box (bx0, by0, bw, bh) of camera `cam`'s
// RGBX plane; the tile is (tx, ty) in units of 32 x 16 pixels of the camera's bordered level-0 rectangle
struct MbWarpJob { int bx0, by0; uint16_t bw, bh; uint16_t tmap, cam; uint16_t tx, ty; uint32_t pad[3]; };
constexpr int MB_STAGE = 6144;       // shared-memory stage of the staged warp, pixels (24 KB)
constexpr int MB_STAGE_SMALL = 3072; // tiles whose box fits 12 KB run as 128-thread CTAs, twice as many resident
constexpr uint32_t MBW_VALID = 0x80000000u;   // entry: stage offset (13 bits) | fy << 13 | fx << 18 | valid

struct MbParams {
    int n, nb;
    MbCam cam[MAX_CAMS];
    int lw[MB_MAX_LEVELS], lh[MB_MAX_LEVELS];
    unsigned long long off_d[MB_MAX_LEVELS];   // offsets of dst level l (short4 units) and dstw (floats)
    const uint2* coords;              // per camera bordered level-0 pixel: table entry (reflection already applied)
    // staged warp (k_mb_warp_staged): per (camera, 32 x 16 tile of its bordered rectangle) a TMA box of the source plane
    const struct MbWarpJob* wjobs; const uint32_t* wentries; const void* wtmaps;
    const uint2* warp_chunks;         // {camera, 256-pixel chunk} of every chunk with at least one valid entry (the others stay 0)
    const uint16_t* tile_cams;        // per level and 32 x 8 tile of dst: bit c set if camera c has a non-zero weight in the tile
    unsigned long long off_t[MB_MAX_LEVELS];   // first tile of level l in tile_cams
    uint32_t* g0;                     // level 0 images, RGBX8888
    short4* g;                        // levels >= 1, 16S x 3 (+pad)
    const float* w;                   // weight pyramids
    short4* dst;                      // blended Laplacian pyramid
    const float* dstw;                // summed band weights
    int clear_wide;                   // 0 only for the forced-wide diagnostic mode (OCTVR_MB_WIDE=1)
    int* wide;                        // wide[l] != 0: some value of dst level l lies outside [-512, 511] (written by the kernels that
                                      // produce the level, cleared by k_mb_warp at the start of a frame); readers then take the 32-bit pyrUp
    // gain
    const uint32_t* rgbx[MAX_CAMS]; int src_pitch[MAX_CAMS];
    const float* gain_f32; const int* gain_flag; const uint8_t* gain_lut; int use_gain;
    // output
    int rx, ry, rw, rh;               // final result roi in the output frame (dst_roi_final_); row-band mappers: the part inside the row window
    int out_w, out_h;
    int oy0, oy1;                     // output rows this mapper writes (row-band mode; default 0 .. out_h)
    int ox0, ox1;                     // output columns this mapper writes (column-band mode; default 0 .. out_w)
    uint8_t* oy; uint8_t* ou; uint8_t* ov; uint32_t oy_pitch, ou_pitch, ov_pitch; int uv_step;
    uint8_t* rgb_out; uint32_t rgb_pitch;
};

struct Multiband {
    MbParams p;
    uint2* d_chunks = nullptr; uint16_t* d_tile_cams = nullptr; unsigned n_chunks = 0;
    MbWarpJob* d_wjobs = nullptr; uint32_t* d_wentries = nullptr; void* d_wtmaps = nullptr;
    unsigned n_wjobs = 0, n_wsmall = 0;                     // staged warp (0: direct); the first n_wsmall jobs have boxes <= MB_STAGE_SMALL
    uint2* d_coords = nullptr; uint32_t* d_g0 = nullptr; short4* d_g = nullptr; float* d_w = nullptr;
    short4* d_dst = nullptr; float* d_dstw = nullptr; int* d_wide = nullptr; bool force_wide = false;
    int max_bw = 0, max_bh = 0;
    int launches = 0;
    // the small-level chain (down 2.., band 2.., collapse ..3) runs on a side stream next to the level-1 band
    cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; bool fork = false;
};

// ------------------------------------------------------------------------------------------------ device
__device__ __forceinline__ int refl101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}
__device__ __forceinline__ int sat16(int v) { return min(max(v, -32768), 32767); }
__device__ __forceinline__ int3 ld3(const short4* s, int idx) { const short4 v = __ldg(s + idx); return make_int3(v.x, v.y, v.z); }
__device__ __forceinline__ int3 ld3(const uint32_t* s, int idx) { const uint32_t v = __ldg(s + idx); return make_int3(v & 255u, (v >> 8) & 255u, (v >> 16) & 255u); }

// ---- pyrUp of FOUR horizontally adjacent pixels by one thread ----
// pyrUp_ (pyramids.cpp:967-1060) is a separable [1 6 1 | 4 4] filter whose border cases are index extensions: -1 maps to
// 1 (0 for a single sample) and len maps to len - 1 -- the border cases of the reference written out.  Four neighbours share
// their source samples: 12 loads and a third of the arithmetic instead of 4 x 9 loads.  x0 may be even or odd.
__device__ __forceinline__ int up_idx(int k, int len) { return k < 0 ? (len > 1 ? 1 : 0) : (k >= len ? len - 1 : k); }
__device__ __forceinline__ int3 add3(int3 a, int3 b) { return make_int3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ int3 mul3(int3 a, int k) { return make_int3(a.x * k, a.y * k, a.z * k); }
template <class T> __device__ __forceinline__ void pyrup_row4(const T* __restrict__ s, int sw, int r, int x0, int3 (&h)[4])
{
    const T* row = s + (size_t)r * sw;
    const int k = x0 >> 1;
    if ((x0 & 1) == 0) {
        const int3 a = ld3(row, up_idx(k - 1, sw)), b = ld3(row, up_idx(k, sw)), c = ld3(row, up_idx(k + 1, sw)), d = ld3(row, up_idx(k + 2, sw));
        h[0] = add3(add3(a, mul3(b, 6)), c); h[1] = mul3(add3(b, c), 4); h[2] = add3(add3(b, mul3(c, 6)), d); h[3] = mul3(add3(c, d), 4);
    } else {
        const int3 a = ld3(row, up_idx(k, sw)), b = ld3(row, up_idx(k + 1, sw)), c = ld3(row, up_idx(k + 2, sw)), d = ld3(row, up_idx(k + 3, sw));
        h[0] = mul3(add3(a, b), 4); h[1] = add3(add3(a, mul3(b, 6)), c); h[2] = mul3(add3(b, c), 4); h[3] = add3(add3(b, mul3(c, 6)), d);
    }
}
template <class T> __device__ __forceinline__ void pyrup4(const T* __restrict__ s, int sw, int sh, int x0, int y, int3 (&out)[4])
{
    const int ky = y >> 1;
    int3 h1[4], h2[4];
    pyrup_row4(s, sw, up_idx(ky, sh), x0, h1);
    pyrup_row4(s, sw, up_idx(ky + 1, sh), x0, h2);
    if ((y & 1) == 0) {
        int3 h0[4];
        pyrup_row4(s, sw, up_idx(ky - 1, sh), x0, h0);
        #pragma unroll
        for (int q = 0; q < 4; q++) out[q] = add3(add3(h0[q], mul3(h1[q], 6)), h2[q]);
    } else {
        #pragma unroll
        for (int q = 0; q < 4; q++) out[q] = mul3(add3(h1[q], h2[q]), 4);
    }
    // FixPtCast<short, 6>: (v + 32) >> 6.  The taps sum to 64, so |v| <= 64 * 32768 and the result is a short already:
    // the saturate_cast of pyramids.cpp never clips here.
    #pragma unroll
    for (int q = 0; q < 4; q++) out[q] = make_int3((out[q].x + 32) >> 6, (out[q].y + 32) >> 6, (out[q].z + 32) >> 6);
}

// ---- the same for a CAMERA's Gaussian levels, on packed 16-bit lanes ----
// Gaussian levels of 8-bit images stay within [0, 255] (pyrDown is a rounded convex combination), so a camera's short4
// {R, G, B, 0} is filtered as two words of two 16-bit lanes: the largest intermediate is 64 * 255 + 32 < 2^16, no lane
// carries into its neighbour, no unpacking, and (v + 32) >> 6 <= 255 needs no saturation.  Same integers as pyrup4.
__device__ __forceinline__ uint2 ldp(const short4* __restrict__ s, int idx) { return __ldg(reinterpret_cast<const uint2*>(s) + idx); }
__device__ __forceinline__ uint2 p_add(uint2 a, uint2 b) { return make_uint2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ uint2 p_161(uint2 a, uint2 b, uint2 c) { return make_uint2(a.x + 6u * b.x + c.x, a.y + 6u * b.y + c.y); }
__device__ __forceinline__ uint2 p_44(uint2 a, uint2 b) { return make_uint2(4u * (a.x + b.x), 4u * (a.y + b.y)); }
__device__ __forceinline__ void pyrup_row4_packed(const short4* __restrict__ s, int sw, int r, int x0, uint2 (&h)[4])
{
    const short4* row = s + (size_t)r * sw;
    const int k = x0 >> 1;
    if ((x0 & 1) == 0) {
        const uint2 a = ldp(row, up_idx(k - 1, sw)), b = ldp(row, up_idx(k, sw)), c = ldp(row, up_idx(k + 1, sw)), d = ldp(row, up_idx(k + 2, sw));
        h[0] = p_161(a, b, c); h[1] = p_44(b, c); h[2] = p_161(b, c, d); h[3] = p_44(c, d);
    } else {
        const uint2 a = ldp(row, up_idx(k, sw)), b = ldp(row, up_idx(k + 1, sw)), c = ldp(row, up_idx(k + 2, sw)), d = ldp(row, up_idx(k + 3, sw));
        h[0] = p_44(a, b); h[1] = p_161(a, b, c); h[2] = p_44(b, c); h[3] = p_161(b, c, d);
    }
}
// out[q] = {R | G << 16, B} of the up-sampled level at (x0 + q, y)
__device__ __forceinline__ void pyrup4_packed(const short4* __restrict__ s, int sw, int sh, int x0, int y, uint2 (&out)[4])
{
    const int ky = y >> 1;
    uint2 h1[4], h2[4];
    pyrup_row4_packed(s, sw, up_idx(ky, sh), x0, h1);
    pyrup_row4_packed(s, sw, up_idx(ky + 1, sh), x0, h2);
    if ((y & 1) == 0) {
        uint2 h0[4];
        pyrup_row4_packed(s, sw, up_idx(ky - 1, sh), x0, h0);
        #pragma unroll
        for (int q = 0; q < 4; q++) out[q] = p_161(h0[q], h1[q], h2[q]);
    } else {
        #pragma unroll
        for (int q = 0; q < 4; q++) out[q] = p_44(h1[q], h2[q]);
    }
    #pragma unroll
    for (int q = 0; q < 4; q++) out[q] = make_uint2(((out[q].x + 0x00200020u) >> 6) & 0x03FF03FFu, (out[q].y + 32u) >> 6);
}
// (short)(lap * weight) for the three channels of one pixel, lap = g - up (within [-255, 255]: the saturating 16S
// subtraction of createLaplacePyr cannot saturate), f32 product truncated toward zero (blenders.cpp:418-420)
__device__ __forceinline__ void lap_weight_acc(int gr, int gg, int gb, uint2 up, bool lap, float w, int (&acc)[3])
{
    if (lap) { gr -= (int)(up.x & 0xFFFFu); gg -= (int)(up.x >> 16); gb -= (int)up.y; }
    acc[0] += __float2int_rz(__fmul_rn((float)gr, w));
    acc[1] += __float2int_rz(__fmul_rn((float)gg, w));
    acc[2] += __float2int_rz(__fmul_rn((float)gb, w));
}

// ---- and for the BLENDED pyramid (signed 16S) when all its values lie within [-512, 511] ----
// v + 512 fits 10 bits, 64 * 1023 + 32 < 2^16, and because the taps always sum to 64,
// (sum w (v + 512) + 32) >> 6 = ((sum w v + 32) >> 6) + 512 exactly: the packed filter on biased lanes gives the same
// integers as pyrup4<short4>.  v + 512 = (v ^ 0x200) & 0x3FF for a 16-bit two's complement v in range.  Whether a level
// is in range is recorded by the kernels that write it (MbParams::wide); out-of-range levels (possible in principle:
// |collapsed value| <= 255 * levels) take the 32-bit path.
__device__ __forceinline__ uint2 ldp_biased(const short4* __restrict__ s, int idx)
{
    const uint2 v = ldp(s, idx);
    return make_uint2((v.x ^ 0x02000200u) & 0x03FF03FFu, (v.y ^ 0x00000200u) & 0x000003FFu);
}
__device__ __forceinline__ void pyrup_row4_biased(const short4* __restrict__ s, int sw, int r, int x0, uint2 (&h)[4])
{
    const short4* row = s + (size_t)r * sw;
    const int k = x0 >> 1;
    if ((x0 & 1) == 0) {
        const uint2 a = ldp_biased(row, up_idx(k - 1, sw)), b = ldp_biased(row, up_idx(k, sw)), c = ldp_biased(row, up_idx(k + 1, sw)), d = ldp_biased(row, up_idx(k + 2, sw));
        h[0] = p_161(a, b, c); h[1] = p_44(b, c); h[2] = p_161(b, c, d); h[3] = p_44(c, d);
    } else {
        const uint2 a = ldp_biased(row, up_idx(k, sw)), b = ldp_biased(row, up_idx(k + 1, sw)), c = ldp_biased(row, up_idx(k + 2, sw)), d = ldp_biased(row, up_idx(k + 3, sw));
        h[0] = p_44(a, b); h[1] = p_161(a, b, c); h[2] = p_44(b, c); h[3] = p_161(b, c, d);
    }
}
__device__ __forceinline__ void pyrup4_narrow(const short4* __restrict__ s, int sw, int sh, int x0, int y, int3 (&out)[4])
{
    const int ky = y >> 1;
    uint2 h1[4], h2[4], o[4];
    pyrup_row4_biased(s, sw, up_idx(ky, sh), x0, h1);
    pyrup_row4_biased(s, sw, up_idx(ky + 1, sh), x0, h2);
    if ((y & 1) == 0) {
        uint2 h0[4];
        pyrup_row4_biased(s, sw, up_idx(ky - 1, sh), x0, h0);
        #pragma unroll
        for (int q = 0; q < 4; q++) o[q] = p_161(h0[q], h1[q], h2[q]);
    } else {
        #pragma unroll
        for (int q = 0; q < 4; q++) o[q] = p_44(h1[q], h2[q]);
    }
    #pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t lo = (o[q].x + 0x00200020u) >> 6, hi = (o[q].y + 32u) >> 6;
        out[q] = make_int3((int)(lo & 0x3FFu) - 512, (int)((lo >> 16) & 0x3FFu) - 512, (int)hi - 512);
    }
}
// pyrUp of a blended level: packed when the level is known to be narrow, 32-bit otherwise (uniform over the grid)
__device__ __forceinline__ void pyrup4_dst(const MbParams& p, int l, int x0, int y, int3 (&out)[4])
{
    if (__ldg(p.wide + l) == 0) pyrup4_narrow(p.dst + p.off_d[l], p.lw[l], p.lh[l], x0, y, out);
    else pyrup4(p.dst + p.off_d[l], p.lw[l], p.lh[l], x0, y, out);
}
__device__ __forceinline__ void note_range(const MbParams& p, int l, int r, int g, int b)
{
    if ((unsigned)(r + 512) > 1023u || (unsigned)(g + 512) > 1023u || (unsigned)(b + 512) > 1023u) atomicOr(p.wide + l, 1);
}

// ---- k_mb_warp: one CTA per FOUR 256-pixel chunks (of any cameras' bordered level-0 images) that have a valid entry.
//      The chain table entry -> four taps is two dependent DRAM round trips; with one pixel per thread the kernel was bound
//      by that latency (ncu: 26 long-scoreboard stalls per issue at 85 % occupancy).  A thread now loads its four table
//      entries first, then all sixteen taps, then does the arithmetic: four independent chains in flight per thread. ----
constexpr int MB_WARP_CHUNKS = 4;
// The four chunks of a CTA belong to ONE camera (the host pads every camera's part of the list to a multiple of four with
// chunk index 0xFFFFFFFF), so the source plane, pitch and gain are CTA-uniform and stay out of the per-pixel code.
__global__ void __launch_bounds__(256) k_mb_warp(const __grid_constant__ MbParams p)
{
    if (blockIdx.x == 0 && threadIdx.x < MB_MAX_LEVELS && p.clear_wide) p.wide[threadIdx.x] = 0;   // new frame: every level narrow until proven wide
    const uint2* list = p.warp_chunks + (size_t)blockIdx.x * MB_WARP_CHUNKS;
    const int c = (int)__ldg(&list[0].x);
    const MbCam& cam = p.cam[c];
    const unsigned npx = (unsigned)(cam.bw * cam.bh);
    const uint2* __restrict__ coords = p.coords + cam.off_g[0];
    uint32_t* __restrict__ g0 = p.g0 + cam.off_g[0];
    const uint32_t* __restrict__ src = p.rgbx[c];
    const int pitch = p.src_pitch[c];
    unsigned t[MB_WARP_CHUNKS];
    uint2 cc[MB_WARP_CHUNKS];
    #pragma unroll
    for (int j = 0; j < MB_WARP_CHUNKS; j++) {
        const unsigned k = __ldg(&list[j].y);
        t[j] = k == 0xFFFFFFFFu ? 0xFFFFFFFFu : k * 256u + threadIdx.x;
        if (t[j] >= npx) t[j] = 0xFFFFFFFFu;
    }
    #pragma unroll
    for (int j = 0; j < MB_WARP_CHUNKS; j++) cc[j] = t[j] != 0xFFFFFFFFu ? __ldcs(coords + t[j]) : make_uint2(0u, 0u);
    uint32_t tap[MB_WARP_CHUNKS][4];
    #pragma unroll
    for (int j = 0; j < MB_WARP_CHUNKS; j++) {
        tap[j][0] = tap[j][1] = tap[j][2] = tap[j][3] = 0u;
        if (cc[j].y & C_VALID) fetch_taps(src, pitch, cc[j], tap[j][0], tap[j][1], tap[j][2], tap[j][3]);
    }
    // gain: verified f32 multiplier, or the exact LUT when the camera is flagged (kernels.cu, gain_tables)
    const bool use_lut = p.use_gain && __ldg(p.gain_flag + c) != 0;
    const float g32 = p.use_gain ? __ldg(p.gain_f32 + c) : 1.f;
    const uint8_t* lut = p.gain_lut + c * 256;
    #pragma unroll
    for (int j = 0; j < MB_WARP_CHUNKS; j++) {
        if (t[j] == 0xFFFFFFFFu) continue;
        uint32_t px = 0;
        if (cc[j].y & C_VALID) {
            int r, g, b;
            bilerp_rgbx(tap[j][0], tap[j][1], tap[j][2], tap[j][3], cc[j].y & 31u, (cc[j].y >> 5) & 31u, r, g, b);
            if (use_lut) { r = __ldg(lut + r); g = __ldg(lut + g); b = __ldg(lut + b); }
            else if (p.use_gain) { r = (int)gain_apply_f32((float)r, g32); g = (int)gain_apply_f32((float)g, g32); b = (int)gain_apply_f32((float)b, g32); }
            px = (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
        }
        g0[t[j]] = px;
    }
}

// ---- k_mb_warp_staged: the same warp with the source footprint of a 32 x 16 tile brought into shared memory by ONE TMA
//      box copy (UTMALDG, zero fill outside the plane = BORDER_CONSTANT 0) and the four taps read from there.  The direct
//      kernel's scattered 4-byte gathers made DRAM fetch random sectors (L2 hit rate 38 %); the box copy reads whole rows,
//      and a 4-byte entry (stage offset | fractions) replaces the 8-byte one.  256 threads, two pixels each. ----
__device__ __forceinline__ void mbw_mbar_init(uint64_t* mbar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbw_tma_box(void* smem_dst, const void* tmap, int x, int y, uint32_t bytes, uint64_t* mbar)
{
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(mbar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"((uint64_t)tmap), "r"(mb), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void mbw_wait(uint64_t* mbar)
{
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "MBW_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0, %1;\n\t"
        "@P1 bra MBW_DONE;\n\t"
        "bra MBW_WAIT;\n\t"
        "MBW_DONE:\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(0x989680u) : "memory");
}
// THREADS x (512 / THREADS) pixels per CTA, STAGE_PX pixels of stage: the kernel is bound by the latency of a tile (job
// record -> box copy -> taps -> store) times the tiles in flight per SM, so tiles whose box fits 12 KB run as 128-thread CTAs
// (16 resident per SM instead of 8); the few larger ones take the 256-thread / 24 KB variant.
template <int THREADS, int STAGE_PX>
__global__ void __launch_bounds__(THREADS) k_mb_warp_staged(const __grid_constant__ MbParams p, const unsigned first_job)
{
    __shared__ __align__(128) uint32_t s_buf[STAGE_PX];
    __shared__ __align__(8) uint64_t s_mbar;
    if (blockIdx.x == 0 && first_job == 0 && threadIdx.x < MB_MAX_LEVELS && p.clear_wide) p.wide[threadIdx.x] = 0;   // new frame: every level narrow until proven wide
    constexpr int ROWS = THREADS / 32, PPT = TILE_H / ROWS;   // rows of the tile covered at once, pixels per thread
    const int tid = threadIdx.x, lx = tid & 31, ly = tid >> 5;
    const unsigned j = first_job + blockIdx.x;
    const MbWarpJob job = p.wjobs[j];
    if (tid == 0) mbw_mbar_init(&s_mbar);
    __syncthreads();
    if (tid == 0) mbw_tma_box(s_buf, (const char*)p.wtmaps + (size_t)job.tmap * 128, job.bx0, job.by0, (uint32_t)job.bw * job.bh * 4u, &s_mbar);
    const uint32_t* ent = p.wentries + (size_t)j * TILE_PX + tid;
    uint32_t e[PPT];
    #pragma unroll
    for (int h = 0; h < PPT; h++) e[h] = __ldcs(ent + h * THREADS);
    const int c = job.cam;
    const MbCam& cam = p.cam[c];
    const bool use_lut = p.use_gain && __ldg(p.gain_flag + c) != 0;
    const float g32 = p.use_gain ? __ldg(p.gain_f32 + c) : 1.f, gb = gain_bias_f32(g32);
    const uint8_t* lut = p.gain_lut + c * 256;
    const uint8_t* s0 = reinterpret_cast<const uint8_t*>(s_buf);
    const uint8_t* s1 = s0 + job.bw * 4;
    // whole tiles (all but the right / bottom edge of a rectangle) skip the per-pixel bounds checks
    const int x = job.tx * TILE_W + lx, yb = job.ty * TILE_H + ly;
    const bool whole = (job.tx + 1) * TILE_W <= cam.bw && (job.ty + 1) * TILE_H <= cam.bh;
    uint32_t* __restrict__ orow = p.g0 + cam.off_g[0] + (size_t)yb * cam.bw + x;
    const size_t ostep = (size_t)ROWS * cam.bw;
    if (tid < 32) mbw_wait(&s_mbar);                        // one warp polls, the others park at the barrier
    __syncthreads();
    #pragma unroll
    for (int h = 0; h < PPT; h++, orow += ostep) {
        if (!whole && (x >= cam.bw || yb + ROWS * h >= cam.bh)) continue;
        // the gather of device_common.cuh (fused_pair) without the blend weight: IDP.2A bilinear with the +512 rounding term,
        // floor(x / 1024) by a round-down fma that leaves 2^23 + v in the register, gain on that biased value.  An invalid
        // entry reads the start of the stage and is masked at the end: no divergence inside the tile.
        const uint32_t off = (e[h] & 0x1FFFu) << 2;
        const uint32_t t00 = *reinterpret_cast<const uint32_t*>(s0 + off), t01 = *reinterpret_cast<const uint32_t*>(s0 + off + 4);
        const uint32_t t10 = *reinterpret_cast<const uint32_t*>(s1 + off), t11 = *reinterpret_cast<const uint32_t*>(s1 + off + 4);
        const uint32_t fx = (e[h] >> 18) & 31u, fy = (e[h] >> 13) & 31u;
        const uint32_t wx = fx * 65535u + 32u;                 // (32-fx) | fx << 16
        const uint32_t wb = wx * fy, wt = wx * 32u - wb;
        const uint32_t rg0 = __byte_perm(t00, t01, 0x5140), bb0 = __byte_perm(t00, t01, 0x6262);
        const uint32_t rg1 = __byte_perm(t10, t11, 0x5140), bb1 = __byte_perm(t10, t11, 0x6262);
        const uint32_t r = __dp2a_lo(wb, rg1, __dp2a_lo(wt, rg0, 512u));
        const uint32_t g = __dp2a_hi(wb, rg1, __dp2a_hi(wt, rg0, 512u));
        const uint32_t b = __dp2a_lo(wb, bb1, __dp2a_lo(wt, bb0, 512u));
        const float rf = __fmaf_rd(__uint2float_rn(r), 0.0009765625f, MAGIC_RD);   // bits = 0x4B000000 + floor(r / 1024)
        const float gf = __fmaf_rd(__uint2float_rn(g), 0.0009765625f, MAGIC_RD);
        const float bf = __fmaf_rd(__uint2float_rn(b), 0.0009765625f, MAGIC_RD);
        uint32_t R = __float_as_uint(rf) & 255u, G = __float_as_uint(gf) & 255u, B = __float_as_uint(bf) & 255u;
        if (use_lut) { R = __ldg(lut + R); G = __ldg(lut + G); B = __ldg(lut + B); }
        else if (p.use_gain) {
            // fma(2^23 + v, g, MAGIC_RN - 2^23 g) = MAGIC_RN + rint(v g) (gain_apply_biased): the integer sits in the mantissa
            R = min(__float_as_uint(__fmaf_rn(rf, g32, gb)) - 0x4B400000u, 255u);
            G = min(__float_as_uint(__fmaf_rn(gf, g32, gb)) - 0x4B400000u, 255u);
            B = min(__float_as_uint(__fmaf_rn(bf, g32, gb)) - 0x4B400000u, 255u);
        }
        const uint32_t px = R | (G << 8) | (B << 16);
        *orow = (e[h] & MBW_VALID) ? px : 0u;
    }
}

// ---- k_mb_down: level l -> l+1 for every camera.  grid (ceil(w/32), ceil(h/8), cameras) at level l+1 ----
// Column-strip version: a thread produces FOUR vertically adjacent pixels of level l+1.  It walks the 11 source rows they
// depend on, forms the five-tap horizontal sum of each row once and scatters it (x 1, 4, 6, 4, 1) into the outputs that use
// the row -- 55 loads for four outputs instead of 100, no shared memory, no barrier.  Integer arithmetic, so the separable
// order gives the same numbers as the 5 x 5 form of pyramids.cpp:849-964.  Threads of a warp walk adjacent
// columns, so every load instruction is a contiguous (stride-2) row segment.
//
// Levels >= 1 of a CAMERA pyramid: values within [0, 255] stored as short4 {R, G, B, 0}, so the words {R | G << 16, B} are
// filtered as packed 16-bit lanes (5 x 5 sum <= 256 * 255 < 2^16) with 16-byte loads in the interior, rows fetched in
// groups ahead of their use and no early exit (see mb_down_strip_u8).  The same integers as the 5 x 5 form.
__device__ __forceinline__ void mb_down_strip_p16(const short4* __restrict__ src, short4* __restrict__ dst, int sw, int sh, int dw, int dh, int x, int y0)
{
    const bool interior = 2 * x - 2 >= 0 && 2 * x + 2 < sw;
    int xi[5];
    #pragma unroll
    for (int k = 0; k < 5; k++) xi[k] = refl101(2 * x + k - 2, sw);
    uint32_t alo[4], ahi[4];
    #pragma unroll
    for (int j = 0; j < 4; j++) alo[j] = ahi[j] = 0u;
    const bool inner_rows = 2 * y0 - 2 >= 0 && 2 * y0 + 8 < sh;
    #pragma unroll
    for (int r0 = 0; r0 < 11; r0 += 3) {
        uint2 P[3][5];
        #pragma unroll
        for (int i = 0; i < 3; i++) {
            const int r = r0 + i;
            if (r >= 11) break;
            const short4* row = src + (size_t)(inner_rows ? 2 * y0 - 2 + r : refl101(2 * y0 - 2 + r, sh)) * sw;
            if (interior) {
                const uint4 ab = __ldg(reinterpret_cast<const uint4*>(row + 2 * x - 2)), cd = __ldg(reinterpret_cast<const uint4*>(row + 2 * x));
                P[i][0] = make_uint2(ab.x, ab.y); P[i][1] = make_uint2(ab.z, ab.w); P[i][2] = make_uint2(cd.x, cd.y); P[i][3] = make_uint2(cd.z, cd.w);
                P[i][4] = ldp(row, 2 * x + 2);
            } else {
                #pragma unroll
                for (int k = 0; k < 5; k++) P[i][k] = ldp(row, xi[k]);
            }
        }
        #pragma unroll
        for (int i = 0; i < 3; i++) {
            const int r = r0 + i;
            if (r >= 11) break;
            const uint32_t hlo = (P[i][0].x + P[i][4].x) + 4u * (P[i][1].x + P[i][3].x) + 6u * P[i][2].x;
            const uint32_t hhi = (P[i][0].y + P[i][4].y) + 4u * (P[i][1].y + P[i][3].y) + 6u * P[i][2].y;
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int t = r - 2 * j;
                if (t < 0 || t > 4) continue;
                const uint32_t kw = t == 2 ? 6u : (t == 1 || t == 3) ? 4u : 1u;
                alo[j] += kw * hlo; ahi[j] += kw * hhi;
            }
        }
    }
    #pragma unroll
    for (int j = 0; j < 4; j++)
        if (y0 + j < dh)
            *reinterpret_cast<uint2*>(dst + (size_t)(y0 + j) * dw + x) = make_uint2(((alo[j] + 0x00800080u) >> 8) & 0x00FF00FFu, (ahi[j] + 128u) >> 8);
}

// Level 0 -> 1: the source is RGBX8888 with X = 0.  ncu showed the generic strip ALU-pipe bound (75 % of the LOP3 / SHF /
// PRMT / IADD3 pipe: byte unpacking and index arithmetic, 107 instructions per source row).  Here the five pixels of a row
// come from two 8-byte loads and one 4-byte load (interior columns), their bytes are transposed per channel with seven
// PRMTs, the [1 4 6 4 | 1] row sums are two IDP.4A per channel (IMAD pipe), and R, B then travel as two 16-bit lanes: the
// full 5 x 5 sum is at most 256 * 255 < 2^16, so no lane carries and ((sum + 128) >> 8) <= 255 needs no saturation.  The
// same integers as the 5 x 5 form.
__device__ __forceinline__ void mb_down_strip_u8(const uint32_t* __restrict__ src, short4* __restrict__ dst, int sw, int sh, int dw, int dh, int x, int y0)
{
    const bool interior = 2 * x - 2 >= 0 && 2 * x + 2 < sw;           // no column reflection, 8-byte aligned pairs (sw is even)
    int xi[5];
    #pragma unroll
    for (int k = 0; k < 5; k++) xi[k] = refl101(2 * x + k - 2, sw);
    uint32_t arb[4], ag[4];
    #pragma unroll
    for (int j = 0; j < 4; j++) arb[j] = ag[j] = 0u;
    const bool inner_rows = 2 * y0 - 2 >= 0 && 2 * y0 + 8 < sh;       // no row reflection for any of the 11 rows
    // The kernel is bound by load latency once the arithmetic is this short (ncu: 14 long-scoreboard stalls per issue): no
    // early exit for ragged strips (rows past the image are reflected like any other and their outputs never stored), so
    // every load of the strip is independent of control flow and the rows are fetched in groups ahead of their use.
    #pragma unroll
    for (int r0 = 0; r0 < 11; r0 += 4) {
      uint32_t A[4], B[4], C[4], D[4], E[4];
      #pragma unroll
      for (int i = 0; i < 4; i++) {
        const int r = r0 + i;
        if (r >= 11) break;
        const int sy = inner_rows ? 2 * y0 - 2 + r : refl101(2 * y0 - 2 + r, sh);
        const uint32_t* row = src + (size_t)sy * sw;
        if (interior) {
            const uint2 ab = __ldg(reinterpret_cast<const uint2*>(row + 2 * x - 2)), cd = __ldg(reinterpret_cast<const uint2*>(row + 2 * x));
            A[i] = ab.x; B[i] = ab.y; C[i] = cd.x; D[i] = cd.y; E[i] = __ldg(row + 2 * x + 2);
        } else {
            A[i] = __ldg(row + xi[0]); B[i] = __ldg(row + xi[1]); C[i] = __ldg(row + xi[2]); D[i] = __ldg(row + xi[3]); E[i] = __ldg(row + xi[4]);
        }
      }
      #pragma unroll
      for (int i = 0; i < 4; i++) {
        const int r = r0 + i;
        if (r >= 11) break;
        const uint32_t a = A[i], b = B[i], c = C[i], d = D[i], e = E[i];
        const uint32_t rg_ab = __byte_perm(a, b, 0x5140), rg_cd = __byte_perm(c, d, 0x5140);       // Ra Rb Ga Gb | Rc Rd Gc Gd
        const uint32_t r4 = __byte_perm(rg_ab, rg_cd, 0x5410), g4 = __byte_perm(rg_ab, rg_cd, 0x7632);
        const uint32_t b4 = __byte_perm(__byte_perm(a, b, 0x4462), __byte_perm(c, d, 0x4462), 0x5410);
        const uint32_t W = 0x04060401u;                                // weights of a, b, c, d; e has weight 1
        const uint32_t hr = __dp4a(r4, W, __dp4a(e, 0x00000001u, 0u));
        const uint32_t hg = __dp4a(g4, W, __dp4a(e, 0x00000100u, 0u));
        const uint32_t hb = __dp4a(b4, W, __dp4a(e, 0x00010000u, 0u));
        const uint32_t hrb = hb * 65536u + hr;                         // R | B << 16 (each <= 16 * 255)
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const int t = r - 2 * j;                                  // tap of output j that reads this row
            if (t < 0 || t > 4) continue;
            const uint32_t kw = t == 2 ? 6u : (t == 1 || t == 3) ? 4u : 1u;
            arb[j] += kw * hrb; ag[j] += kw * hg;
        }
      }
    }
    #pragma unroll
    for (int j = 0; j < 4; j++)
        if (y0 + j < dh) {
            const uint32_t rb = ((arb[j] + 0x00800080u) >> 8) & 0x00FF00FFu, g = (ag[j] + 128u) >> 8;
            // short4 {R, G, B, 0} as two words
            *reinterpret_cast<uint2*>(dst + (size_t)(y0 + j) * dw + x) = make_uint2((rb & 0xFFFFu) | (g << 16), rb >> 16);
        }
}

template <bool L0> __global__ void __launch_bounds__(256) k_mb_down(const __grid_constant__ MbParams p, int l)
{
    const MbCam& cam = p.cam[blockIdx.z];
    const int sw = cam.bw >> l, sh = cam.bh >> l, dw = sw >> 1, dh = sh >> 1;
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = (blockIdx.y * 8 + threadIdx.y) * 4;
    if (x >= dw || y0 >= dh) return;
    if (L0) mb_down_strip_u8(p.g0 + cam.off_g[0], p.g + cam.off_g[1], sw, sh, dw, dh, x, y0);
    else mb_down_strip_p16(p.g + cam.off_g[l], p.g + cam.off_g[l + 1], sw, sh, dw, dh, x, y0);
}

// ---- k_mb_band: one thread per FOUR horizontally adjacent pixels of destination level l (CTA = 128 x 8 pixels) ----
//      Levels do not depend on each other, so one launch covers all of them (blockIdx.z + 1 = level): the small levels,
//      latency-bound on their own (15 us each for a few thousand pixels), hide under level 1.
__global__ void __launch_bounds__(256, 4) k_mb_band(const __grid_constant__ MbParams p, const int first_level)
{
    const int l = (int)blockIdx.z + first_level;
    const int X0 = (blockIdx.x * 32 + threadIdx.x) * 4, Y = blockIdx.y * 8 + threadIdx.y;
    const int lw = p.lw[l], lh = p.lh[l];
    if ((int)blockIdx.x * 128 >= lw || (int)blockIdx.y * 8 >= lh) return;
    // only the cameras that have a non-zero weight somewhere in the CTA's four 32 x 8 tiles (at level 0 the weights are the
    // hard seam masks: usually one camera); uniform over the CTA
    unsigned cams = 0;
    {
        const int tiles_x = (lw + 31) / 32;
        const uint16_t* tm = p.tile_cams + p.off_t[l] + (size_t)blockIdx.y * tiles_x;
        #pragma unroll
        for (int k = 0; k < 4; k++) if ((int)blockIdx.x * 4 + k < tiles_x) cams |= __ldg(tm + blockIdx.x * 4 + k);
    }
    if (X0 >= lw || Y >= lh) return;
    int acc[4][3];
    #pragma unroll
    for (int q = 0; q < 4; q++) acc[q][0] = acc[q][1] = acc[q][2] = 0;
    for (; cams; cams &= cams - 1) {
        const int c = __ffs(cams) - 1;
        const MbCam& cam = p.cam[c];
        const int x0 = X0 - (cam.x0 >> l), y = Y - (cam.y0 >> l);
        const int w_l = cam.bw >> l, h_l = cam.bh >> l;
        if (y < 0 || y >= h_l || x0 + 3 < 0 || x0 >= w_l) continue;
        // interior groups: no per-pixel bounds checks or branches (a zero weight adds (short)(lap * 0) = 0); the weight and
        // the level-l samples are fetched together
        const float* wrow = p.w + cam.off_w[l] + (size_t)y * w_l + x0;
        const short4* grow = p.g + cam.off_g[l] + (size_t)y * w_l + x0;
        const bool in4 = x0 >= 0 && x0 + 3 < w_l;
        float w[4];
        uint2 gl[4];
        #pragma unroll
        for (int q = 0; q < 4; q++) {
            const bool in = in4 || (x0 + q >= 0 && x0 + q < w_l);
            w[q] = in ? __ldg(wrow + q) : 0.f;
            gl[q] = in ? ldp(grow, q) : make_uint2(0u, 0u);
        }
        if (w[0] == 0.f && w[1] == 0.f && w[2] == 0.f && w[3] == 0.f) continue;
        uint2 up[4];
        const bool lap = l < p.nb;                                    // createLaplacePyr: pyr[l] -= pyrUp(pyr[l+1]); the top level stays Gaussian
        if (lap) pyrup4_packed(p.g + cam.off_g[l + 1], w_l >> 1, h_l >> 1, x0, y, up);
        #pragma unroll
        for (int q = 0; q < 4; q++)
            lap_weight_acc((int)(gl[q].x & 0xFFFFu), (int)(gl[q].x >> 16), (int)gl[q].y, up[q], lap, w[q], acc[q]);
    }
    #pragma unroll
    for (int q = 0; q < 4; q++) {
        if (X0 + q >= lw) break;
        const size_t di = p.off_d[l] + (size_t)Y * lw + X0 + q;
        // normalizeUsingWeightMap (blenders.cpp:788-797): (short)(v / (w + 1e-5f))
        const float den = __fadd_rn(__ldg(p.dstw + di), 1e-5f);
        const short r = (short)__float2int_rz(__fdiv_rn((float)(short)acc[q][0], den));
        const short g = (short)__float2int_rz(__fdiv_rn((float)(short)acc[q][1], den));
        const short b = (short)__float2int_rz(__fdiv_rn((float)(short)acc[q][2], den));
        note_range(p, l, r, g, b);
        p.dst[di] = make_short4(r, g, b, 0);
    }
}

// ---- k_mb_collapse: dst_{l-1} = sat(pyrUp(dst_l) + dst_{l-1}), l >= 2; four pixels per thread ----
__global__ void __launch_bounds__(256) k_mb_collapse(const __grid_constant__ MbParams p, int l)
{
    const int X0 = (blockIdx.x * 32 + threadIdx.x) * 4, Y = blockIdx.y * 8 + threadIdx.y;
    const int w = p.lw[l - 1], h = p.lh[l - 1];
    if (X0 >= w || Y >= h) return;
    short4 curv[4];                                                  // fetched before the pyrUp: both load groups in flight together
    #pragma unroll
    for (int q = 0; q < 4; q++) curv[q] = X0 + q < w ? p.dst[p.off_d[l - 1] + (size_t)Y * w + X0 + q] : make_short4(0, 0, 0, 0);
    int3 up[4];
    pyrup4_dst(p, l, X0, Y, up);
    #pragma unroll
    for (int q = 0; q < 4; q++) {
        if (X0 + q >= w) break;
        const size_t di = p.off_d[l - 1] + (size_t)Y * w + X0 + q;
        const short4 cur = curv[q];
        const int r = sat16(up[q].x + cur.x), g = sat16(up[q].y + cur.y), b = sat16(up[q].z + cur.z);
        note_range(p, l - 1, r, g, b);
        p.dst[di] = make_short4((short)r, (short)g, (short)b, 0);
    }
}

// ---- k_mb_final: destination level 0 never exists in memory.  One thread per FOUR OUTPUT-FRAME pixels computes the level-0
//      band (k_mb_band with l = 0: Laplacian, weighting, accumulation over the cameras, normalisation), adds pyrUp(dst_1)
//      (the last collapse step), masks, narrows to 8 bit and stores RGB / YUV 4:2:0.  Pixels outside the result roi are
//      black.  Saves the 8 B/px write + read of dst_0 and one launch. ----
__global__ void __launch_bounds__(256, 4) k_mb_final(const __grid_constant__ MbParams p)
{
    const int X0 = p.ox0 + (blockIdx.x * 32 + threadIdx.x) * 4, Y = p.oy0 + blockIdx.y * 8 + threadIdx.y;
    // cameras with a non-zero level-0 weight somewhere under this CTA (128 x 8 output pixels = up to 5 x 2 weight tiles)
    unsigned cams = 0;
    {
        const int tiles_x = (p.lw[0] + 31) / 32, tiles_y = (p.lh[0] + 7) / 8;
        const int xs = p.ox0 + (int)blockIdx.x * 128 - p.rx, ys = p.oy0 + (int)blockIdx.y * 8 - p.ry;
        const int tx0 = max(xs, 0) >> 5, tx1 = min((xs + 127) >> 5, tiles_x - 1);
        const int ty0 = max(ys, 0) >> 3, ty1 = min((ys + 7) >> 3, tiles_y - 1);
        if (xs + 127 >= 0 && ys + 7 >= 0 && tiles_x > 0 && tiles_y > 0)
            for (int ty = ty0; ty <= ty1; ty++)
                for (int tx = tx0; tx <= tx1; tx++) cams |= __ldg(p.tile_cams + p.off_t[0] + (size_t)ty * tiles_x + tx);
    }
    if (X0 >= p.ox1 || Y >= p.oy1) return;
    int R[4], G[4], B[4];
    #pragma unroll
    for (int q = 0; q < 4; q++) R[q] = G[q] = B[q] = 0;
    const int x0 = X0 - p.rx, y = Y - p.ry;
    if (y >= 0 && y < p.rh && x0 + 3 >= 0 && x0 < p.rw) {
        // Interior groups (all four pixels inside the roi / the camera rectangle: almost all of them) run without per-pixel
        // bounds checks or branches: a zero weight contributes (short)(lap * 0) = 0 whether or not it is skipped.
        const float* dwrow = p.dstw + p.off_d[0] + (size_t)y * p.lw[0] + x0;
        const bool full4 = x0 >= 0 && x0 + 3 < p.rw;
        float dw[4];
        #pragma unroll
        for (int q = 0; q < 4; q++) dw[q] = (full4 || (x0 + q >= 0 && x0 + q < p.rw)) ? __ldg(dwrow + q) : 0.f;
        bool ok[4];
        #pragma unroll
        for (int q = 0; q < 4; q++) ok[q] = dw[q] > 1e-5f;             // dst_mask = dst_band_weights_[0] > WEIGHT_EPS (blenders.cpp:472)
        if (ok[0] || ok[1] || ok[2] || ok[3]) {
            int acc[4][3];
            #pragma unroll
            for (int q = 0; q < 4; q++) acc[q][0] = acc[q][1] = acc[q][2] = 0;
            for (; cams; cams &= cams - 1) {
                const int c = __ffs(cams) - 1;
                const MbCam& cam = p.cam[c];
                const int bw = cam.bw, cx0 = x0 - cam.x0, cy = y - cam.y0;
                if (cy < 0 || cy >= cam.bh || cx0 + 3 < 0 || cx0 >= bw) continue;
                const float* wrow = p.w + cam.off_w[0] + (size_t)cy * bw + cx0;
                const uint32_t* grow = p.g0 + cam.off_g[0] + (size_t)cy * bw + cx0;
                const bool in4 = cx0 >= 0 && cx0 + 3 < bw;
                float w[4];
                uint32_t px[4];
                #pragma unroll
                for (int q = 0; q < 4; q++) {
                    const bool in = in4 || (cx0 + q >= 0 && cx0 + q < bw);
                    w[q] = in ? __ldg(wrow + q) : 0.f;
                    px[q] = in ? __ldg(grow + q) : 0u;
                    w[q] = ok[q] ? w[q] : 0.f;
                }
                if (w[0] == 0.f && w[1] == 0.f && w[2] == 0.f && w[3] == 0.f) continue;
                uint2 up[4];
                if (p.nb > 0) pyrup4_packed(p.g + cam.off_g[1], bw >> 1, cam.bh >> 1, cx0, cy, up);
                #pragma unroll
                for (int q = 0; q < 4; q++)
                    lap_weight_acc((int)(px[q] & 255u), (int)((px[q] >> 8) & 255u), (int)((px[q] >> 16) & 255u), up[q], p.nb > 0, w[q], acc[q]);
            }
            int3 up[4];
            if (p.nb > 0) pyrup4_dst(p, 1, x0, y, up);
            #pragma unroll
            for (int q = 0; q < 4; q++) {
                const float den = __fadd_rn(dw[q], 1e-5f);            // normalizeUsingWeightMap (blenders.cpp:788-797)
                int3 v = make_int3((short)__float2int_rz(__fdiv_rn((float)(short)acc[q][0], den)),
                                   (short)__float2int_rz(__fdiv_rn((float)(short)acc[q][1], den)),
                                   (short)__float2int_rz(__fdiv_rn((float)(short)acc[q][2], den)));
                if (p.nb > 0) v = make_int3(sat16(up[q].x + v.x), sat16(up[q].y + v.y), sat16(up[q].z + v.z));
                R[q] = ok[q] ? clamp255(v.x) : 0; G[q] = ok[q] ? clamp255(v.y) : 0; B[q] = ok[q] ? clamp255(v.z) : 0;   // convertTo(CV_8U), masked
            }
        }
    }
    const bool full = X0 + 3 < p.ox1;
    if (p.rgb_out) {
        uint8_t* o = p.rgb_out + (size_t)Y * p.rgb_pitch + 3 * X0;
        #pragma unroll
        for (int q = 0; q < 4; q++) if (X0 + q < p.ox1) { o[3 * q] = (uint8_t)R[q]; o[3 * q + 1] = (uint8_t)G[q]; o[3 * q + 2] = (uint8_t)B[q]; }
    }
    if (p.oy) {
        uint8_t* oyp = p.oy + (size_t)Y * p.oy_pitch + X0;
        const uint32_t l4 = rgb_luma(R[0], G[0], B[0]) | (rgb_luma(R[1], G[1], B[1]) << 8) | (rgb_luma(R[2], G[2], B[2]) << 16) | (rgb_luma(R[3], G[3], B[3]) << 24);
        if (full && (((uintptr_t)oyp) & 3) == 0) *reinterpret_cast<uint32_t*>(oyp) = l4;
        else for (int q = 0; q < 4 && X0 + q < p.ox1; q++) oyp[q] = (uint8_t)(l4 >> (8 * q));
        if ((Y & 1) == 0) {                                            // X0 is even: chroma from pixels 0 and 2 of the group
            #pragma unroll
            for (int q = 0; q < 4; q += 2) {
                if (X0 + q >= p.ox1) break;
                const size_t co = (size_t)((X0 + q) >> 1) * p.uv_step;
                p.ou[(size_t)(Y >> 1) * p.ou_pitch + co] = (uint8_t)rgb_cb(R[q], G[q], B[q]);
                p.ov[(size_t)(Y >> 1) * p.ov_pitch + co] = (uint8_t)rgb_cr(R[q], G[q], B[q]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
namespace {
inline int mirror(int p, int len)      // BORDER_REFLECT (fedcba|abcdefgh|hgfedcb)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p - 1 : 2 * len - 1 - p;
    return p;
}
template <class T> T* upload(const std::vector<T>& v)
{
    T* d = nullptr;
    OB_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) OB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}
uint2 mk_entry(int32_t sx, int32_t sy, int src_w, int src_h, bool valid)
{
    if (!valid) return make_uint2(0u, 0u);
    int ix = std::min(32767, std::max(-32768, sx >> 5)), iy = std::min(32767, std::max(-32768, sy >> 5));
    const uint32_t fx = (uint32_t)(sx & 31), fy = (uint32_t)(sy & 31);
    const bool x0 = ix >= 0 && ix < src_w, x1 = ix + 1 >= 0 && ix + 1 < src_w, y0 = iy >= 0 && iy < src_h, y1 = iy + 1 >= 0 && iy + 1 < src_h;
    const uint32_t taps = (uint32_t)(x0 && y0) | ((uint32_t)(x1 && y0) << 1) | ((uint32_t)(x0 && y1) << 2) | ((uint32_t)(x1 && y1) << 3);
    if (!taps) return make_uint2(0u, 0u);
    uint32_t flags = C_VALID;
    if (taps != 15u) flags |= C_BORDER | (taps << C_TAP_SHIFT);
    return make_uint2((uint32_t)(iy * src_w + ix), fx | (fy << 5) | flags);
}
}  // namespace


// ------------------------------------------------------------------------------------------------ device-side set-up
// The static tables of a multiband mapper packed by CUDA kernels (the host code below does the same and stays for mappers
// whose warp cannot be staged, and for OCTVR_PACK=host): fixed-point coordinates, staged-warp jobs / boxes / entries of every
// camera rectangle (BORDER_REFLECT baked in), the f32 weight pyramids of the seam masks (cv::pyrDown's float arithmetic in
// its exact association, prep.cpp pyrdown_f32), the summed band weights and the per-tile camera sets.
struct MbGeom { int top, left, width, height, ys, ye, ch, xs, xe, cw; };      // one camera rectangle and its window (multiband_create, phase 0)
namespace {
struct MbpCam {
    const float* map1; const float* map2; const uint8_t* mask; const uint8_t* seam; int2* sxy;
    int roi_w, roi_h, src_w, src_h;
    MbGeom g;
    float* wfull[MB_MAX_LEVELS];      // weight pyramid of the FULL rectangle (temporary)
};
struct MbpParams { MbpCam cam[MAX_CAMS]; int n, nb; float shift; };

__device__ __forceinline__ int mbp_mirror(int p, int len)      // BORDER_REFLECT (fedcba|abcdefgh|hgfedcb)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p - 1 : 2 * len - 1 - p;
    return p;
}
__device__ __forceinline__ int mbp_mirror101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}
__global__ void __launch_bounds__(256) k_mbp_quantise(const __grid_constant__ MbpParams p, int c)
{
    const MbpCam& k = p.cam[c];
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)k.roi_w * k.roi_h) return;
    const float fw = (float)(double)k.src_w, fh = (float)(double)k.src_h;
    float px = __fadd_rn(__fmul_rn(k.map1[i], fw), 0.f), py = __fadd_rn(__fmul_rn(k.map2[i], fh), 0.f);
    if (p.shift != 0.f) { px = __fsub_rn(px, p.shift); py = __fsub_rn(py, p.shift); }
    k.sxy[i] = make_int2(__float2int_rn(__fmul_rn(px, 32.f)), __float2int_rn(__fmul_rn(py, 32.f)));
}
// table entry of window pixel (x, y) of camera k: valid (mk_entry's C_VALID), integer tap position, fractions
__device__ __forceinline__ bool mbp_entry(const MbpCam& k, int x, int y, int& ix, int& iy, int& fx, int& fy)
{
    const int ly = mbp_mirror(y + k.g.ys - k.g.top, k.roi_h), lx = mbp_mirror(x + k.g.xs - k.g.left, k.roi_w);
    const size_t o = (size_t)ly * k.roi_w + lx;
    if (!k.mask[o]) return false;
    const int2 s = k.sxy[o];
    ix = min(32767, max(-32768, s.x >> 5)); iy = min(32767, max(-32768, s.y >> 5));
    fx = s.x & 31; fy = s.y & 31;
    const bool x0 = ix >= 0 && ix < k.src_w, x1 = ix + 1 >= 0 && ix + 1 < k.src_w, y0 = iy >= 0 && iy < k.src_h, y1 = iy + 1 >= 0 && iy + 1 < k.src_h;
    return (x0 || x1) && (y0 || y1);
}
// one CTA per 32 x 16 tile of camera c's window: {xmin, xmax, ymin, ymax} of the taps of its valid pixels (xmin > xmax: none);
// rows[0 / 1] = min / max source row over the camera's valid entries
__global__ void __launch_bounds__(128) k_mbp_boxes(const __grid_constant__ MbpParams p, int c, int tiles_x, int4* boxes, int* rows)
{
    __shared__ int s[4];
    const MbpCam& k = p.cam[c];
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    if (threadIdx.x == 0) { s[0] = INT_MAX; s[1] = INT_MIN; s[2] = INT_MAX; s[3] = INT_MIN; }
    __syncthreads();
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
    #pragma unroll
    for (int h = 0; h < 4; h++) {
        const int px = threadIdx.x + h * 128, x = tx * TILE_W + (px & (TILE_W - 1)), y = ty * TILE_H + px / TILE_W;
        int ix, iy, fx, fy;
        if (x >= k.g.cw || y >= k.g.ch || !mbp_entry(k, x, y, ix, iy, fx, fy)) continue;
        xmin = min(xmin, ix); xmax = max(xmax, ix + 1); ymin = min(ymin, iy); ymax = max(ymax, iy + 1);
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s[0], xmin); atomicMax(&s[1], xmax); atomicMin(&s[2], ymin); atomicMax(&s[3], ymax); }
    __syncthreads();
    if (threadIdx.x == 0) {
        boxes[blockIdx.x] = make_int4(s[0], s[1], s[2], s[3]);
        if (s[0] <= s[1]) { atomicMin(rows + 2 * c, s[2]); atomicMax(rows + 2 * c + 1, s[3]); }
    }
}
// the 512 four-byte entries of staged-warp job j (final order): stage offset | fy << 13 | fx << 18 | valid
__global__ void __launch_bounds__(128) k_mbp_entries(const __grid_constant__ MbpParams p, const MbWarpJob* jobs, uint32_t* entries)
{
    const MbWarpJob job = jobs[blockIdx.x];
    const MbpCam& k = p.cam[job.cam];
    #pragma unroll
    for (int h = 0; h < 4; h++) {
        const int px = threadIdx.x + h * 128, x = job.tx * TILE_W + (px & (TILE_W - 1)), y = job.ty * TILE_H + px / TILE_W;
        int ix, iy, fx, fy;
        uint32_t e = 0u;
        if (x < k.g.cw && y < k.g.ch && mbp_entry(k, x, y, ix, iy, fx, fy))
            e = (uint32_t)((iy - job.by0) * job.bw + (ix - job.bx0)) | ((uint32_t)fy << 13) | ((uint32_t)fx << 18) | MBW_VALID;
        entries[(size_t)blockIdx.x * TILE_PX + px] = e;
    }
}
// level-0 weight map of the full rectangle: seam mask / 255 inside the roi, 0 in the border (BORDER_CONSTANT), blenders.cpp:345-366
__global__ void __launch_bounds__(256) k_mbp_w0(const __grid_constant__ MbpParams p, int c)
{
    const MbpCam& k = p.cam[c];
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= k.g.width || y >= k.g.height) return;
    const int lx = x - k.g.left, ly = y - k.g.top;
    float w = 0.f;
    if (lx >= 0 && ly >= 0 && lx < k.roi_w && ly < k.roi_h) w = __fadd_rn(__fmul_rn((float)k.seam[(size_t)ly * k.roi_w + lx], (float)(1. / 255.)), 0.f);
    k.wfull[0][(size_t)y * k.g.width + x] = w;
}
// cv::pyrDown for 32FC1 (pyramids.cpp:849-964 incl. the SSE association of :143-185) -- prep.cpp pyrdown_f32
__device__ __forceinline__ float mbp_hrow(const float* __restrict__ s, int sw, int c)
{
    const float a = __fmul_rn(s[c], 6.f), b = __fmul_rn(__fadd_rn(s[mbp_mirror101(c - 1, sw)], s[mbp_mirror101(c + 1, sw)]), 4.f);
    return __fadd_rn(__fadd_rn(__fadd_rn(a, b), s[mbp_mirror101(c - 2, sw)]), s[mbp_mirror101(c + 2, sw)]);
}
__global__ void __launch_bounds__(256) k_mbp_pyrdown(const float* __restrict__ src, int sw, int sh, float* __restrict__ dst)
{
    const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const float r0 = mbp_hrow(src + (size_t)mbp_mirror101(2 * y - 2, sh) * sw, sw, 2 * x), r1 = mbp_hrow(src + (size_t)mbp_mirror101(2 * y - 1, sh) * sw, sw, 2 * x);
    const float r2 = mbp_hrow(src + (size_t)(2 * y) * sw, sw, 2 * x);
    const float r3 = mbp_hrow(src + (size_t)mbp_mirror101(2 * y + 1, sh) * sw, sw, 2 * x), r4 = mbp_hrow(src + (size_t)mbp_mirror101(2 * y + 2, sh) * sw, sw, 2 * x);
    float v;
    if (x < (dw & ~7)) {
        const float a = __fadd_rn(__fadd_rn(r0, r4), __fadd_rn(r2, r2)), b = __fadd_rn(__fadd_rn(r1, r3), r2);
        v = __fmul_rn(__fadd_rn(a, __fmul_rn(b, 4.f)), 1.f / 256);
    } else
        v = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2, 6.f), __fmul_rn(__fadd_rn(r1, r3), 4.f)), r0), r4), (float)(1. / 256));
    dst[(size_t)y * dw + x] = v;
}
// window (c0.., r0..) of a full-rectangle weight level -> the packed weight pool
__global__ void __launch_bounds__(256) k_mbp_crop(const float* __restrict__ src, int sw, int c0, int r0, float* __restrict__ dst, int dw, int dh)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= dw || y >= dh) return;
    dst[(size_t)y * dw + x] = src[(size_t)(r0 + y) * sw + c0 + x];
}
// dst_band_weights_[l] (cameras added in order, blenders.cpp:421) and the camera set of every 32 x 8 tile of level l
__global__ void __launch_bounds__(256) k_mbp_dstw(const __grid_constant__ MbParams p, int l, float* dstw, uint16_t* tile_cams)
{
    __shared__ unsigned s_cams;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (threadIdx.x == 0 && threadIdx.y == 0) s_cams = 0u;
    __syncthreads();
    const int lw = p.lw[l], lh = p.lh[l];
    unsigned cams = 0;
    if (x < lw && y < lh) {
        float sum = 0.f;
        for (int c = 0; c < p.n; c++) {
            const MbCam& cam = p.cam[c];
            const int cx = x - (cam.x0 >> l), cy = y - (cam.y0 >> l), wl = cam.bw >> l, hl = cam.bh >> l;
            if (cx < 0 || cy < 0 || cx >= wl || cy >= hl) continue;
            const float w = p.w[cam.off_w[l] + (size_t)cy * wl + cx];
            sum = __fadd_rn(sum, w);
            if (w != 0.f) cams |= 1u << c;
        }
        dstw[p.off_d[l] + (size_t)y * lw + x] = sum;
    }
    cams = __reduce_or_sync(0xffffffffu, cams);
    if (threadIdx.x == 0 && cams) atomicOr(&s_cams, cams);
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) tile_cams[p.off_t[l] + (size_t)blockIdx.y * ((lw + 31) / 32) + blockIdx.x] = (uint16_t)s_cams;
}
template <class T> struct MbpBuf {
    T* p = nullptr;
    MbpBuf() {}
    explicit MbpBuf(size_t n) { OB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); }
    MbpBuf(const T* h, size_t n) : MbpBuf(n) { if (n) OB_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice)); }
    ~MbpBuf() { cudaFree(p); }
    MbpBuf(const MbpBuf&) = delete;
    T* release() { T* q = p; p = nullptr; return q; }
};
}  // namespace

// false: not applicable (a tile's footprint does not fit the stage ...) -- the caller runs the host set-up instead
static bool multiband_pack_gpu(octvr_mapper& m, const octvr_template& t, const std::vector<Img<uint8_t>>& seams, const std::vector<MbGeom>& geo,
                               Multiband& mb, size_t dst_floats, size_t n_tiles, int64_t& table_bytes)
{
    InitTrace tr("multiband_pack_gpu");
    MbParams& p = mb.p;
    const int n = p.n, nb = p.nb;
    MbpParams q;
    memset(&q, 0, sizeof(q));
    q.n = n; q.nb = nb; q.shift = m.texel_shift;
    std::vector<std::unique_ptr<MbpBuf<float>>> d_m1(n), d_m2(n);
    std::vector<std::unique_ptr<MbpBuf<uint8_t>>> d_mask(n), d_seam(n);
    std::vector<std::unique_ptr<MbpBuf<int2>>> d_sxy(n);
    std::vector<std::unique_ptr<MbpBuf<float>>> d_wfull(n);
    size_t w_total = 0;
    for (int i = 0; i < n; i++) {
        const TInput& in = t.inputs[i];
        const size_t px = (size_t)in.roi.w * in.roi.h;
        d_m1[i].reset(new MbpBuf<float>(in.map1.d.data(), px)); d_m2[i].reset(new MbpBuf<float>(in.map2.d.data(), px));
        d_mask[i].reset(new MbpBuf<uint8_t>(in.mask.d.data(), px)); d_seam[i].reset(new MbpBuf<uint8_t>(seams[i].d.data(), px));
        d_sxy[i].reset(new MbpBuf<int2>(px));
        MbpCam& k = q.cam[i];
        k.map1 = d_m1[i]->p; k.map2 = d_m2[i]->p; k.mask = d_mask[i]->p; k.seam = d_seam[i]->p; k.sxy = d_sxy[i]->p;
        k.roi_w = in.roi.w; k.roi_h = in.roi.h; k.src_w = m.in_w[i]; k.src_h = m.in_h[i]; k.g = geo[i];
        // full-rectangle weight pyramid (temporary): levels packed one after the other
        size_t tot = 0; int w = geo[i].width, h = geo[i].height;
        for (int l = 0; l <= nb; l++) { tot += (size_t)w * h; w = (w + 1) / 2; h = (h + 1) / 2; }
        d_wfull[i].reset(new MbpBuf<float>(tot));
        size_t off = 0; w = geo[i].width; h = geo[i].height;
        for (int l = 0; l <= nb; l++) { k.wfull[l] = d_wfull[i]->p + off; off += (size_t)w * h; w = (w + 1) / 2; h = (h + 1) / 2; }
        for (int l = 0; l <= nb; l++) { p.cam[i].off_w[l] = w_total; w_total += (size_t)(geo[i].cw >> l) * (geo[i].ch >> l); }
    }
    tr.lap("upload maps + masks");
    for (int i = 0; i < n; i++) k_mbp_quantise<<<(unsigned)(((size_t)t.inputs[i].roi.w * t.inputs[i].roi.h + 255) / 256), 256>>>(q, i);
    OB_CUDA(cudaGetLastError());
    // ---- staged warp: per camera the tiles of its window, their tap boxes
    MbpBuf<int> d_rows(2 * MAX_CAMS);
    {
        std::vector<int> init(2 * MAX_CAMS);
        for (int i = 0; i < MAX_CAMS; i++) { init[2 * i] = INT_MAX; init[2 * i + 1] = INT_MIN; }
        OB_CUDA(cudaMemcpy(d_rows.p, init.data(), init.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    std::vector<std::vector<int4>> boxes(n);
    for (int i = 0; i < n; i++) {
        const int tiles_x = (geo[i].cw + TILE_W - 1) / TILE_W, tiles_y = (geo[i].ch + TILE_H - 1) / TILE_H;
        const size_t nt = (size_t)tiles_x * tiles_y;
        if (geo[i].ch <= 0 || nt == 0) continue;
        if (tiles_x > 65535 || tiles_y > 65535) return false;
        MbpBuf<int4> d_box(nt);
        k_mbp_boxes<<<(unsigned)nt, 128>>>(q, i, tiles_x, d_box.p, d_rows.p);
        OB_CUDA(cudaGetLastError());
        boxes[i].resize(nt);
        OB_CUDA(cudaMemcpy(boxes[i].data(), d_box.p, nt * sizeof(int4), cudaMemcpyDeviceToHost));
    }
    std::vector<int> rows(2 * MAX_CAMS);
    OB_CUDA(cudaMemcpy(rows.data(), d_rows.p, rows.size() * sizeof(int), cudaMemcpyDeviceToHost));
    tr.lap("quantise + tile boxes");
    // jobs in the host packer's order: cameras in order, tiles row by row; size classes, tensor-map slots
    std::vector<MbWarpJob> wjobs;
    std::map<uint64_t, int> tmap_index;
    std::vector<int> col_lo(n, INT_MAX), col_hi(n, INT_MIN);      // source columns the taps of each camera touch
    auto size_class = [](int v) { return v <= 64 ? (v + 7) / 8 * 8 : v <= 128 ? (v + 15) / 16 * 16 : (v + 31) / 32 * 32; };
    for (int i = 0; i < n; i++) {
        const int tiles_x = (geo[i].cw + TILE_W - 1) / TILE_W;
        size_t keys_of_cam = 0;
        for (size_t k = 0; k < boxes[i].size(); k++) {
            const int4 b = boxes[i][k];
            if (b.x > b.y) continue;
            col_lo[i] = std::min(col_lo[i], b.x); col_hi[i] = std::max(col_hi[i], b.y);
            MbWarpJob job;
            memset(&job, 0, sizeof(job));
            job.bx0 = (int)std::floor(b.x / 4.0) * 4; job.by0 = b.z;
            const int bw = size_class(b.y - job.bx0 + 1), bh = size_class(b.w - b.z + 1);
            if (bw > 256 || bh > 256 || (int64_t)bw * bh > MB_STAGE) return false;
            job.bw = (uint16_t)bw; job.bh = (uint16_t)bh; job.cam = (uint16_t)i; job.tx = (uint16_t)(k % tiles_x); job.ty = (uint16_t)(k / tiles_x);
            const uint64_t key = ((uint64_t)i << 32) | ((uint64_t)bw << 16) | (uint64_t)bh;
            auto it = tmap_index.find(key);
            if (it == tmap_index.end()) { it = tmap_index.emplace(key, (int)tmap_index.size()).first; keys_of_cam++; }
            if (keys_of_cam > 60000 || tmap_index.size() > 65535) return false;
            job.tmap = (uint16_t)it->second;
            wjobs.push_back(job);
        }
    }
    if (wjobs.empty()) return false;
    {   // small-box jobs first (stable), as the host packer orders them
        std::vector<MbWarpJob> j2;
        j2.reserve(wjobs.size());
        auto small = [&](const MbWarpJob& j) { return (int)j.bw * j.bh <= MB_STAGE_SMALL; };
        for (auto& j : wjobs) if (small(j)) j2.push_back(j);
        mb.n_wsmall = (unsigned)j2.size();
        for (auto& j : wjobs) if (!small(j)) j2.push_back(j);
        wjobs.swap(j2);
    }
    MbpBuf<MbWarpJob> d_jobs(wjobs.data(), wjobs.size());
    MbpBuf<uint32_t> d_entries(wjobs.size() * TILE_PX);
    k_mbp_entries<<<(unsigned)wjobs.size(), 128>>>(q, d_jobs.p, d_entries.p);
    OB_CUDA(cudaGetLastError());
    tr.lap("jobs + entries");
    // ---- weights: full-rectangle pyramids, windows cropped into the pool, summed band weights, tile camera sets
    MbpBuf<float> d_w(w_total), d_dstw(dst_floats);
    MbpBuf<uint16_t> d_tc(n_tiles);
    OB_CUDA(cudaMemset(d_dstw.p, 0, std::max<size_t>(dst_floats, 1) * sizeof(float)));
    OB_CUDA(cudaMemset(d_tc.p, 0, std::max<size_t>(n_tiles, 1) * sizeof(uint16_t)));
    for (int i = 0; i < n; i++) {
        const MbGeom& g = geo[i];
        k_mbp_w0<<<dim3((g.width + 31) / 32, (g.height + 7) / 8), dim3(32, 8)>>>(q, i);
        int w = g.width, h = g.height;
        for (int l = 1; l <= nb; l++) {
            k_mbp_pyrdown<<<dim3(((w + 1) / 2 + 31) / 32, ((h + 1) / 2 + 7) / 8), dim3(32, 8)>>>(q.cam[i].wfull[l - 1], w, h, q.cam[i].wfull[l]);
            w = (w + 1) / 2; h = (h + 1) / 2;
        }
        w = g.width;
        for (int l = 0; l <= nb; l++) {
            const int dw = g.cw >> l, dh = g.ch >> l;
            if (dw > 0 && dh > 0)
                k_mbp_crop<<<dim3((dw + 31) / 32, (dh + 7) / 8), dim3(32, 8)>>>(q.cam[i].wfull[l], w, g.xs >> l, g.ys >> l, d_w.p + p.cam[i].off_w[l], dw, dh);
            w = (w + 1) / 2;
        }
    }
    OB_CUDA(cudaGetLastError());
    p.w = d_w.p;
    for (int l = 0; l <= nb; l++)
        if (p.lw[l] > 0 && p.lh[l] > 0) k_mbp_dstw<<<dim3((p.lw[l] + 31) / 32, (p.lh[l] + 7) / 8), dim3(32, 8)>>>(p, l, d_dstw.p, d_tc.p);
    OB_CUDA(cudaGetLastError());
    OB_CUDA(cudaDeviceSynchronize());
    tr.lap("weight pyramids + band weights");
    // ---- commit
    std::vector<uint8_t> tmaps(tmap_index.size() * 128);
    for (auto& kv : tmap_index) {
        const int cam = (int)(kv.first >> 32), bw = (int)((kv.first >> 16) & 0xFFFF), bh = (int)(kv.first & 0xFFFF);
        encode_rgbx_tensor_map(tmaps.data() + (size_t)kv.second * 128, m.d_rgbx[cam], m.in_w[cam], m.in_h[cam], bw, bh);
    }
    MbpBuf<uint8_t> d_tm(tmaps.data(), tmaps.size());
    mb.d_wtmaps = d_tm.release();
    mb.d_wjobs = d_jobs.release(); mb.d_wentries = d_entries.release(); mb.n_wjobs = (unsigned)wjobs.size();
    mb.d_w = d_w.release(); mb.d_dstw = d_dstw.release(); mb.d_tile_cams = d_tc.release();
    p.wjobs = mb.d_wjobs; p.wentries = mb.d_wentries; p.wtmaps = mb.d_wtmaps;
    if (m.band_y0 != 0 || m.band_y1 != t.out_h || m.band_x0 != 0 || m.band_x1 != t.out_w)
        for (int i = 0; i < n; i++) {
            const int lo = rows[2 * i], hi = rows[2 * i + 1];
            if (lo > hi) m.src_row0[i] = m.src_row1[i] = 0;
            else { m.src_row0[i] = std::max(0, lo) & ~1; m.src_row1[i] = std::min(m.in_h[i], hi + 1); }
        }
    for (int i = 0; i < n; i++) {
        if (col_lo[i] > col_hi[i]) { m.src_col0[i] = m.src_col1[i] = 0; continue; }
        m.src_col0[i] = std::max(0, col_lo[i]) & ~7; m.src_col1[i] = std::min(m.in_w[i], col_hi[i] + 1);
    }
    table_bytes = (int64_t)(wjobs.size() * TILE_PX * 4 + wjobs.size() * sizeof(MbWarpJob) + w_total * 4 + dst_floats * 4);
    return true;
}

Multiband* multiband_create(octvr_mapper& m, const octvr_template& t,
                            const std::function<void(std::vector<Img<int32_t>>&, std::vector<Img<int32_t>>&)>& host_xy)
{
    const int n = m.n;
    InitTrace tr("multiband_create");
    std::vector<Img<int32_t>> sx, sy;                       // fixed-point coordinates on the host: only the host set-up needs them
    std::vector<Img<uint8_t>> seams = t.seam_masks;
    bool have = (int)seams.size() == n;
    for (auto& s : seams) have = have && !s.empty();
    if (!have) seams = distance_seam_masks(t.inputs, t.out_w, m.device);     // Mapper requires mt.seam_masks (mapper.cpp:106)

    std::unique_ptr<Multiband> mb(new Multiband);
    MbParams& p = mb->p;
    memset(&p, 0, sizeof(p));
    const int req = int(std::ceil(std::log((double)m.blend) / std::log(2.)) - 1.);    // mapper.cpp:161
    OB_CHECK(req >= 1, "multiband needs blend >= 3 (at least one band; blenders.cpp:595)");
    Rect Rf = t.inputs[0].roi;
    for (int i = 1; i < n; i++) Rf = rect_union(Rf, t.inputs[i].roi);
    const double max_len = (double)std::max(Rf.w, Rf.h);
    const int nb = std::min(req, (int)std::ceil(std::log(max_len) / std::log(2.0)));  // blenders.cpp:242-243
    OB_CHECK(nb >= 0 && nb + 1 <= MB_MAX_LEVELS, "too many bands");
    const int al = 1 << nb;
    const int PW = Rf.w + (al - Rf.w % al) % al, PH = Rf.h + (al - Rf.h % al) % al;   // blenders.cpp:246-247
    p.n = n; p.nb = nb;
    // Row window [E0, E1) of the padded dst roi this mapper computes.  A row-band mapper (multi-GPU partition of one frame,
    // SURVEY.md 8e) keeps its band plus a halo and crops every camera rectangle, weight pyramid and dst level to it.  Cutting
    // the pyramids at a window edge corrupts, per level, at most 2 rows of the Gaussian levels next to the cut, 6 rows of the
    // Laplacian levels (pyrUp reads one coarser row either side) and r_l = 2 r_{l+1} + 2 rows of the collapsed levels, i.e.
    // 2^(bands+2) - 2 rows at level 0; a halo of 4 * 2^bands rows therefore leaves the band itself bit-identical to the
    // full-frame result.  Weights are computed on the full rectangles and cropped, so they are exact everywhere.
    // A column-band mapper does the same with a column window [F0, F1): for frames wider than tall (C4's 7680 x 1920 eyes)
    // the halos of column bands are a much smaller share of the work than those of row bands.
    int E0 = 0, E1 = PH, F0 = 0, F1 = PW;
    {
        const int halo = 4 * al;
        auto fl = [&](int v) { return v >= 0 ? v / al * al : -((-v + al - 1) / al * al); };
        if (m.band_y0 != 0 || m.band_y1 != t.out_h) {
            E0 = std::min(PH, std::max(0, fl(m.band_y0 - Rf.y - halo)));
            E1 = std::min(PH, std::max(0, fl(m.band_y1 - Rf.y + halo + al - 1)));
        }
        if (m.band_x0 != 0 || m.band_x1 != t.out_w) {
            F0 = std::min(PW, std::max(0, fl(m.band_x0 - Rf.x - halo)));
            F1 = std::min(PW, std::max(0, fl(m.band_x1 - Rf.x + halo + al - 1)));
        }
        if (E1 <= E0 || F1 <= F0) { E0 = E1 = 0; F0 = F1 = 0; }
    }
    const int WH = E1 - E0, WW = F1 - F0;
    p.rx = Rf.x + F0; p.ry = Rf.y + E0; p.rw = std::max(0, std::min(Rf.w, F1) - F0); p.rh = std::max(0, std::min(Rf.h, E1) - E0); p.out_w = t.out_w; p.out_h = t.out_h;
    p.oy0 = m.band_y0; p.oy1 = m.band_y1; p.ox0 = m.band_x0; p.ox1 = m.band_x1;
    size_t doff = 0;
    for (int l = 0; l <= nb; l++) {
        p.lw[l] = l == 0 ? WW : (p.lw[l - 1] + 1) / 2; p.lh[l] = l == 0 ? WH : (p.lh[l - 1] + 1) / 2;
        p.off_d[l] = doff; doff += (size_t)p.lw[l] * p.lh[l];
    }
    std::vector<float> dstw(doff, 0.f);
    std::vector<uint2> chunks;                              // k_mb_warp work list
    std::vector<uint16_t> tile_cams;                        // k_mb_band camera masks
    {
        size_t toff = 0;
        for (int l = 0; l <= nb; l++) { p.off_t[l] = toff; toff += (size_t)((p.lw[l] + 31) / 32) * ((p.lh[l] + 7) / 8); }
        tile_cams.assign(toff, 0);
    }
    std::vector<uint2> coords;
    std::vector<float> wts;
    size_t g0_total = 0, g_total = 0;
    // staged warp tables (k_mb_warp_staged); given up for the whole mapper if a tile's footprint does not fit the stage or a
    // source width is not a multiple of 4 pixels (TMA: 16-byte row pitch), or when OCTVR_MB_WARP=direct
    bool staged = true;
    for (int i = 0; i < n; i++) staged = staged && m.in_w[i] % 4 == 0;
    if (const char* e = getenv("OCTVR_MB_WARP")) staged = staged && std::string(e) != "direct";
    std::vector<MbWarpJob> wjobs;
    std::vector<uint32_t> wentries;
    std::map<uint64_t, int> tmap_index;                     // (camera, box w, box h) -> descriptor slot

    // ---- phase 0 (sequential, cheap): geometry of every camera rectangle, offsets into the shared arrays ----
    struct CamBuild {
        int top = 0, left = 0, width = 0, height = 0, ys = 0, ye = 0, ch = 0, xs = 0, xe = 0, cw = 0;
        std::vector<MbWarpJob> jobs; std::vector<uint32_t> entries; std::vector<uint64_t> keys;   // job.tmap = index into keys
        std::vector<uint2> chunks;
        std::vector<Img<float>> wl;                         // weight pyramid of the full rectangle, levels 0 .. nb
        bool staged_ok = true;
        int src_lo = INT32_MAX, src_hi = INT32_MIN;
    };
    std::vector<CamBuild> cb(n);
    for (int i = 0; i < n; i++) {
        const TInput& in = t.inputs[i];
        MbCam& c = p.cam[i];
        // MultiBandBlender::feed geometry (blenders.cpp:306-341); dst_roi_ = (Rf.x, Rf.y, PW, PH)
        const int gap = 3 * al;
        int tnx = std::max(Rf.x, in.roi.x - gap), tny = std::max(Rf.y, in.roi.y - gap);
        int bnx = std::min(Rf.x + PW, in.roi.x + in.roi.w + gap), bny = std::min(Rf.y + PH, in.roi.y + in.roi.h + gap);
        tnx = Rf.x + (((tnx - Rf.x) >> nb) << nb); tny = Rf.y + (((tny - Rf.y) >> nb) << nb);
        int width = bnx - tnx, height = bny - tny;
        width += (al - width % al) % al; height += (al - height % al) % al;
        bnx = tnx + width; bny = tny + height;
        const int dy = std::max(bny - (Rf.y + PH), 0), dx = std::max(bnx - (Rf.x + PW), 0);
        tnx -= dx; bnx -= dx; tny -= dy; bny -= dy;
        CamBuild& b = cb[i];
        b.top = in.roi.y - tny; b.left = in.roi.x - tnx; b.width = width; b.height = height;
        // rows [ys, ye) of the full rectangle fall inside the row window (all of it without row bands)
        // and columns [xs, xe) inside the column window
        const int fy0 = tny - Rf.y, fx0 = tnx - Rf.x;
        b.ys = std::min(height, std::max(0, E0 - fy0)); b.ye = std::max(b.ys, std::min(height, E1 - fy0));
        b.xs = std::min(width, std::max(0, F0 - fx0)); b.xe = std::max(b.xs, std::min(width, F1 - fx0));
        if (b.xe == b.xs) { b.ye = b.ys; b.xs = 0; b.xe = width; }      // nothing of this camera in the window: same state as an empty row range
        b.ch = b.ye - b.ys; b.cw = b.xe - b.xs;
        c.x0 = fx0 + b.xs - F0; c.y0 = fy0 + b.ys - E0; c.bw = b.cw; c.bh = b.ch;
        if (b.ch > 0) { mb->max_bw = std::max(mb->max_bw, b.cw); mb->max_bh = std::max(mb->max_bh, b.ch); }
        c.off_g[0] = g0_total; g0_total += (size_t)b.cw * b.ch;
        // every level starts at an even element: 16-byte aligned rows for the vector loads of mb_down_strip_p16
        for (int l = 1; l <= nb; l++) { c.off_g[l] = g_total; g_total += (size_t)(b.cw >> l) * (b.ch >> l); g_total += g_total & 1; }
    }
    std::vector<MbGeom> geo(n);
    for (int i = 0; i < n; i++) geo[i] = MbGeom{ cb[i].top, cb[i].left, cb[i].width, cb[i].height, cb[i].ys, cb[i].ye, cb[i].ch, cb[i].xs, cb[i].xe, cb[i].cw };
    // ---- the tables packed by CUDA kernels (default); the host set-up below when that does not apply ----
    bool on_device = false;
    {
        const char* e = getenv("OCTVR_PACK");
        if (staged && !(e && std::string(e) == "host")) {
            int64_t tb = 0;
            on_device = multiband_pack_gpu(m, t, seams, geo, *mb, doff, tile_cams.size(), tb);
            if (on_device) m.table_bytes = tb;
            else {      // partial state of a failed attempt
                cudaFree(mb->d_wjobs); cudaFree(mb->d_wentries); cudaFree(mb->d_wtmaps); cudaFree(mb->d_w); cudaFree(mb->d_dstw); cudaFree(mb->d_tile_cams);
                mb->d_wjobs = nullptr; mb->d_wentries = nullptr; mb->d_wtmaps = nullptr; mb->d_w = nullptr; mb->d_dstw = nullptr; mb->d_tile_cams = nullptr;
                mb->n_wjobs = mb->n_wsmall = 0; p.w = nullptr; p.wjobs = nullptr; p.wentries = nullptr; p.wtmaps = nullptr;
            }
        }
    }
    tr.lap(on_device ? "tables packed on the device" : "device packing not applicable");
    if (!on_device) {
    host_xy(sx, sy);
    coords.resize(g0_total);

    // ---- phase 1 (one host thread per camera): level-0 remap table with BORDER_REFLECT baked in, the f32 weight map
    //      (BORDER_CONSTANT 0) and its Gaussian pyramid, the staged warp jobs of the camera ----
    const bool want_staged = staged;
    auto build_cam = [&](int i) {
        const TInput& in = t.inputs[i];
        CamBuild& b = cb[i];
        const int width = b.width, height = b.height, ys = b.ys, ye = b.ye, ch = b.ch, top = b.top, left = b.left, xs = b.xs, xe = b.xe, cw = b.cw;
        Img<float> wmap(width, height, 0.f);
        const float inv255 = (float)(1. / 255.);
        uint2* ce = coords.data() + p.cam[i].off_g[0];
        std::vector<int16_t> qx, qy;                            // integer tap position of every valid pixel (staged warp)
        if (want_staged) { qx.assign((size_t)cw * ch, 0); qy.assign((size_t)cw * ch, 0); }
        for (int y = 0; y < height; y++) {
            const int ly = mirror(y - top, in.roi.h);
            const bool in_y = y - top >= 0 && y - top < in.roi.h;
            const bool in_rows = y >= ys && y < ye;
            if (!in_rows && !in_y) continue;
            for (int x = 0; x < width; x++) {
                const int lx = mirror(x - left, in.roi.w);
                if (in_rows && x >= xs && x < xe) {
                    const int32_t fsx = sx[i].row(ly)[lx], fsy = sy[i].row(ly)[lx];
                    const size_t at = (size_t)(y - ys) * cw + (x - xs);
                    ce[at] = mk_entry(fsx, fsy, m.in_w[i], m.in_h[i], in.mask.row(ly)[lx] != 0);
                    if (ce[at].y & C_VALID) { b.src_lo = std::min(b.src_lo, fsy >> 5); b.src_hi = std::max(b.src_hi, (fsy >> 5) + 1); }
                    if (want_staged) { qx[at] = (int16_t)std::min(32767, std::max(-32768, fsx >> 5)); qy[at] = (int16_t)std::min(32767, std::max(-32768, fsy >> 5)); }
                }
                if (in_y && x - left >= 0 && x - left < in.roi.w) wmap.row(y)[x] = seams[i].row(y - top)[x - left] * inv255 + 0.f;
            }
        }
        // staged warp: one job per 32 x 16 tile of the rectangle that has a valid pixel
        std::map<uint64_t, int> local_keys;
        for (int ty = 0; want_staged && b.staged_ok && ty < (ch + TILE_H - 1) / TILE_H; ty++)
            for (int tx = 0; b.staged_ok && tx < (cw + TILE_W - 1) / TILE_W; tx++) {
                int xmin = INT32_MAX, xmax = INT32_MIN, ymin = INT32_MAX, ymax = INT32_MIN;
                for (int py = ty * TILE_H; py < std::min(ch, (ty + 1) * TILE_H); py++)
                    for (int px = tx * TILE_W; px < std::min(cw, (tx + 1) * TILE_W); px++) {
                        const size_t at = (size_t)py * cw + px;
                        if (!(ce[at].y & C_VALID)) continue;
                        xmin = std::min(xmin, (int)qx[at]); xmax = std::max(xmax, qx[at] + 1); ymin = std::min(ymin, (int)qy[at]); ymax = std::max(ymax, qy[at] + 1);
                    }
                if (xmin > xmax) continue;                          // nothing valid: the tile keeps the zeros of the one-time memset
                auto size_class = [](int v) { return v <= 64 ? (v + 7) / 8 * 8 : v <= 128 ? (v + 15) / 16 * 16 : (v + 31) / 32 * 32; };
                MbWarpJob job;
                memset(&job, 0, sizeof(job));
                job.bx0 = (int)std::floor(xmin / 4.0) * 4; job.by0 = ymin;          // TMA: the innermost start coordinate must be 16-byte aligned
                const int bw = size_class(xmax - job.bx0 + 1), bh = size_class(ymax - ymin + 1);
                if (bw > 256 || bh > 256 || (int64_t)bw * bh > MB_STAGE || tx > 65535 || ty > 65535) { b.staged_ok = false; break; }
                job.bw = (uint16_t)bw; job.bh = (uint16_t)bh; job.cam = (uint16_t)i; job.tx = (uint16_t)tx; job.ty = (uint16_t)ty;
                const uint64_t key = ((uint64_t)i << 32) | ((uint64_t)bw << 16) | (uint64_t)bh;
                auto it = local_keys.find(key);
                if (it == local_keys.end()) { it = local_keys.emplace(key, (int)b.keys.size()).first; b.keys.push_back(key); }
                job.tmap = (uint16_t)it->second;                    // local index; renumbered when the cameras are merged
                if (b.keys.size() > 60000) { b.staged_ok = false; break; }
                const size_t e0 = b.entries.size();
                b.entries.resize(e0 + TILE_PX, 0u);
                for (int py = ty * TILE_H; py < std::min(ch, (ty + 1) * TILE_H); py++)
                    for (int px = tx * TILE_W; px < std::min(cw, (tx + 1) * TILE_W); px++) {
                        const size_t at = (size_t)py * cw + px;
                        if (!(ce[at].y & C_VALID)) continue;
                        const uint32_t off = (uint32_t)((qy[at] - job.by0) * bw + (qx[at] - job.bx0));
                        b.entries[e0 + (size_t)(py - ty * TILE_H) * TILE_W + (px - tx * TILE_W)] =
                            off | ((ce[at].y >> 5) & 31u) << 13 | (ce[at].y & 31u) << 18 | MBW_VALID;      // fy, fx of the table entry
                    }
                b.jobs.push_back(job);
            }
        for (size_t k = 0; k < ((size_t)cw * ch + 255) / 256; k++) {
            bool any = false;
            for (size_t e = k * 256; e < std::min((k + 1) * 256, (size_t)cw * ch) && !any; e++) any = (ce[e].y & C_VALID) != 0;
            if (any) b.chunks.push_back(make_uint2((uint32_t)i, (uint32_t)k));
        }
        while (b.chunks.size() % MB_WARP_CHUNKS) b.chunks.push_back(make_uint2((uint32_t)i, 0xFFFFFFFFu));   // a CTA's chunks share a camera
        b.wl.resize(nb + 1);
        b.wl[0] = std::move(wmap);
        for (int l = 1; l <= nb; l++) b.wl[l] = pyrdown_f32(b.wl[l - 1]);
    };
    {
        std::vector<std::thread> th;
        std::vector<std::exception_ptr> errs(n);
        for (int i = 0; i < n; i++)
            th.emplace_back([&, i] { try { build_cam(i); } catch (...) { errs[i] = std::current_exception(); } });
        for (auto& q : th) q.join();
        for (auto& e : errs) if (e) std::rethrow_exception(e);
    }

    // ---- phase 2 (sequential): merge in camera order ----
    for (int i = 0; i < n; i++) staged = staged && cb[i].staged_ok;
    for (int i = 0; i < n; i++) {
        CamBuild& b = cb[i];
        MbCam& c = p.cam[i];
        if (m.band_y0 != 0 || m.band_y1 != t.out_h || m.band_x0 != 0 || m.band_x1 != t.out_w) {   // band mapper: only these source rows need converting
            if (b.src_lo > b.src_hi) m.src_row0[i] = m.src_row1[i] = 0;
            else { m.src_row0[i] = std::max(0, b.src_lo) & ~1; m.src_row1[i] = std::min(m.in_h[i], b.src_hi + 1); }
        }
        if (staged) {
            std::vector<int> slot(b.keys.size());
            for (size_t k = 0; k < b.keys.size(); k++) {
                auto it = tmap_index.find(b.keys[k]);
                if (it == tmap_index.end()) it = tmap_index.emplace(b.keys[k], (int)tmap_index.size()).first;
                slot[k] = it->second;
            }
            if (tmap_index.size() > 65535) staged = false;
            for (MbWarpJob& j : b.jobs) j.tmap = (uint16_t)slot[j.tmap];
            wjobs.insert(wjobs.end(), b.jobs.begin(), b.jobs.end());
            wentries.insert(wentries.end(), b.entries.begin(), b.entries.end());
        }
        std::vector<MbWarpJob>().swap(b.jobs); std::vector<uint32_t>().swap(b.entries);
        chunks.insert(chunks.end(), b.chunks.begin(), b.chunks.end());
        for (int l = 0; l <= nb; l++) {
            c.off_w[l] = wts.size();
            const int r0 = b.ys >> l, r1 = b.ye >> l, c0 = b.xs >> l, c1 = b.xe >> l;      // the window's part of the full level-l weight map
            for (int r = r0; r < r1; r++) wts.insert(wts.end(), b.wl[l].row(r) + c0, b.wl[l].row(r) + c1);
        }
    }
    if (!staged) { wjobs.clear(); wentries.clear(); tmap_index.clear(); }
    // dst_band_weights_[l](rc) += weight (blenders.cpp:421), cameras in order for every element; threads split the dst rows
    for (int l = 0; l <= nb; l++) {
        const int rows = p.lh[l];
        const int nt = std::max(1, std::min(16, rows / 64));
        std::vector<std::thread> th;
        for (int q = 0; q < nt; q++)
            th.emplace_back([&, q] {
                // whole 8-row groups per thread: tile_cams words are not shared between threads
                const int g0 = (rows + 7) / 8 * q / nt * 8, g1 = std::min(rows, (rows + 7) / 8 * (q + 1) / nt * 8);
                for (int i = 0; i < n; i++) {
                    const CamBuild& b = cb[i];
                    const MbCam& c = p.cam[i];
                    const int xt = c.x0 >> l, yt = c.y0 >> l, r0 = b.ys >> l, r1 = b.ye >> l, c0 = b.xs >> l, c1 = b.xe >> l;
                    const Img<float>& wl = b.wl[l];
                    for (int y = std::max(0, g0 - yt); y < std::min(r1 - r0, g1 - yt); y++) {
                        float* dr = dstw.data() + p.off_d[l] + (size_t)(yt + y) * p.lw[l] + xt;
                        const float* wr = wl.row(r0 + y) + c0;
                        uint16_t* tc = tile_cams.data() + p.off_t[l] + (size_t)((yt + y) / 8) * ((p.lw[l] + 31) / 32);
                        for (int x = 0; x < c1 - c0; x++) {
                            dr[x] += wr[x];
                            if (wr[x] != 0.f) tc[(xt + x) / 32] |= (uint16_t)(1u << i);
                        }
                    }
                }
            });
        for (auto& q : th) q.join();
    }
    cb.clear();
    if (staged && !wjobs.empty()) {
        std::vector<uint8_t> tmaps(tmap_index.size() * 128);
        for (auto& kv : tmap_index) {
            const int cam = (int)(kv.first >> 32), bw = (int)((kv.first >> 16) & 0xFFFF), bh = (int)(kv.first & 0xFFFF);
            encode_rgbx_tensor_map(tmaps.data() + (size_t)kv.second * 128, m.d_rgbx[cam], m.in_w[cam], m.in_h[cam], bw, bh);
        }
        void* dt = nullptr;
        OB_CUDA(cudaMalloc(&dt, tmaps.size()));
        OB_CUDA(cudaMemcpy(dt, tmaps.data(), tmaps.size(), cudaMemcpyHostToDevice));
        mb->d_wtmaps = dt;
        {   // small-box jobs first (stable), entries moved along
            std::vector<uint32_t> order(wjobs.size());
            for (size_t k = 0; k < order.size(); k++) order[k] = (uint32_t)k;
            auto small = [&](uint32_t k) { return (int)wjobs[k].bw * wjobs[k].bh <= MB_STAGE_SMALL; };
            std::stable_partition(order.begin(), order.end(), small);
            std::vector<MbWarpJob> j2(wjobs.size());
            std::vector<uint32_t> e2(wentries.size());
            unsigned ns = 0;
            for (size_t k = 0; k < order.size(); k++) {
                j2[k] = wjobs[order[k]];
                memcpy(e2.data() + k * TILE_PX, wentries.data() + (size_t)order[k] * TILE_PX, TILE_PX * sizeof(uint32_t));
                ns += small(order[k]) ? 1u : 0u;
            }
            wjobs.swap(j2); wentries.swap(e2);
            mb->n_wsmall = ns;
        }
        mb->d_wjobs = upload(wjobs); mb->d_wentries = upload(wentries); mb->n_wjobs = (unsigned)wjobs.size();
        p.wjobs = mb->d_wjobs; p.wentries = mb->d_wentries; p.wtmaps = mb->d_wtmaps;
        coords.clear(); coords.shrink_to_fit();             // the 8-byte table is only read by the direct kernel
        chunks.clear();
    }
    mb->d_coords = upload(coords);
    mb->d_w = upload(wts);
    mb->d_dstw = upload(dstw);
    mb->d_chunks = upload(chunks); mb->n_chunks = (unsigned)chunks.size();
    mb->d_tile_cams = upload(tile_cams);
    m.table_bytes = (int64_t)(coords.size() * sizeof(uint2) + wentries.size() * 4 + wjobs.size() * sizeof(MbWarpJob) + wts.size() * 4 + dstw.size() * 4);
    tr.lap("tables packed on the host");
    }   // !on_device
    p.warp_chunks = mb->d_chunks; p.tile_cams = mb->d_tile_cams;
    OB_CUDA(cudaMalloc(&mb->d_g0, std::max<size_t>(g0_total, 1) * sizeof(uint32_t)));
    OB_CUDA(cudaMemset(mb->d_g0, 0, std::max<size_t>(g0_total, 1) * sizeof(uint32_t)));       // chunks without valid entries stay 0
    OB_CUDA(cudaMalloc(&mb->d_g, std::max<size_t>(g_total, 1) * sizeof(short4)));
    OB_CUDA(cudaMalloc(&mb->d_dst, std::max<size_t>(doff, 1) * sizeof(short4)));
    OB_CUDA(cudaMalloc(&mb->d_wide, MB_MAX_LEVELS * sizeof(int)));
    {   // OCTVR_MB_WIDE=1 (diagnostic / tests): every blended level is treated as wide, i.e. the 32-bit pyrUp everywhere
        const char* e = getenv("OCTVR_MB_WIDE");
        mb->force_wide = e && atoi(e) != 0;
        std::vector<int> init(MB_MAX_LEVELS, mb->force_wide ? 1 : 0);
        OB_CUDA(cudaMemcpy(mb->d_wide, init.data(), init.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    p.wide = mb->d_wide; p.clear_wide = mb->force_wide ? 0 : 1;
    {   // side stream for the small-level chain (OCTVR_MB_FORK=0: everything on the caller's stream)
        const char* e = getenv("OCTVR_MB_FORK");
        mb->fork = !(e && atoi(e) == 0);
        if (mb->fork) {
            int lo = 0, hi = 0;
            OB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            OB_CUDA(cudaStreamCreateWithPriority(&mb->side, cudaStreamNonBlocking, hi));
            OB_CUDA(cudaEventCreateWithFlags(&mb->ev_fork, cudaEventDisableTiming));
            OB_CUDA(cudaEventCreateWithFlags(&mb->ev_join, cudaEventDisableTiming));
        }
    }
    p.coords = mb->d_coords; p.g0 = mb->d_g0; p.g = mb->d_g; p.w = mb->d_w; p.dst = mb->d_dst; p.dstw = mb->d_dstw;
    for (int i = 0; i < n; i++) { p.rgbx[i] = m.d_rgbx[i]; p.src_pitch[i] = m.in_w[i]; }
    mb->launches = (mb->n_wjobs ? (mb->n_wsmall ? 1 : 0) + (mb->n_wjobs > mb->n_wsmall ? 1 : 0) : 1) + nb + (nb >= 1 ? 1 : 0) + (mb->fork && nb >= 3 ? 1 : 0) + std::max(0, nb - 1) + 1;
    return mb.release();
}

void multiband_stitch(octvr_mapper& m, const octvr_frame* out, cudaStream_t s)
{
    Multiband& mb = *m.mb;
    MbParams p = mb.p;
    p.gain_f32 = m.d_gain_f32; p.gain_flag = m.d_gain_flag; p.gain_lut = m.d_gain_lut; p.use_gain = m.gain ? 1 : 0;
    if (out) {
        p.oy = out->y; p.ou = out->u; p.ov = out->v;
        p.oy_pitch = (uint32_t)out->y_pitch; p.ou_pitch = (uint32_t)out->u_pitch; p.ov_pitch = (uint32_t)out->v_pitch;
        p.uv_step = out->uv_pixel_stride;
    }
    p.rgb_out = m.rgb_this_frame ? m.d_rgb : nullptr; p.rgb_pitch = (uint32_t)m.out_w * 3;
    const int nb = p.nb, n = p.n;
    if (mb.n_wjobs) {
        if (mb.n_wsmall) k_mb_warp_staged<128, MB_STAGE_SMALL><<<mb.n_wsmall, 128, 0, s>>>(p, 0u);
        if (mb.n_wjobs > mb.n_wsmall) k_mb_warp_staged<256, MB_STAGE><<<mb.n_wjobs - mb.n_wsmall, 256, 0, s>>>(p, mb.n_wsmall);
    } else if (mb.n_chunks) k_mb_warp<<<mb.n_chunks / MB_WARP_CHUNKS, 256, 0, s>>>(p);
    if (p.lh[0] > 0 && p.lw[0] > 0 && mb.max_bh > 0) {     // an empty window (a band outside the result roi) only writes black
        auto down = [&](int l, cudaStream_t st) {
            const dim3 grid(((mb.max_bw >> (l + 1)) + 31) / 32, ((mb.max_bh >> (l + 1)) + 31) / 32, n);
            if (l == 0) k_mb_down<true><<<grid, dim3(32, 8), 0, st>>>(p, l);
            else k_mb_down<false><<<grid, dim3(32, 8), 0, st>>>(p, l);
        };
        auto band = [&](int l0, int l1, cudaStream_t st) {      // destination levels l0 .. l1 in one launch (level 0 is computed inside k_mb_final)
            k_mb_band<<<dim3((p.lw[l0] + 127) / 128, (p.lh[l0] + 7) / 8, l1 - l0 + 1), dim3(32, 8), 0, st>>>(p, l0);
        };
        auto collapse = [&](int l, cudaStream_t st) {
            k_mb_collapse<<<dim3((p.lw[l - 1] + 127) / 128, (p.lh[l - 1] + 7) / 8), dim3(32, 8), 0, st>>>(p, l);
        };
        if (mb.fork && nb >= 3) {
            // The levels >= 2 are a chain of small, latency-bound launches (for C3: down 10 + 8 + 8 us, collapse 5 + 5 + 8 us).
            // They depend on level 1 only through G_2, and level 1's own band (the large one) does not depend on them: the chain
            // runs on a side stream next to it and joins before the last collapse step.
            down(0, s); down(1, s);
            cudaEventRecord(mb.ev_fork, s);
            cudaStreamWaitEvent(mb.side, mb.ev_fork, 0);
            for (int l = 2; l < nb; l++) down(l, mb.side);
            band(2, nb, mb.side);
            for (int l = nb; l >= 3; l--) collapse(l, mb.side);
            cudaEventRecord(mb.ev_join, mb.side);
            band(1, 1, s);
            cudaStreamWaitEvent(s, mb.ev_join, 0);
            collapse(2, s);
        } else {
            for (int l = 0; l < nb; l++) down(l, s);
            if (nb >= 1) band(1, nb, s);
            for (int l = nb; l >= 2; l--) collapse(l, s);
        }
    }
    k_mb_final<<<dim3((p.ox1 - p.ox0 + 127) / 128, (p.oy1 - p.oy0 + 7) / 8), dim3(32, 8), 0, s>>>(p);
}

int multiband_launches(const octvr_mapper& m) { return m.mb ? m.mb->launches : 0; }

void multiband_destroy(Multiband* mb)
{
    if (!mb) return;
    if (mb->side) cudaStreamDestroy(mb->side);
    if (mb->ev_fork) cudaEventDestroy(mb->ev_fork);
    if (mb->ev_join) cudaEventDestroy(mb->ev_join);
    cudaFree(mb->d_chunks); cudaFree(mb->d_tile_cams); cudaFree(mb->d_wjobs); cudaFree(mb->d_wentries); cudaFree(mb->d_wtmaps); cudaFree(mb->d_coords); cudaFree(mb->d_g0); cudaFree(mb->d_g); cudaFree(mb->d_w); cudaFree(mb->d_dst); cudaFree(mb->d_dstw); cudaFree(mb->d_wide);
    delete mb;
}

}  // namespace ob
