// csrc/fast.cu -- vr::FastMapper (modules/octvr/include/octvr.hpp:123-144, src/mapper_fast.cpp:27-195) on B200.
//
// The reference's NV12 path keeps the frame in 4:2:0 from end to end: per camera three cv::remap_weighted launches
// (imgproc/src/opencl/remap_weighted.cl:20-77: bilinear taps with 5-bit fractions, times a u8 feather weight, rounded
// and ADDED to a 16-bit accumulator plane), full-resolution tables for luma and half-resolution tables for the two chroma
// planes, then convertTo(CV_8U, 1 / 255).  3 n + 6 launches and three 16-bit accumulator planes through memory per frame.
//
// Here: ONE launch, accumulators in registers, two device paths with the same arithmetic:
//   k_fast_staged (default)  tiles of 16 rows x 16 threads; per (tile, camera) job the source footprint is copied into shared
//                            memory (cp.async, zero fill outside the plane) and the taps are read from there; 4-byte entries.
//   k_fast_nv12              tiles of 8 rows x 32 threads; taps gathered byte by byte through L1 / L2; 8-byte entries
//                              { int16 ax, int16 ay, u8 fx, u8 fy, u8 weight, u8 flags }
//                            laid out [job][thread][pixel], streamed with 16-byte loads.  Serves frames that are not 16-byte
//                            aligned and mappers with a footprint larger than the 16 KB stage (OCTVR_FAST=direct forces it).
// A luma thread owns 4 consecutive pixels, a chroma thread 2 consecutive chroma positions (both channels), so every thread
// ends with one 32-bit store.  A *job* is a (tile, camera) pair with at least one non-zero weight.
//
// Arithmetic (bit-exact with the kernel above on any IEEE device, see oracle/refgen/ref_fast.cpp):
//   V = a (32-fx)(32-fy) + b fx (32-fy) + c (32-fx) fy + d fx fy           (integer = 1024 x the float expression, exact)
//   q = rte( fl32(V * w) / 1024 )                                            (the kernel's single rounding, then sat_rte)
//   acc = (acc + q) mod 2^16;   out = rte( fl32(acc * fl32(1 / 255)) )       (cvtScale16u8u with a float scale, cvRound)
// Taps outside the source read 0 (the OUTSIDE() test of the kernel).  As in the reference, output chroma channel 0 is
// remapped from input chroma channel 1 and channel 1 from channel 0 (mapper_fast.cpp:179-180).
#include "prep.h"
#include "template.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

namespace ob {
namespace {

constexpr int FT_ROWS = 8, FT_THREADS = 256;
constexpr uint32_t F_INSIDE = 1u;                    // all four taps inside the source plane

struct FastCam { const uint8_t* y; const uint8_t* uv; int y_pitch, uv_pitch, w, h; };   // w, h: luma size; chroma plane: w / 2 x h / 2 pairs
struct FastParams {
    FastCam cam[MAX_CAMS];
    const uint32_t* tile_job_start;                  // [n_tiles + 1]
    const uint8_t* job_cam;
    const uint4* entries;                            // luma jobs: 2 x uint4 per thread, chroma jobs: 1 x uint4 per thread
    const uint32_t* job_entry_ofs;                   // per job: offset in uint4 units
    uint8_t* out; int out_pitch, W, H;
    int luma_tiles_x, luma_tiles, chroma_tiles_x;
};

__device__ __forceinline__ uint32_t fast_term(const uint8_t* __restrict__ p, int pitch, int w, int h, uint2 e)
{
    // e.x = ax | ay << 16 (int16 each), e.y = fx | fy << 8 | weight << 16 | flags << 24
    const uint32_t wgt = (e.y >> 16) & 0xFF;
    if (wgt == 0) return 0;
    const int ax = (short)(e.x & 0xFFFF), ay = (short)(e.x >> 16);
    const int fx = e.y & 0xFF, fy = (e.y >> 8) & 0xFF;
    int a, b, c, d;
    if ((e.y >> 24) & F_INSIDE) {
        const uint8_t* q = p + (size_t)ay * pitch + ax;
        a = q[0]; b = q[1]; c = q[pitch]; d = q[pitch + 1];
    } else {
        const bool x0 = ax >= 0 && ax < w, x1 = ax + 1 >= 0 && ax + 1 < w, y0 = ay >= 0 && ay < h, y1 = ay + 1 >= 0 && ay + 1 < h;
        a = (x0 && y0) ? p[(size_t)ay * pitch + ax] : 0;
        b = (x1 && y0) ? p[(size_t)ay * pitch + ax + 1] : 0;
        c = (x0 && y1) ? p[(size_t)(ay + 1) * pitch + ax] : 0;
        d = (x1 && y1) ? p[(size_t)(ay + 1) * pitch + ax + 1] : 0;
    }
    const int V = (a * (32 - fx) + b * fx) * (32 - fy) + (c * (32 - fx) + d * fx) * fy;
    return (uint32_t)__float2int_rn(__fmul_rn(__int2float_rn(V), (float)wgt) * 0.0009765625f);
}

// both chroma channels of one position: returns q(channel 0) | q(channel 1) << 16 of the INPUT plane
__device__ __forceinline__ uint32_t fast_term_uv(const uint8_t* __restrict__ p, int pitch, int w, int h, uint2 e)
{
    const uint32_t wgt = (e.y >> 16) & 0xFF;
    if (wgt == 0) return 0;
    const int ax = (short)(e.x & 0xFFFF), ay = (short)(e.x >> 16);
    const int fx = e.y & 0xFF, fy = (e.y >> 8) & 0xFF;
    uint32_t a, b, c, d;                             // channel 0 in bits 0-7, channel 1 in bits 8-15
    if ((e.y >> 24) & F_INSIDE) {
        const uint8_t* q = p + (size_t)ay * pitch + 2 * ax;
        a = *reinterpret_cast<const uint16_t*>(q); b = *reinterpret_cast<const uint16_t*>(q + 2);
        c = *reinterpret_cast<const uint16_t*>(q + pitch); d = *reinterpret_cast<const uint16_t*>(q + pitch + 2);
    } else {
        const bool x0 = ax >= 0 && ax < w, x1 = ax + 1 >= 0 && ax + 1 < w, y0 = ay >= 0 && ay < h, y1 = ay + 1 >= 0 && ay + 1 < h;
        a = (x0 && y0) ? *reinterpret_cast<const uint16_t*>(p + (size_t)ay * pitch + 2 * ax) : 0;
        b = (x1 && y0) ? *reinterpret_cast<const uint16_t*>(p + (size_t)ay * pitch + 2 * ax + 2) : 0;
        c = (x0 && y1) ? *reinterpret_cast<const uint16_t*>(p + (size_t)(ay + 1) * pitch + 2 * ax) : 0;
        d = (x1 && y1) ? *reinterpret_cast<const uint16_t*>(p + (size_t)(ay + 1) * pitch + 2 * ax + 2) : 0;
    }
    const float fw = (float)wgt;
    const int k00 = (32 - fx) * (32 - fy), k01 = fx * (32 - fy), k10 = (32 - fx) * fy, k11 = fx * fy;
    const int V0 = (int)(a & 0xFF) * k00 + (int)(b & 0xFF) * k01 + (int)(c & 0xFF) * k10 + (int)(d & 0xFF) * k11;
    const int V1 = (int)(a >> 8) * k00 + (int)(b >> 8) * k01 + (int)(c >> 8) * k10 + (int)(d >> 8) * k11;
    const uint32_t q0 = (uint32_t)__float2int_rn(__fmul_rn(__int2float_rn(V0), fw) * 0.0009765625f);
    const uint32_t q1 = (uint32_t)__float2int_rn(__fmul_rn(__int2float_rn(V1), fw) * 0.0009765625f);
    return q0 | (q1 << 16);
}

__device__ __forceinline__ uint32_t fast_out(uint32_t acc16)
{
    return (uint32_t)__float2int_rn(__fmul_rn(__int2float_rn((int)(acc16 & 0xFFFFu)), (float)(1.0 / 255.0)));   // <= 257
}

__global__ void __launch_bounds__(FT_THREADS) k_fast_nv12(const __grid_constant__ FastParams p)
{
    const int tile = blockIdx.x, tid = threadIdx.x, lane_x = tid & 31, row = tid >> 5;
    const uint32_t j0 = __ldg(p.tile_job_start + tile), j1 = __ldg(p.tile_job_start + tile + 1);
    if (tile < p.luma_tiles) {
        const int tx = tile % p.luma_tiles_x, ty = tile / p.luma_tiles_x;
        const int x = tx * 128 + lane_x * 4, y = ty * FT_ROWS + row;
        uint32_t acc[4] = { 0, 0, 0, 0 };
        for (uint32_t j = j0; j < j1; j++) {
            const FastCam& c = p.cam[__ldg(p.job_cam + j)];
            const uint4* e = p.entries + __ldg(p.job_entry_ofs + j) + tid * 2;
            const uint4 e0 = __ldg(e), e1 = __ldg(e + 1);
            acc[0] += fast_term(c.y, c.y_pitch, c.w, c.h, make_uint2(e0.x, e0.y));
            acc[1] += fast_term(c.y, c.y_pitch, c.w, c.h, make_uint2(e0.z, e0.w));
            acc[2] += fast_term(c.y, c.y_pitch, c.w, c.h, make_uint2(e1.x, e1.y));
            acc[3] += fast_term(c.y, c.y_pitch, c.w, c.h, make_uint2(e1.z, e1.w));
        }
        if (y >= p.H || x >= p.W) return;
        uint32_t o[4];
        #pragma unroll
        for (int k = 0; k < 4; k++) o[k] = min(fast_out(acc[k]), 255u);
        uint8_t* d = p.out + (size_t)y * p.out_pitch + x;
        if (x + 3 < p.W && ((p.out_pitch & 3) == 0)) *reinterpret_cast<uint32_t*>(d) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        else for (int k = 0; k < 4 && x + k < p.W; k++) d[k] = (uint8_t)o[k];
    } else {
        const int ct = tile - p.luma_tiles;
        const int tx = ct % p.chroma_tiles_x, ty = ct / p.chroma_tiles_x;
        const int cw = p.W / 2, ch = p.H / 2;
        const int x = tx * 64 + lane_x * 2, y = ty * FT_ROWS + row;
        uint32_t acc[2] = { 0, 0 };
        for (uint32_t j = j0; j < j1; j++) {
            const FastCam& c = p.cam[__ldg(p.job_cam + j)];
            const uint4 e = __ldg(p.entries + __ldg(p.job_entry_ofs + j) + tid);
            // 16-bit lanes wrap independently: add with the carry between the halves cut
            const uint32_t q0 = fast_term_uv(c.uv, c.uv_pitch, c.w / 2, c.h / 2, make_uint2(e.x, e.y));
            const uint32_t q1 = fast_term_uv(c.uv, c.uv_pitch, c.w / 2, c.h / 2, make_uint2(e.z, e.w));
            acc[0] = (((acc[0] & 0xFFFFu) + (q0 & 0xFFFFu)) & 0xFFFFu) | (((acc[0] >> 16) + (q0 >> 16)) << 16);
            acc[1] = (((acc[1] & 0xFFFFu) + (q1 & 0xFFFFu)) & 0xFFFFu) | (((acc[1] >> 16) + (q1 >> 16)) << 16);
        }
        if (y >= ch || x >= cw) return;
        // output channel 0 <- input channel 1 (high half), output channel 1 <- input channel 0 (mapper_fast.cpp:179-180)
        uint32_t o[4] = { min(fast_out(acc[0] >> 16), 255u), min(fast_out(acc[0]), 255u), min(fast_out(acc[1] >> 16), 255u), min(fast_out(acc[1]), 255u) };
        uint8_t* d = p.out + (size_t)(p.H + y) * p.out_pitch + 2 * x;
        if (x + 1 < cw && ((p.out_pitch & 3) == 0)) *reinterpret_cast<uint32_t*>(d) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        else { d[0] = (uint8_t)o[0]; d[1] = (uint8_t)o[1]; if (x + 1 < cw) { d[2] = (uint8_t)o[2]; d[3] = (uint8_t)o[3]; } }
    }
}


// ---- k_fast_staged (default): the same arithmetic with the taps read from shared memory ------------------------------------
// Tiles of 16 rows x 16 threads (luma: 64 x 16 pixels, a thread owns 4; chroma: 32 x 16 positions, a thread owns 2).  Per job
// the source footprint of the tile -- rows [by0, by0 + bh) x bytes [bx0, bx0 + 16 bwc) of the camera's plane, <= 16 KB -- is
// copied into one of two shared-memory stages with 16-byte cp.async (zero-filled where it leaves the plane: that IS the
// OUTSIDE() rule of the OpenCL kernel), the next job's copy in flight while this job's pixels are gathered.  Entries shrink
// to 4 bytes: stage offset (14 bits) | fx << 14 | fy << 19 | weight << 24.
constexpr int FS_ROWS = 16, FS_STAGE = 16384;
struct FsJob { short bx0, by0; unsigned short bwc, bh; uint32_t rcp; uint32_t cam; };      // rcp = floor(2^32 / bwc) + 1: c / bwc == __umulhi(c, rcp) for c < 2^16
static_assert(sizeof(FsJob) == 16, "FsJob");
struct FsParams {
    FastCam cam[MAX_CAMS];
    const uint32_t* tile_job_start; const FsJob* jobs; const uint32_t* entries;      // job j owns 1024 (luma) / 512 (chroma) entries, in job order
    uint8_t* out; int out_pitch, W, H;
    int luma_tiles_x, luma_tiles, chroma_tiles_x;
    uint32_t luma_jobs; int stage_bytes;
};
__device__ __forceinline__ void fs_issue(const FsJob& job, const uint8_t* __restrict__ plane, int pitch, int wbytes, int h, uint8_t* stage)
{
    const int n = (int)job.bwc * job.bh;
    for (int c = threadIdx.x; c < n; c += FT_THREADS) {
        const int r = (int)__umulhi((unsigned)c, job.rcp), k = c - r * job.bwc;
        const int gy = job.by0 + r, gx = job.bx0 + 16 * k;
        int valid = (gy >= 0 && gy < h && gx >= 0) ? min(max(wbytes - gx, 0), 16) : 0;
        const uint8_t* src = valid ? plane + (size_t)gy * pitch + gx : plane;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + c * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(valid) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
// q = rte(fl32(V * w) / 1024) without the XU pipe (I2F / F2I run at a quarter of the FP32 rate): V < 2^18 and w < 2^8 become
// floats by the 2^23 bias trick (exact), the product is rounded once (the OpenCL kernel's rounding), the scaling by 2^-10 is
// exact, and adding 1.5 * 2^23 rounds to nearest even into the low mantissa bits (one FFMA: x * 2^-10 is exact, so the fused
// rounding is the rounding of the sum).  Returns q + 0x4B400000; the caller subtracts the bias when it accumulates.
__device__ __forceinline__ float fs_float(uint32_t v) { return __fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.f); }
__device__ __forceinline__ uint32_t fs_q(int V, float fw)
{
    return __float_as_uint(__fmaf_rn(__fmul_rn(fs_float((uint32_t)V), fw), 0.0009765625f, 12582912.f)) - 0x4B400000u;
}
template <int MINB>
__global__ void __launch_bounds__(FT_THREADS, MINB) k_fast_staged(const __grid_constant__ FsParams p)
{
    extern __shared__ __align__(16) uint8_t s_dyn[];          // two stages of p.stage_bytes (the mapper's largest footprint)
    const int tile = blockIdx.x, tid = threadIdx.x, lane_x = tid & 15, row = tid >> 4;
    const uint32_t j0 = __ldg(p.tile_job_start + tile), j1 = __ldg(p.tile_job_start + tile + 1);
    const bool luma = tile < p.luma_tiles;
    uint32_t acc[4] = { 0, 0, 0, 0 };                  // luma: four pixels; chroma: {pos 0 ch 0, pos 0 ch 1, pos 1 ch 0, pos 1 ch 1}
    FsJob cur, nxt;
    if (j0 < j1) {
        cur = p.jobs[j0];
        const FastCam& c = p.cam[cur.cam];
        fs_issue(cur, luma ? c.y : c.uv, luma ? c.y_pitch : c.uv_pitch, luma ? c.w : 2 * (c.w / 2), luma ? c.h : c.h / 2, s_dyn);
    }
    for (uint32_t j = j0; j < j1; j++) {
        const int buf = (int)(j - j0) & 1;
        if (j + 1 < j1) {
            nxt = p.jobs[j + 1];
            const FastCam& c = p.cam[nxt.cam];
            fs_issue(nxt, luma ? c.y : c.uv, luma ? c.y_pitch : c.uv_pitch, luma ? c.w : 2 * (c.w / 2), luma ? c.h : c.h / 2, s_dyn + (buf ^ 1) * p.stage_bytes);
        }
        const uint8_t* st = s_dyn + buf * p.stage_bytes;
        const int sp = (int)cur.bwc * 16;
        if (luma) {
            const uint4 e4 = __ldg(reinterpret_cast<const uint4*>(p.entries + (size_t)j * 1024) + tid);
            if (j + 1 < j1) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const uint32_t e[4] = { e4.x, e4.y, e4.z, e4.w };
            #pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t wgt = e[k] >> 24;
                const int off = e[k] & 0x3FFF, fx = (e[k] >> 14) & 31, fy = (e[k] >> 19) & 31;
                const int a = st[off], b = st[off + 1], c2 = st[off + sp], d = st[off + sp + 1];
                const int V = (a * (32 - fx) + b * fx) * (32 - fy) + (c2 * (32 - fx) + d * fx) * fy;
                acc[k] += fs_q(V, fs_float(wgt));       // weight 0 (no contribution): q = 0 whatever the taps are
            }
        } else {
            const uint2 e2 = __ldg(reinterpret_cast<const uint2*>(p.entries + (size_t)p.luma_jobs * 1024 + (size_t)(j - p.luma_jobs) * 512) + tid);
            if (j + 1 < j1) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const uint32_t e[2] = { e2.x, e2.y };
            #pragma unroll
            for (int k = 0; k < 2; k++) {
                const float fw = fs_float(e[k] >> 24);
                const int off = e[k] & 0x3FFF, fx = (e[k] >> 14) & 31, fy = (e[k] >> 19) & 31;
                const uint32_t a = *reinterpret_cast<const uint16_t*>(st + off), b = *reinterpret_cast<const uint16_t*>(st + off + 2);
                const uint32_t c2 = *reinterpret_cast<const uint16_t*>(st + off + sp), d = *reinterpret_cast<const uint16_t*>(st + off + sp + 2);
                const int k00 = (32 - fx) * (32 - fy), k01 = fx * (32 - fy), k10 = (32 - fx) * fy, k11 = fx * fy;
                const int V0 = (int)(a & 0xFF) * k00 + (int)(b & 0xFF) * k01 + (int)(c2 & 0xFF) * k10 + (int)(d & 0xFF) * k11;
                const int V1 = (int)(a >> 8) * k00 + (int)(b >> 8) * k01 + (int)(c2 >> 8) * k10 + (int)(d >> 8) * k11;
                acc[2 * k] += fs_q(V0, fw); acc[2 * k + 1] += fs_q(V1, fw);
            }
        }
        __syncthreads();                               // everyone is done with this stage before the job after next overwrites it
        cur = nxt;
    }
    if (luma) {
        const int tx = tile % p.luma_tiles_x, ty = tile / p.luma_tiles_x;
        const int x = tx * 64 + lane_x * 4, y = ty * FS_ROWS + row;
        if (y >= p.H || x >= p.W) return;
        uint32_t o[4];
        #pragma unroll
        for (int k = 0; k < 4; k++) o[k] = min(fast_out(acc[k]), 255u);
        uint8_t* d = p.out + (size_t)y * p.out_pitch + x;
        if (x + 3 < p.W && ((p.out_pitch & 3) == 0)) *reinterpret_cast<uint32_t*>(d) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        else for (int k = 0; k < 4 && x + k < p.W; k++) d[k] = (uint8_t)o[k];
    } else {
        const int ct = tile - p.luma_tiles;
        const int tx = ct % p.chroma_tiles_x, ty = ct / p.chroma_tiles_x;
        const int cw = p.W / 2, ch = p.H / 2;
        const int x = tx * 32 + lane_x * 2, y = ty * FS_ROWS + row;
        if (y >= ch || x >= cw) return;
        // output channel 0 <- input channel 1, output channel 1 <- input channel 0 (mapper_fast.cpp:179-180)
        const uint32_t o[4] = { min(fast_out(acc[1]), 255u), min(fast_out(acc[0]), 255u), min(fast_out(acc[3]), 255u), min(fast_out(acc[2]), 255u) };
        uint8_t* d = p.out + (size_t)(p.H + y) * p.out_pitch + 2 * x;
        if (x + 1 < cw && ((p.out_pitch & 3) == 0)) *reinterpret_cast<uint32_t*>(d) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
        else { d[0] = (uint8_t)o[0]; d[1] = (uint8_t)o[1]; if (x + 1 < cw) { d[2] = (uint8_t)o[2]; d[3] = (uint8_t)o[3]; } }
    }
}

// cv::resize(src, Size(cols / 2, rows / 2)) INTER_LINEAR as the FastMapper constructor calls it (mapper_fast.cpp:57-58,70,98-99):
// exact 2x goes through the INTER_AREA fast path (imgwarp.cpp:3299-3303); f32: SSE body ((a + b) + (c + d)) * 0.25f for
// dx <= w - 4 (:2283-2318), scalar tail (((0 + a) + b) + c + d) * 0.25f (:2425-2437); u8: (a + b + c + d + 2) >> 2
Img<float> resize_half(const Img<float>& s)
{
    const int dw = s.w / 2, dh = s.h / 2;
    if (s.w != 2 * dw || s.h != 2 * dh) return resize_linear(s, dw, dh);
    Img<float> d(dw, dh);
    const int body = dw & ~3;
    for (int y = 0; y < dh; y++) {
        const float* r0 = s.row(2 * y), *r1 = s.row(2 * y + 1);
        float* o = d.row(y);
        for (int x = 0; x < body; x++) o[x] = ((r0[2 * x] + r0[2 * x + 1]) + (r1[2 * x] + r1[2 * x + 1])) * 0.25f;
        for (int x = body; x < dw; x++) { float sum = 0.f; sum += r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1]; o[x] = sum * 0.25f; }
    }
    return d;
}
Img<uint8_t> resize_half(const Img<uint8_t>& s)
{
    const int dw = s.w / 2, dh = s.h / 2;
    if (s.w != 2 * dw || s.h != 2 * dh) return resize_linear(s, dw, dh);
    Img<uint8_t> d(dw, dh);
    for (int y = 0; y < dh; y++) {
        const uint8_t* r0 = s.row(2 * y), *r1 = s.row(2 * y + 1);
        uint8_t* o = d.row(y);
        for (int x = 0; x < dw; x++) o[x] = (uint8_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
    }
    return d;
}

struct FastTable { Img<int32_t> sx, sy; Img<uint8_t> wgt; };      // 1/32-px fixed point (cv::convertMaps) and the u8 feather weight

inline short sat16(int v) { return (short)std::min(32767, std::max(-32768, v)); }

// entries + job lists of one plane pass.  tile_w: pixels per tile row (32 threads x px); sw, sh: source plane size in positions
void pack_plane(const std::vector<FastTable>& tb, int W, int H, int px, int sw_div, const std::vector<int>& in_w, const std::vector<int>& in_h,
                std::vector<uint32_t>& tile_job_start, std::vector<uint8_t>& job_cam, std::vector<uint32_t>& job_ofs, std::vector<uint4>& entries,
                int& tiles_x, int64_t& pairs)
{
    const int n = (int)tb.size(), tile_w = 32 * px;
    tiles_x = (W + tile_w - 1) / tile_w;
    const int tiles_y = (H + FT_ROWS - 1) / FT_ROWS;
    for (int ty = 0; ty < tiles_y; ty++)
        for (int tx = 0; tx < tiles_x; tx++) {
            for (int c = 0; c < n; c++) {
                bool any = false;
                for (int y = ty * FT_ROWS; y < std::min(H, (ty + 1) * FT_ROWS) && !any; y++) {
                    const uint8_t* w = tb[c].wgt.row(y);
                    for (int x = tx * tile_w; x < std::min(W, (tx + 1) * tile_w); x++) if (w[x]) { any = true; break; }
                }
                if (!any) continue;
                job_cam.push_back((uint8_t)c);
                job_ofs.push_back((uint32_t)entries.size());
                const size_t base = entries.size();
                entries.resize(base + (size_t)FT_THREADS * px / 2, make_uint4(0, 0, 0, 0));
                uint2* e = reinterpret_cast<uint2*>(entries.data() + base);
                const int sw = in_w[c] / sw_div, sh = in_h[c] / sw_div;
                for (int r = 0; r < FT_ROWS; r++) {
                    const int y = ty * FT_ROWS + r;
                    if (y >= H) break;
                    for (int l = 0; l < 32; l++)
                        for (int k = 0; k < px; k++) {
                            const int x = tx * tile_w + l * px + k;
                            if (x >= W) continue;
                            const uint32_t wv = tb[c].wgt.row(y)[x];
                            if (!wv) continue;
                            const int ix = tb[c].sx.row(y)[x], iy = tb[c].sy.row(y)[x];
                            const int ax = sat16(ix >> 5), ay = sat16(iy >> 5);
                            const bool inside = ax >= 0 && ay >= 0 && ax + 1 < sw && ay + 1 < sh;
                            const bool any_tap = ax + 1 >= 0 && ay + 1 >= 0 && ax < sw && ay < sh;
                            if (!any_tap) continue;                                   // all four taps read 0: contributes nothing
                            pairs++;
                            e[(size_t)(r * 32 + l) * px + k] = make_uint2((uint32_t)(uint16_t)ax | ((uint32_t)(uint16_t)ay << 16),
                                                                          (uint32_t)(ix & 31) | ((uint32_t)(iy & 31) << 8) | (wv << 16) | ((inside ? F_INSIDE : 0u) << 24));
                        }
                }
            }
            tile_job_start.push_back((uint32_t)job_cam.size());
        }
}


// staged tables of one plane pass (k_fast_staged).  px: positions per thread (4 luma, 2 chroma); unit: bytes per source position
// (1 luma, 2 interleaved chroma).  false: some job's footprint does not fit the stage -> the direct kernel serves the mapper.
bool pack_plane_staged(const std::vector<FastTable>& tb, int W, int H, int px, int unit, int sdiv, const std::vector<int>& in_w, const std::vector<int>& in_h,
                       std::vector<uint32_t>& tile_job_start, std::vector<FsJob>& jobs, std::vector<uint32_t>& entries, int& tiles_x, int64_t& pairs, int& max_box)
{
    const int n = (int)tb.size(), tile_w = 16 * px;
    tiles_x = (W + tile_w - 1) / tile_w;
    const int tiles_y = (H + FS_ROWS - 1) / FS_ROWS;
    for (int ty = 0; ty < tiles_y; ty++)
        for (int tx = 0; tx < tiles_x; tx++) {
            for (int c = 0; c < n; c++) {
                const int sw = in_w[c] / sdiv, sh = in_h[c] / sdiv;
                int xmin = INT32_MAX, xmax = INT32_MIN, ymin = INT32_MAX, ymax = INT32_MIN;
                auto tap = [&](int x, int y, int& ax, int& ay) {
                    if (!tb[c].wgt.row(y)[x]) return false;
                    ax = sat16(tb[c].sx.row(y)[x] >> 5); ay = sat16(tb[c].sy.row(y)[x] >> 5);
                    return ax + 1 >= 0 && ay + 1 >= 0 && ax < sw && ay < sh;              // at least one tap inside
                };
                for (int y = ty * FS_ROWS; y < std::min(H, (ty + 1) * FS_ROWS); y++)
                    for (int x = tx * tile_w; x < std::min(W, (tx + 1) * tile_w); x++) {
                        int ax, ay;
                        if (!tap(x, y, ax, ay)) continue;
                        xmin = std::min(xmin, ax); xmax = std::max(xmax, ax + 1); ymin = std::min(ymin, ay); ymax = std::max(ymax, ay + 1);
                    }
                if (xmin > xmax) continue;
                const int x0b = unit * xmin, x1b = unit * xmax + unit - 1;               // byte range of the taps in a source row
                const int bx0 = (int)std::floor(x0b / 16.0) * 16, bwc = std::max(2, (x1b - bx0) / 16 + 1), bh = ymax - ymin + 1;      // >= 2: rcp below must fit 32 bits
                if ((int64_t)bwc * bh * 16 > FS_STAGE || bx0 < -32768 || ymin < -32768 || ymin > 32767 || bwc > 65535 || bh > 65535) return false;
                max_box = std::max(max_box, bwc * bh * 16);
                FsJob job;
                job.bx0 = (short)bx0; job.by0 = (short)ymin; job.bwc = (unsigned short)bwc; job.bh = (unsigned short)bh;
                job.rcp = (uint32_t)(0x100000000ull / (uint64_t)bwc) + 1u; job.cam = (uint32_t)c;
                const size_t base = entries.size();
                entries.resize(base + (size_t)FT_THREADS * px, 0u);
                for (int r = 0; r < FS_ROWS; r++) {
                    const int y = ty * FS_ROWS + r;
                    if (y >= H) break;
                    for (int l = 0; l < 16; l++)
                        for (int k = 0; k < px; k++) {
                            const int x = tx * tile_w + l * px + k;
                            int ax, ay;
                            if (x >= W || !tap(x, y, ax, ay)) continue;
                            pairs++;
                            const uint32_t off = (uint32_t)((ay - ymin) * bwc * 16 + (unit * ax - bx0));
                            entries[base + (size_t)(r * 16 + l) * px + k] = off | ((uint32_t)(tb[c].sx.row(y)[x] & 31) << 14) | ((uint32_t)(tb[c].sy.row(y)[x] & 31) << 19) |
                                                                            ((uint32_t)tb[c].wgt.row(y)[x] << 24);
                        }
                }
                jobs.push_back(job);
            }
            tile_job_start.push_back((uint32_t)jobs.size());
        }
    return true;
}
}  // namespace
}  // namespace ob

using namespace ob;

struct octvr_fast {
    int device = 0, n = 0, W = 0, H = 0;
    std::vector<int> in_w, in_h;
    // the reference constructor's tables (host copies, kept for octvr_fast_debug_table)
    std::vector<FastTable> full, half;
    uint32_t* d_tile_job_start = nullptr; uint8_t* d_job_cam = nullptr; uint32_t* d_job_ofs = nullptr; uint4* d_entries = nullptr;
    int luma_tiles_x = 0, luma_tiles = 0, chroma_tiles_x = 0, tiles = 0;
    int64_t pairs_luma = 0, pairs_chroma = 0, table_bytes = 0;
    // staged layout (k_fast_staged, the default); the direct tables above are built only when it does not apply
    bool staged = false;
    uint32_t* d_s_tile_job_start = nullptr; FsJob* d_s_jobs = nullptr; uint32_t* d_s_entries = nullptr;
    int s_luma_tiles_x = 0, s_luma_tiles = 0, s_chroma_tiles_x = 0, s_tiles = 0, s_stage_bytes = 0;
    int s_occ = 8;                      // resident CTAs / SM asked of the staged kernel (measured on B200: 6 / 7 / 8 -> 70.3 / 64.3 / 64.2 us)
    uint32_t s_luma_jobs = 0;
    void build_direct();
    ~octvr_fast() { cudaFree(d_tile_job_start); cudaFree(d_job_cam); cudaFree(d_job_ofs); cudaFree(d_entries); cudaFree(d_s_tile_job_start); cudaFree(d_s_jobs); cudaFree(d_s_entries); }
};


void octvr_fast::build_direct()
{
    if (d_entries) return;
    std::vector<uint32_t> tjs{ 0 }, job_ofs;
    std::vector<uint8_t> job_cam;
    std::vector<uint4> entries;
    int64_t pl = 0, pc = 0;
    pack_plane(full, W, H, 4, 1, in_w, in_h, tjs, job_cam, job_ofs, entries, luma_tiles_x, pl);
    luma_tiles = (int)tjs.size() - 1;
    pack_plane(half, W / 2, H / 2, 2, 2, in_w, in_h, tjs, job_cam, job_ofs, entries, chroma_tiles_x, pc);
    tiles = (int)tjs.size() - 1;
    if (job_cam.empty()) fail(OCTVR_ERR_INVALID, "FastMapper: no camera contributes to the output");
    OB_CUDA(cudaSetDevice(device));
    auto up = [](auto*& d, const auto& v) {
        OB_CUDA(cudaMalloc(&d, v.size() * sizeof(v[0])));
        OB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
    };
    up(d_tile_job_start, tjs); up(d_job_cam, job_cam); up(d_job_ofs, job_ofs); up(d_entries, entries);
    if (!staged) { pairs_luma = pl; pairs_chroma = pc; table_bytes = (int64_t)(entries.size() * 16 + job_ofs.size() * 4 + job_cam.size() + tjs.size() * 4); }
}

extern "C" {

octvr_status octvr_fast_create(const octvr_template* t, const int* in_sizes_wh, int n_inputs, int device, octvr_fast** out)
{
    return guard([&] {
        OB_CHECK(t && in_sizes_wh && out, "null argument");
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) fail(OCTVR_ERR_CUDA, "no CUDA device (octvr_b200 has no CPU path)");
        OB_CHECK(device >= 0 && device < count, "bad device");
        const int n = (int)t->inputs.size();
        OB_CHECK(n_inputs == n && n >= 1 && n <= MAX_CAMS, "input count");
        if (!t->overlays.empty()) fail(OCTVR_ERR_INVALID, "FastMapper: overlay inputs are not supported (mapper_fast.cpp:31)");
        const int W = t->out_w, H = t->out_h;
        OB_CHECK(W > 0 && H > 1 && (W % 2) == 0, "FastMapper needs an even output width");
        std::unique_ptr<octvr_fast> f(new octvr_fast);
        f->device = device; f->n = n; f->W = W; f->H = H;
        for (int i = 0; i < n; i++) {
            const TInput& in = t->inputs[i];
            // "does not support ROI yet" (mapper_fast.cpp:50-51) + remap_weighted's dst.size() == map.size() (imgwarp.cpp:4645)
            if (in.roi.x != 0 || in.roi.y != 0 || in.roi.w != W || in.roi.h != H)
                fail(OCTVR_ERR_UNSUPPORTED, "FastMapper: inputs must cover the whole output frame (build the template with use_roi = 0)");
            const int iw = in_sizes_wh[2 * i], ih = in_sizes_wh[2 * i + 1];
            OB_CHECK(iw >= 2 && ih >= 2 && (iw % 2) == 0, "FastMapper needs NV12 inputs of even width");
            f->in_w.push_back(iw); f->in_h.push_back(ih);
        }
        // ---- the constructor's tables (mapper_fast.cpp:37-101)
        f->full.resize(n); f->half.resize(n);
        std::vector<Img<float>> wf(n);
        Img<float> total(W, H, 1e-5f);
        for (int i = 0; i < n; i++) {
            const TInput& in = t->inputs[i];
            quantise_map(in.map1, in.map2, f->in_w[i], f->in_h[i], f->full[i].sx, f->full[i].sy);
            Img<float> h1 = resize_half(in.map1), h2 = resize_half(in.map2);
            quantise_map(h1, h2, f->in_w[i] / 2, f->in_h[i] / 2, f->half[i].sx, f->half[i].sy);
            wf[i] = chamfer_l2(in.mask);
            for (size_t k = 0; k < wf[i].d.size(); k++) {
                const float v = wf[i].d[k] - 5.f;
                wf[i].d[k] = v > 0.f ? v : 0.f;
                total.d[k] = wf[i].d[k] + total.d[k];
            }
        }
        for (int i = 0; i < n; i++) {
            f->full[i].wgt = Img<uint8_t>(W, H);
            for (size_t k = 0; k < wf[i].d.size(); k++) {
                const float q = wf[i].d[k] / total.d[k];
                const long r = lrintf(q * 255.f + 0.f);
                f->full[i].wgt.d[k] = (uint8_t)std::min(255l, std::max(0l, r));
            }
            f->half[i].wgt = resize_half(f->full[i].wgt);
        }
        // ---- device tables: staged layout when every footprint fits the stage (OCTVR_FAST=direct forces the direct-gather tables)
        OB_CUDA(cudaSetDevice(device));
        auto up = [](auto*& d, const auto& v) {
            OB_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(v[0])));
            if (!v.empty()) OB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(v[0]), cudaMemcpyHostToDevice));
        };
        const char* mode = getenv("OCTVR_FAST");
        if (!(mode && std::string(mode) == "direct")) {
            std::vector<uint32_t> tjs{ 0 }, entries;
            std::vector<FsJob> jobs;
            int64_t pl = 0, pc = 0;
            int max_box = 0;
            bool ok = pack_plane_staged(f->full, W, H, 4, 1, 1, f->in_w, f->in_h, tjs, jobs, entries, f->s_luma_tiles_x, pl, max_box);
            f->s_luma_tiles = (int)tjs.size() - 1; f->s_luma_jobs = (uint32_t)jobs.size();
            ok = ok && pack_plane_staged(f->half, W / 2, H / 2, 2, 2, 2, f->in_w, f->in_h, tjs, jobs, entries, f->s_chroma_tiles_x, pc, max_box);
            f->s_tiles = (int)tjs.size() - 1;
            if (ok && !jobs.empty()) {
                up(f->d_s_tile_job_start, tjs); up(f->d_s_jobs, jobs); up(f->d_s_entries, entries);
                f->staged = true; f->pairs_luma = pl; f->pairs_chroma = pc;
                f->s_stage_bytes = (max_box + 1023) / 1024 * 1024;
                if (const char* e = getenv("OCTVR_FAST_OCC")) f->s_occ = atoi(e);
                f->table_bytes = (int64_t)(entries.size() * 4 + jobs.size() * sizeof(FsJob) + tjs.size() * 4);
            }
        }
        if (!f->staged) f->build_direct();
        *out = f.release();
    });
}

octvr_status octvr_fast_stitch_nv12(octvr_fast* f, const uint8_t* const* d_inputs, const size_t* pitches, int n_inputs,
                                    uint8_t* d_output, size_t out_pitch, void* stream)
{
    return guard([&] {
        OB_CHECK(f && d_inputs && pitches && d_output, "null argument");
        OB_CHECK(n_inputs == f->n, "input count");                                   // mapper_fast.cpp:156-160
        OB_CHECK(out_pitch >= (size_t)f->W, "output pitch");
        bool aligned = true;                              // the staged kernel copies 16-byte chunks
        for (int i = 0; i < f->n; i++) {
            OB_CHECK(d_inputs[i] && pitches[i] >= (size_t)f->in_w[i] && (pitches[i] % 2) == 0 && ((uintptr_t)d_inputs[i] % 2) == 0, "input frame");
            aligned = aligned && (pitches[i] % 16) == 0 && ((uintptr_t)d_inputs[i] % 16) == 0;
        }
        OB_CUDA(cudaSetDevice(f->device));
        if (f->staged && aligned) {
            FsParams p;
            memset(&p, 0, sizeof(p));
            for (int i = 0; i < f->n; i++)
                p.cam[i] = FastCam{ d_inputs[i], d_inputs[i] + (size_t)f->in_h[i] * pitches[i], (int)pitches[i], (int)pitches[i], f->in_w[i], f->in_h[i] };
            p.tile_job_start = f->d_s_tile_job_start; p.jobs = f->d_s_jobs; p.entries = f->d_s_entries;
            p.out = d_output; p.out_pitch = (int)out_pitch; p.W = f->W; p.H = f->H;
            p.luma_tiles_x = f->s_luma_tiles_x; p.luma_tiles = f->s_luma_tiles; p.chroma_tiles_x = f->s_chroma_tiles_x; p.luma_jobs = f->s_luma_jobs;
            p.stage_bytes = f->s_stage_bytes;
            const size_t smem = 2 * (size_t)f->s_stage_bytes;
            if (f->s_occ >= 8) k_fast_staged<8><<<f->s_tiles, FT_THREADS, smem, (cudaStream_t)stream>>>(p);
            else if (f->s_occ == 7) k_fast_staged<7><<<f->s_tiles, FT_THREADS, smem, (cudaStream_t)stream>>>(p);
            else k_fast_staged<6><<<f->s_tiles, FT_THREADS, smem, (cudaStream_t)stream>>>(p);
            OB_CUDA(cudaGetLastError());
            return;
        }
        f->build_direct();                                // frames that are not 16-byte aligned: direct-gather tables, built on first use
        FastParams p;
        memset(&p, 0, sizeof(p));
        for (int i = 0; i < f->n; i++)
            p.cam[i] = FastCam{ d_inputs[i], d_inputs[i] + (size_t)f->in_h[i] * pitches[i], (int)pitches[i], (int)pitches[i], f->in_w[i], f->in_h[i] };
        p.tile_job_start = f->d_tile_job_start; p.job_cam = f->d_job_cam; p.entries = f->d_entries; p.job_entry_ofs = f->d_job_ofs;
        p.out = d_output; p.out_pitch = (int)out_pitch; p.W = f->W; p.H = f->H;
        p.luma_tiles_x = f->luma_tiles_x; p.luma_tiles = f->luma_tiles; p.chroma_tiles_x = f->chroma_tiles_x;
        k_fast_nv12<<<f->tiles, FT_THREADS, 0, (cudaStream_t)stream>>>(p);
        OB_CUDA(cudaGetLastError());
    });
}

octvr_status octvr_fast_info(const octvr_fast* f, int* out_w, int* out_h, long long* pairs_luma, long long* pairs_chroma, long long* table_bytes)
{
    return guard([&] {
        OB_CHECK(f, "null argument");
        if (out_w) *out_w = f->W;
        if (out_h) *out_h = f->H;
        if (pairs_luma) *pairs_luma = f->pairs_luma;
        if (pairs_chroma) *pairs_chroma = f->pairs_chroma;
        if (table_bytes) *table_bytes = f->table_bytes;
    });
}

octvr_status octvr_fast_debug_table(const octvr_fast* f, int cam, int which, void* h_out)
{
    return guard([&] {
        OB_CHECK(f && h_out && cam >= 0 && cam < f->n && which >= 0 && which < 6, "bad argument");
        const FastTable& tb = (which % 2) ? f->half[cam] : f->full[cam];
        const size_t px = tb.wgt.d.size();
        if (which / 2 == 0) {            // map1: CV_16SC2
            short* o = (short*)h_out;
            for (size_t k = 0; k < px; k++) { o[2 * k] = sat16(tb.sx.d[k] >> 5); o[2 * k + 1] = sat16(tb.sy.d[k] >> 5); }
        } else if (which / 2 == 1) {     // map2: CV_16UC1
            uint16_t* o = (uint16_t*)h_out;
            for (size_t k = 0; k < px; k++) o[k] = (uint16_t)((tb.sy.d[k] & 31) * 32 + (tb.sx.d[k] & 31));
        } else memcpy(h_out, tb.wgt.d.data(), px);
    });
}

void octvr_fast_destroy(octvr_fast* f) { delete f; }

}  // extern "C"
