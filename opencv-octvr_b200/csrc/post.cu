// csrc/post.cu -- the tail of Mapper::stitch after the blender (modules/octvr/src/mapper.cpp:279-312): overlay inputs
// copied over the blended result through their masks, the optional resize of the result to `scale_output`, the RGB ->
// YUV 4:2:0 conversion of the (scaled) result and the preview resize.  These stages only run for mappers that have
// overlays, a scale_output different from the template size, or a preview buffer; the plain path writes its 4:2:0
// output straight from the blend kernel and never comes here.
//
// Arithmetic contract = the reference's CPU primitives: cv::remap INTER_LINEAR for the overlay warp
// (same 1/32-px fixed point as the blend kernels), cv::resize INTER_LINEAR for 8UC3 (imgwarp.cpp:3224-3500 dispatch and
// coefficients, :1387-1419 horizontal, :1477-1500 vertical; exactly-2x reductions take the INTER_AREA fast path,
// imgwarp.cpp:3299-3303 + ResizeAreaFastVec :2349-2390; equal sizes are a copy, :3261-3265), cv::cvtColor RGB2YUV_I420.
#include "post.h"
#include "prep.h"
#include "device_common.cuh"
#include <algorithm>

namespace ob {

// ---- overlay: remap the overlay's RGBX plane through its table and copy into the result where its mask is set ----
struct OverlayParams {
    const uint32_t* rgbx; int src_pitch;
    const uint2* coords;              // per roi pixel: table entry, C_VALID <=> mask != 0
    int rx, ry, rw, rh;
    uint8_t* rgb; uint32_t rgb_pitch;
};

__global__ void __launch_bounds__(256) k_overlay(const __grid_constant__ OverlayParams p)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= p.rw || y >= p.rh) return;
    const uint2 c = __ldg(p.coords + (size_t)y * p.rw + x);
    if (!(c.y & C_VALID)) return;                                     // copyTo(result(roi), mask): untouched outside the mask
    uint32_t t00, t01, t10, t11;
    fetch_taps(p.rgbx, p.src_pitch, c, t00, t01, t10, t11);
    int r, g, b;
    bilerp_rgbx(t00, t01, t10, t11, c.y & 31u, (c.y >> 5) & 31u, r, g, b);
    uint8_t* o = p.rgb + (size_t)(p.ry + y) * p.rgb_pitch + 3 * (p.rx + x);
    o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b;
}

// ---- cv::resize(8UC3, INTER_LINEAR) ----
struct ResizeParams {
    const uint8_t* src; uint32_t src_pitch; int sw, sh;
    uint8_t* dst; uint32_t dst_pitch; int dw, dh;
    int mode;                         // 0 copy, 1 exact 2x reduction (area), 2 linear
    const int* xofs; const short2* xa; const int* yofs; const short2* yb;
};

__global__ void __launch_bounds__(256) k_resize_rgb(const __grid_constant__ ResizeParams p)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= p.dw || y >= p.dh) return;
    uint8_t* o = p.dst + (size_t)y * p.dst_pitch + 3 * x;
    if (p.mode == 0) {
        const uint8_t* s = p.src + (size_t)y * p.src_pitch + 3 * x;
        o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
    } else if (p.mode == 1) {                                         // (a + b + c + d + 2) >> 2
        const uint8_t* s0 = p.src + (size_t)(2 * y) * p.src_pitch + 6 * x;
        const uint8_t* s1 = s0 + p.src_pitch;
        #pragma unroll
        for (int c = 0; c < 3; c++) o[c] = (uint8_t)((s0[c] + s0[c + 3] + s1[c] + s1[c + 3] + 2) >> 2);
    } else {
        const int sx = __ldg(p.xofs + x), sx1 = min(sx + 1, p.sw - 1);      // the second coefficient is 0 on the last column
        const short2 a = __ldg(p.xa + x), b = __ldg(p.yb + y);
        const int sy = __ldg(p.yofs + y);
        const uint8_t* s0 = p.src + (size_t)min(max(sy, 0), p.sh - 1) * p.src_pitch;
        const uint8_t* s1 = p.src + (size_t)min(max(sy + 1, 0), p.sh - 1) * p.src_pitch;
        #pragma unroll
        for (int c = 0; c < 3; c++) {
            const int top = s0[3 * sx + c] * a.x + s0[3 * sx1 + c] * a.y;
            const int bot = s1[3 * sx + c] * a.x + s1[3 * sx1 + c] * a.y;
            o[c] = (uint8_t)((((b.x * (top >> 4)) >> 16) + ((b.y * (bot >> 4)) >> 16) + 2) >> 2);
        }
    }
}

// ---- RGB888 -> 4:2:0 planes (cvtColor RGB2YUV_I420: chroma from the top-left pixel of each 2 x 2 block) ----
struct YuvParams {
    const uint8_t* rgb; uint32_t rgb_pitch; int w, h;
    uint8_t* oy; uint8_t* ou; uint8_t* ov; uint32_t oy_pitch, ou_pitch, ov_pitch; int uv_step;
};

__global__ void __launch_bounds__(256) k_rgb_to_yuv420(const __grid_constant__ YuvParams p)
{
    const int bx = blockIdx.x * 32 + threadIdx.x, by = blockIdx.y * 8 + threadIdx.y;     // 2 x 2 block index
    if (2 * bx >= p.w || 2 * by >= p.h) return;
    #pragma unroll
    for (int dy = 0; dy < 2; dy++) {
        const uint8_t* s = p.rgb + (size_t)(2 * by + dy) * p.rgb_pitch + 6 * bx;
        const uint32_t y0 = rgb_luma(s[0], s[1], s[2]), y1 = rgb_luma(s[3], s[4], s[5]);
        uint8_t* o = p.oy + (size_t)(2 * by + dy) * p.oy_pitch + 2 * bx;
        o[0] = (uint8_t)y0; o[1] = (uint8_t)y1;
        if (dy == 0) {
            p.ou[(size_t)by * p.ou_pitch + (size_t)bx * p.uv_step] = (uint8_t)rgb_cb(s[0], s[1], s[2]);
            p.ov[(size_t)by * p.ov_pitch + (size_t)bx * p.uv_step] = (uint8_t)rgb_cr(s[0], s[1], s[2]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ host
ResizePlan::~ResizePlan() { cudaFree(d_xofs); cudaFree(d_xa); cudaFree(d_yofs); cudaFree(d_yb); }

template <class T> static T* upload(const std::vector<T>& v)
{
    T* d = nullptr;
    OB_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
    if (!v.empty()) OB_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

ResizePlan* resize_plan_create(int sw, int sh, int dw, int dh)
{
    OB_CHECK(sw > 0 && sh > 0 && dw > 0 && dh > 0, "resize: empty size");
    std::unique_ptr<ResizePlan> r(new ResizePlan);
    r->sw = sw; r->sh = sh; r->dw = dw; r->dh = dh;
    if (sw == dw && sh == dh) { r->mode = 0; return r.release(); }
    if (sw == 2 * dw && sh == 2 * dh) { r->mode = 1; return r.release(); }       // scale_x == scale_y == 2 exactly
    r->mode = 2;
    std::vector<int> xofs, yofs;
    std::vector<short> xa, yb;
    resize_linear_tables(sw, dw, true, xofs, xa);
    resize_linear_tables(sh, dh, false, yofs, yb);
    r->d_xofs = upload(xofs); r->d_yofs = upload(yofs);
    r->d_xa = reinterpret_cast<short2*>(upload(xa)); r->d_yb = reinterpret_cast<short2*>(upload(yb));
    return r.release();
}

void launch_resize_rgb(const ResizePlan& r, const uint8_t* src, size_t src_pitch, uint8_t* dst, size_t dst_pitch, cudaStream_t s)
{
    ResizeParams p;
    p.src = src; p.src_pitch = (uint32_t)src_pitch; p.sw = r.sw; p.sh = r.sh;
    p.dst = dst; p.dst_pitch = (uint32_t)dst_pitch; p.dw = r.dw; p.dh = r.dh;
    p.mode = r.mode; p.xofs = r.d_xofs; p.xa = r.d_xa; p.yofs = r.d_yofs; p.yb = r.d_yb;
    k_resize_rgb<<<dim3((r.dw + 31) / 32, (r.dh + 7) / 8), dim3(32, 8), 0, s>>>(p);
}

void launch_overlay(const uint32_t* rgbx, int src_pitch, const uint2* coords, const Rect& roi, uint8_t* rgb, size_t rgb_pitch, cudaStream_t s)
{
    if (roi.w <= 0 || roi.h <= 0) return;
    OverlayParams p;
    p.rgbx = rgbx; p.src_pitch = src_pitch; p.coords = coords;
    p.rx = roi.x; p.ry = roi.y; p.rw = roi.w; p.rh = roi.h; p.rgb = rgb; p.rgb_pitch = (uint32_t)rgb_pitch;
    k_overlay<<<dim3((roi.w + 31) / 32, (roi.h + 7) / 8), dim3(32, 8), 0, s>>>(p);
}

void launch_rgb_to_yuv420(const uint8_t* rgb, size_t rgb_pitch, int w, int h, const octvr_frame& out, cudaStream_t s)
{
    YuvParams p;
    p.rgb = rgb; p.rgb_pitch = (uint32_t)rgb_pitch; p.w = w; p.h = h;
    p.oy = out.y; p.ou = out.u; p.ov = out.v;
    p.oy_pitch = (uint32_t)out.y_pitch; p.ou_pitch = (uint32_t)out.u_pitch; p.ov_pitch = (uint32_t)out.v_pitch; p.uv_step = out.uv_pixel_stride;
    k_rgb_to_yuv420<<<dim3((w / 2 + 31) / 32, (h / 2 + 7) / 8), dim3(32, 8), 0, s>>>(p);
}

// ---- k_crop_frames: source columns [col0, col0 + cw) of up to 16 packed (1.5 h x w) frames -> packed (1.5 h x cw) frames,
//      one launch (the ingest side of octvr_mapper_set_input_window: only the columns that are read travel between GPUs) ----
struct CropFrame { const uint8_t* src; uint8_t* dst; uint32_t src_pitch, dst_pitch; int w, h, col0, cw; };
struct CropParams { CropFrame f[MAX_CAMS]; };
__global__ void __launch_bounds__(256) k_crop_frames(const __grid_constant__ CropParams p)
{
    const CropFrame& f = p.f[blockIdx.z];
    const int chunks = f.cw / 16, rows = f.h + f.h / 2;
    const int c = blockIdx.x * 256 + threadIdx.x, r = blockIdx.y;
    if (c >= chunks || r >= rows) return;
    int sx = f.col0 + c * 16;                                       // luma row
    if (r >= f.h) {                                                 // U | V halves side by side
        const int half = f.cw / 2, x = c * 16;
        sx = x < half ? f.col0 / 2 + x : f.w / 2 + f.col0 / 2 + (x - half);
    }
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(f.src + (size_t)r * f.src_pitch + sx));
    *reinterpret_cast<uint4*>(f.dst + (size_t)r * f.dst_pitch + c * 16) = v;
}

}  // namespace ob

extern "C" octvr_status octvr_crop_packed_frames(int n, const uint8_t* const* d_src, const size_t* src_pitch, const int* wh, const int* col0,
                                                 const int* width, uint8_t* const* d_dst, const size_t* dst_pitch, void* stream)
{
    using namespace ob;
    return guard([&] {
        OB_CHECK(n >= 1 && n <= MAX_CAMS && d_src && src_pitch && wh && col0 && width && d_dst && dst_pitch, "bad argument");
        CropParams p;
        memset(&p, 0, sizeof(p));
        int max_chunks = 0, max_rows = 0;
        for (int i = 0; i < n; i++) {
            const int w = wh[2 * i], h = wh[2 * i + 1];
            OB_CHECK(d_src[i] && d_dst[i] && col0[i] >= 0 && width[i] > 0 && col0[i] + width[i] <= w, "window outside the frame");
            // 16-byte copies: the luma and both chroma segments must start on 16-byte boundaries
            OB_CHECK(col0[i] % 32 == 0 && width[i] % 32 == 0 && w % 32 == 0 && src_pitch[i] % 16 == 0 && dst_pitch[i] % 16 == 0 &&
                     (uintptr_t)d_src[i] % 16 == 0 && (uintptr_t)d_dst[i] % 16 == 0, "windows, widths and pitches must be multiples of 32 / 16 bytes");
            OB_CHECK(src_pitch[i] >= (size_t)w && dst_pitch[i] >= (size_t)width[i] && h % 2 == 0, "pitch");
            p.f[i] = CropFrame{ d_src[i], d_dst[i], (uint32_t)src_pitch[i], (uint32_t)dst_pitch[i], w, h, col0[i], width[i] };
            max_chunks = std::max(max_chunks, width[i] / 16); max_rows = std::max(max_rows, h + h / 2);
        }
        k_crop_frames<<<dim3((max_chunks + 255) / 256, max_rows, n), 256, 0, (cudaStream_t)stream>>>(p);
        OB_CUDA(cudaGetLastError());
    });
}
