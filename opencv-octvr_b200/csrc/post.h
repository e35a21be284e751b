// csrc/post.h -- stages of Mapper::stitch after the blender (mapper.cpp:279-312): overlays, scale_output, preview.
#pragma once
#include "common.h"
#include "../../include/octvr_b200.h"
#include <memory>

namespace ob {

// cv::resize(8UC3, INTER_LINEAR) from (sw, sh) to (dw, dh): coefficient tables on the device
struct ResizePlan {
    int sw = 0, sh = 0, dw = 0, dh = 0, mode = 0;
    int* d_xofs = nullptr; short2* d_xa = nullptr; int* d_yofs = nullptr; short2* d_yb = nullptr;
    ~ResizePlan();
};
ResizePlan* resize_plan_create(int sw, int sh, int dw, int dh);
void launch_resize_rgb(const ResizePlan& r, const uint8_t* src, size_t src_pitch, uint8_t* dst, size_t dst_pitch, cudaStream_t s);
// warped overlay copied into the RGB888 result through its mask (mapper.cpp:279-282; the evident intent -- the reference
// copies from a buffer it never fills)
void launch_overlay(const uint32_t* rgbx, int src_pitch, const uint2* coords, const Rect& roi, uint8_t* rgb, size_t rgb_pitch, cudaStream_t s);
// cvtRGB24toYUV420P of the (scaled) result (mapper.cpp:294-306) with the CPU cvtColor arithmetic
void launch_rgb_to_yuv420(const uint8_t* rgb, size_t rgb_pitch, int w, int h, const octvr_frame& out, cudaStream_t s);

}  // namespace ob
