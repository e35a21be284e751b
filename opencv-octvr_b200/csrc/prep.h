// csrc/prep.h -- init-time host preparation (runs once per Mapper / template, like the reference's
// own CPU init in blenders.cpp:531-572 and template.cpp:155-204).
#pragma once
#include "common.h"

namespace ob {

// cv::distanceTransform(mask, DIST_L2, 3): two-pass 3x3 chamfer in 16.16 fixed point
// (imgproc/src/distransform.cpp:69-139).
Img<float> chamfer_l2(const Img<uint8_t>& mask);
// cv::resize(..., INTER_LINEAR) for 8UC1 (imgwarp.cpp:3224-3500,1387-1500) and 32FC1.
Img<uint8_t> resize_linear(const Img<uint8_t>& src, int dw, int dh);
Img<float> resize_linear(const Img<float>& src, int dw, int dh);
// per destination index: source offset (x: clamped to [0, src-1]; y: unclamped, imgwarp.cpp:3404-3447) and the two
// 11-bit coefficients {round((1-f)*2048), round(f*2048)} of the 8-bit INTER_LINEAR path
void resize_linear_tables(int src, int dst, bool clamp_ofs, std::vector<int>& ofs, std::vector<short>& coef);
// cv::fillPoly(img, {pts}, val) for one contour on an 8UC1 image, lineType 8, shift 0 (imgproc/src/drawing.cpp:1195-1404:
// outline by 8-connected lines through cv::LineIterator + cv::clipLine :80-236, interior by the 16.16 fixed-point scan-line
// edge table).  Camera::Camera draws the `selection` rectangle and the polygonal exclude / include masks with it
// (octvr/src/camera.cpp:96-167).
void fill_poly_u8(uint8_t* img, int w, int h, const int* pts_xy, int npts, uint8_t val);
// The pixels cv::imdecode(bytes, IMREAD_COLOR) yields for a PNG file (camera.cpp:169-187 reads exclude / include masks that
// way): 8-bit B,G,R per pixel.  Grey is replicated, a palette expanded, alpha dropped, 16-bit samples reduced to their
// high byte, 1 / 2 / 4-bit grey scaled to 0..255 -- what libpng does for OpenCV.  Inflate comes from zlib; interlaced files
// are rejected (OCTVR_ERR_UNSUPPORTED).
void png_decode_bgr(const uint8_t* bytes, size_t n, int& w, int& h, std::vector<uint8_t>& bgr);
// cv::pyrDown for 32FC1 (pyramids.cpp:849-964 incl. the SSE association of :143-185)
Img<float> pyrdown_f32(const Img<float>& src);

// FeatherGPUBlender ctor recipe (stitching/src/blenders.cpp:531-572): W_i = N*max(DT_i-border,0)/(1e-5+sum)
std::vector<Img<float>> feather_weights(const std::vector<TInput>& in, int border);
// blend == 0: masked copies in camera order (mapper.cpp:269-275) == weight 1 for the LAST covering camera
std::vector<Img<float>> overwrite_weights(const std::vector<TInput>& in);
// MapperTemplate::create_masks() without images (template.cpp:155-204, seam_finders.cpp:86-133)
// distance_seam_masks: on the GPU (seam.cu) whenever a CUDA device is present; the host routine only serves processes without
// a device (loading / dumping .dat files needs none).  seam_backend() tells which one made the last set (1 = GPU, 0 = host).
std::vector<Img<uint8_t>> distance_seam_masks(const std::vector<TInput>& in, int out_w, int device = -1);
std::vector<Img<uint8_t>> distance_seam_masks_host(const std::vector<TInput>& in, int out_w);
bool distance_seam_masks_gpu(const std::vector<TInput>& in, int out_w, int device, std::vector<Img<uint8_t>>& out);
int seam_backend();
// seam.cu: d_dist[i] = cv::distanceTransform(d_mask[i], DIST_L2, 3) for n DEVICE images (one CTA each, widths <= 8192)
void chamfer_l2_gpu_batch(const uint8_t* const* d_mask, const int* w, const int* h, float* const* d_dist, int n);

// cv::remap coordinate quantisation for planar f32 maps (imgwarp.cpp:4383-4442) applied to
// fl32(map * size) (template.cpp:175-176).  Returns 1/32-px fixed point sx, sy.
// shift: subtracted from the pixel coordinate before it is quantised -- 0 (cv::remap's pixel-index convention, the default)
// or 0.5 (the texel-centre convention of the reference's CUDA path: tex2D at u * W samples index u * W - 0.5)
void quantise_map(const Img<float>& map1, const Img<float>& map2, int src_w, int src_h,
                  Img<int32_t>& sx, Img<int32_t>& sy, float shift = 0.f);

}  // namespace ob
