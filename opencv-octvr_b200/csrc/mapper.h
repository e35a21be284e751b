// csrc/mapper.h -- state behind octvr_mapper (vr::Mapper, modules/octvr/src/mapper.hpp:29-95).
#pragma once
#include "common.h"
#include <functional>
#include "kernels.cuh"
#include "template.h"
#include "post.h"

namespace ob { struct Multiband; }

struct octvr_mapper {
    int device = 0;
    int n = 0, out_w = 0, out_h = 0, blend = 0;
    bool gain = false;
    std::vector<int> in_w, in_h;
    float inv_n = 1.f;
    // per-camera RGBX planes (written by K_convert every frame) and optional full-size vignette maps
    std::vector<uint32_t*> d_rgbx;
    std::vector<float*> d_vig;
    // tile-compacted tables (feather / no-blend)
    int tiles_x = 0, tiles_y = 0;
    int band_y0 = 0, band_y1 = 0;       // output rows this mapper produces (multi-GPU row-band mode); default: all rows
    float texel_shift = 0.f;            // 0.5 with OCTVR_TEXEL_CENTER=1: sample at u * W - 0.5 (the reference's texture path) instead of u * W
    int band_x0 = 0, band_x1 = 0;       // output columns this mapper produces (column-band mode, multiband only); default: all columns
    std::vector<int> gain_col0, gain_col1; // per camera: the source columns the gain samples read
    std::vector<int> win_col0, win_w;   // input window (octvr_mapper_set_input_window): the frames passed to stitch hold source columns
                                        // [win_col0, win_col0 + win_w) only; default: the whole width
    std::vector<int> src_col0, src_col1; // per camera: the source columns [col0, col1) some table entry reads (only those are converted; a fisheye
                                        // circle in a 16:9 frame leaves the sides unused); default: all columns
    std::vector<int> src_row0, src_row1; // per camera: the source rows [row0, row1) some table entry of this mapper reads (a row-band mapper
                                        // converts only those to RGBX); default: all rows
    size_t njobs = 0;
    uint32_t* d_tile_job_start = nullptr;
    uint8_t* d_job_cam = nullptr;
    uint2* d_coords = nullptr;
    float* d_weights = nullptr;
    // staged layout (K_blend_staged)
    bool staged = false;
    ob::JobMeta* d_jobs = nullptr;
    uint2* d_entries = nullptr;
    void* d_tmaps = nullptr;
    int n_tmaps = 0;
    // ring layout (K_blend_ring, the default feather / no-blend kernel): d_tmaps as above
    bool ring = false;
    int ring_ctas = 0;                  // resident CTAs per SM (sizes the persistent grid)
    int sm_count = 0;
    uint4* d_rjobs = nullptr;
    uint4* d_rentries = nullptr;
    unsigned int* d_ring_counter = nullptr;
    unsigned long long* d_dbg_ring = nullptr;
    // fused layout (K_stitch_fused)
    bool fused = false;
    int fused_grid = 0;
    ob::FJob* d_fjobs = nullptr;
    ob::FBin* d_fbins = nullptr;
    uint16_t* d_fitems = nullptr;
    // gain compensation
    ob::GainParams gp;
    uint8_t* d_smask = nullptr; uint2* d_gcoord = nullptr; double* d_partial = nullptr;
    uint4* d_gsamples = nullptr; int2* d_gchunks = nullptr; unsigned long long* d_gtotals = nullptr;
    unsigned int* d_ticket = nullptr; double* d_gains = nullptr; float* d_gain_f32 = nullptr;
    unsigned long long* d_dbg = nullptr;
    int* d_gain_flag = nullptr; uint8_t* d_gain_lut = nullptr; double* h_gains = nullptr;
    // optional RGB result (Mapper::result)
    bool keep_rgb = false;
    bool rgb_this_frame = false;        // keep_rgb, or the after-blend stages below need the RGB888 result of this stitch
    uint8_t* d_rgb = nullptr;
    // after-blend stages (post.cu; mapper.cpp:279-312): overlay inputs, scale_output, preview
    int n_ov = 0;                       // overlay inputs follow the n blended inputs in every per-input array
    std::vector<uint2*> d_ov_coords;
    std::vector<ob::Rect> ov_roi;
    int scaled_w = 0, scaled_h = 0;     // size of the 4:2:0 output (== out_w x out_h unless scale_output differs)
    uint8_t* d_rgb_scaled = nullptr;
    ob::ResizePlan* scale_plan = nullptr;
    ob::ResizePlan* preview_plan = nullptr;
    // multiband state (blend > 0)
    ob::Multiband* mb = nullptr;
    // bookkeeping
    int64_t pairs = 0, roi_area = 0, table_bytes = 0;
    bool profiling = false;
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    cudaStream_t last_stream = nullptr;
    bool last_stream_valid = false;
    ~octvr_mapper();
};

namespace ob {
// mapper.cpp: CUtensorMap (128 bytes, written to out128) over an RGBX plane for box_w x box_h boxes, zero fill outside
void encode_rgbx_tensor_map(void* out128, const uint32_t* plane, int plane_w, int plane_h, int box_w, int box_h);
// pack.cu: the default feather / no-blend tables (K_blend_ring layout) packed by CUDA kernels; false = not applicable
// (then build_mapper's host packer runs).  m.d_rgbx must be allocated.
bool pack_ring_gpu(octvr_mapper& m, const octvr_template& t, int blend);
// multiband.cu
// host_xy(sx, sy) fills the host copies of the fixed-point coordinates; called only when the tables cannot be packed on the device
Multiband* multiband_create(octvr_mapper& m, const octvr_template& t,
                            const std::function<void(std::vector<Img<int32_t>>&, std::vector<Img<int32_t>>&)>& host_xy);
void multiband_stitch(octvr_mapper& m, const octvr_frame* out, cudaStream_t s);
int multiband_launches(const octvr_mapper& m);
void multiband_destroy(Multiband* mb);
}  // namespace ob

