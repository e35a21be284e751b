// csrc/kernels.cuh -- launch interfaces of the per-frame CUDA kernels (sm_100a).
#pragma once
#include "common.h"

namespace ob {

// ---- K_convert: planar/semi-planar 4:2:0 -> RGBX8888 (one u32 per source pixel), optional vignette ----
struct CamSrc {
    const uint8_t* y; const uint8_t* u; const uint8_t* v;
    uint32_t y_pitch, u_pitch, v_pitch;
    int uv_step;              // 1 planar, 2 NV12; packed RGB24 / BGR24 input: 3, with y / u / v = the R / G / B byte of pixel 0 and all
                              // three pitches = the row pitch (rgb = 1)
    int rgb;
    int row0, row1;           // rows [row0, row1) are converted (row0 even); the whole image unless the mapper is a row-band mapper
    int xlo, xhi;             // addressable columns [xlo, xhi] of the frame (the whole width unless the mapper has an input window)
    int col0, col1;           // columns [col0, col1) are converted (col0 a multiple of 8): the part of the frame some table entry reads
    int w, h;
    uint32_t* rgbx;           // w*h, pitch = w pixels
    const float* vignette;    // w*h f32 or null
    int aligned4;             // y rows 4-byte aligned and w % 4 == 0
};
struct ConvertParams {
    CamSrc cam[MAX_CAMS];
    int n;
    int grid_x, grid_y;              // ceil(max w / 256), ceil(max h / 16)
};


// ---- K_gain: working-scale statistics, least-squares solve, exact gain tables ----
struct GainCam {
    int sx, sy, sw, sh;        // working-scale ROI (mapper.cpp:95-99)
    uint32_t off;              // offset of this camera in smask / gcoord / sq
};
struct GainParams {
    GainCam cam[MAX_CAMS];
    CamSrc src[MAX_CAMS];      // this frame's input planes: the gain samples are converted on the fly, so the kernel
                               // does not depend on K_convert and runs concurrently with it on a second stream
    int n;
    uint32_t total;            // sum of sw*sh
    const uint8_t* smask;      // linearly resized masks (mapper.cpp:113-114)
    const uint2* gcoord;       // NEAREST-resized sample (mapper.cpp:235-237): x = (ix+1) | (iy+1)<<16 of the top-left tap, y = fx|fy<<5|flags
    const uint4* samples;      // the same entries regrouped per canvas chunk (512 slots each): {entry.x, entry.y, camera | local pixel << 8, 0}
    const int2* chunks;        // [grid] (unused)
    unsigned long long* totals;   // [n_pairs][5] exact integer sums (count, hi/lo of sum_i, hi/lo of sum_j); zeroed by the last CTA
    int cx0, cy0, cw, ch;      // working-scale canvas = union of the scaled ROIs
    int n_pairs, grid;         // pairs (i<=j); CTAs launched
    double* partial;           // [grid][n_pairs][3] : count, sum_i, sum_j
    unsigned int* ticket;
    double* gains;             // [n] f64 (Mapper::gains())
    float* gain_f32;           // [n] verified f32 multiplier
    int* gain_flag;            // [n] 1 -> use the LUT
    uint8_t* gain_lut;         // [n][256] exact sat_u8(rint(v*g)) in f64
    unsigned long long* dbg;   // optional: %globaltimer stamps of the last CTA (start, ticket, reduced, solved, done), [5] first gain CTA start,
    unsigned long long* dbg_trace;   // == dbg when OCTVR_GAIN_TRACE is set: the conversion CTAs stamp [6] first start, [7] last end, gain CTA c < 1024 stamps
                                     //   [8 + 6c ..]: start, samples loaded, norms done, pair sums done, fence done
};
// gp == nullptr: conversion only
void launch_convert_gain(const ConvertParams& cp, const GainParams* gp, cudaStream_t s);
void launch_gain_finalize(const GainParams& p, cudaStream_t s);   // gains[] already set (predefined gains)

// ---- K_blend: fused remap (1/32-px fixed-point bilinear) + gain + weighted accumulate + normalise
//      + RGB -> YUV 4:2:0 store ----
struct BlendParams {
    const uint32_t* rgbx[MAX_CAMS];
    int src_pitch[MAX_CAMS];
    const uint32_t* tile_job_start;   // [tiles+1]
    const uint8_t* job_cam;           // [jobs]
    const uint2* coords;              // [jobs*TILE_PX]
    const float* weights;             // [jobs*TILE_PX]
    int tiles_x, tiles_y, out_w, out_h;
    int tile_y0, tiles_y_run;                 // row band of this mapper (tile rows [tile_y0, tile_y0 + tiles_y_run)); default: all
    uint8_t* oy; uint8_t* ou; uint8_t* ov;
    uint32_t oy_pitch, ou_pitch, ov_pitch;
    int uv_step;
    uint8_t* rgb_out; uint32_t rgb_pitch;     // optional RGB888 result
    const float* gain_f32; const int* gain_flag; const uint8_t* gain_lut;
    int use_gain;
    float inv_n;
};
void launch_blend(const BlendParams& p, cudaStream_t s);

// ---- K_blend_staged: same arithmetic, but each job's SOURCE FOOTPRINT (bounding box of its bilinear taps in the
//      camera's RGBX plane) is brought into shared memory by TMA (cp.async.bulk.tensor.2d, one instruction per
//      camera, zero-filled outside the image = BORDER_CONSTANT, completion on an mbarrier) for ALL cameras of the
//      tile at once, and the taps are read with LDS.  The table shrinks to 8 B per pair:
//      {stage-relative tap offset | fx<<16 | fy<<21, f32 weight}. ----
constexpr int STAGE_CAP = 6144;        // shared-memory stage capacity in pixels (24 KB): all footprints of a job group
// One 32-byte record per (tile, slot); a tile owns MAX_CAMS consecutive slots, so the CTA finds everything it needs
// with ONE coalesced 512-byte read (no tile -> job-list indirection).  bw x bh = TMA box (px); soff = px offset in
// the stage; tmap = descriptor index; nj / j0 (jobs of the tile, first entry block) are replicated in every slot.
struct JobMeta { int cam; int bx0, by0; int bw, bh; int soff; int grp_nj; int tmap; int j0; int pad[7]; };
static_assert(sizeof(JobMeta) == 64, "JobMeta");
struct StagedParams {
    const uint32_t* rgbx[MAX_CAMS];
    int src_w[MAX_CAMS], src_h[MAX_CAMS];
    const JobMeta* jobs;              // [tiles][MAX_CAMS]
    const void* tmaps;                // CUtensorMap[...]: one per (camera, box size) in use, 128 B each
    const uint2* entries;             // [jobs*TILE_PX]
    int tiles_x, tiles_y, out_w, out_h;
    int tile_y0, tiles_y_run;         // row band (see BlendParams)
    uint8_t* oy; uint8_t* ou; uint8_t* ov;
    uint32_t oy_pitch, ou_pitch, ov_pitch;
    int uv_step;
    uint8_t* rgb_out; uint32_t rgb_pitch;
    const float* gain_f32; const int* gain_flag; const uint8_t* gain_lut;
    int use_gain;
    float inv_n;
};
void launch_blend_staged(const StagedParams& p, cudaStream_t s);

// ---- K_blend_ring (default feather / no-blend kernel; blend_ring.cu): persistent CTAs, a producer warp feeding a byte
//      ring of shared memory with [4 KB of table entries | source box] per job by TMA, four consumer warps with four
//      pixels per thread, packed-FP32 arithmetic, warp-private 4:2:0 epilogue with 128-bit stores. ----
#ifndef RING_KB
#define RING_KB 28
#endif
constexpr int RING_BYTES = RING_KB * 1024;          // shared-memory ring of a CTA (source boxes)
constexpr int RING_ENT_BYTES = TILE_PX * 8;    // table entries of one job
// job record (uint4), tile t owns records [t * nslot, (t + 1) * nslot) (nslot = cameras; unused slots are all zero):
//   x = bx0 (s16) | by0 (s16) << 16 (top-left of the TMA box in the camera's RGBX plane), y = tensor map index | camera << 16,
//   z = box bytes (bw * bh * 4, never 0 for a job), w = box pitch in px (bw, 12 bits) | job index << 12
// entries: per job [2][128] uint4 {ex_a, ex_b, weight_a, weight_b}: thread t of the tile's 128 owns column t & 31 and
//          rows 2w, 2w+1 (first uint4), 2w+8, 2w+9 (second) with w = t >> 5;
//          ex = byte offset of the top-left tap inside the box << 16 | fy << 8 | fx
struct RingParams {
    const uint4* jobs;                // [tiles_x * tiles_y][nslot]
    int nslot;
    const uint4* entries;
    const void* tmaps;                // CUtensorMap[...], 128 B each
    unsigned int* counter;            // tile ticket (zero between launches; re-armed by the kernel)
    unsigned long long* dbg;          // diagnostics (build with -DRING_DEBUG=1): [0] jobs, [1] sum ns waited for a job's data, [2] sum ns between TMA issue and first use,
                                      //   [3] jobs waited for > 200 ns, [4] sum ns issue -> data complete over the jobs waited for
    int tiles_x, out_w, out_h;
    int tile0, ntiles;                // row band: tiles [tile0, tile0 + ntiles)
    uint8_t* oy; uint8_t* ou; uint8_t* ov;
    uint32_t oy_pitch, ou_pitch, ov_pitch;
    int uv_step;
    int fast_store;                   // planar 4:2:0 output with 16-byte aligned rows and no RGB result: 128-bit stores
    uint8_t* rgb_out; uint32_t rgb_pitch;
    const float* gain_f32; const int* gain_flag; const uint8_t* gain_lut;
    int use_gain;
    float inv_n;
};
int ring_ctas_per_sm();
void launch_blend_ring(const RingParams& p, int grid, cudaStream_t s);

// ---- K_stitch_fused: the whole feather / no-blend frame in ONE kernel, no intermediate image in HBM ----
//      Persistent CTAs (4 per SM) walk a host-balanced list of 32x32 output tiles.  For every (tile, camera) job the
//      CTA (1) converts the job's source blocks of the 4:2:0 input planes (L2-resident: 37 MB for the 6 x 2.7K rig)
//      into an RGBX stage in shared memory (same integer BT.601 arithmetic as K_convert), (2) gathers the four
//      bilinear taps per output pixel from that stage with LDS, interpolates (IDP.2A), applies gain and weight and
//      accumulates in registers; a finished tile is normalised, converted to YUV 4:2:0 and stored with 128-bit stores.
//      Everything a job needs is in flight one or two jobs ahead: its 8 B/pair table entries by a TMA bulk copy
//      (cp.async.bulk + mbarrier transaction count), its input bytes by per-thread cp.async into thread-private
//      slots, its job record and item descriptors by plain loads into registers.
constexpr int FT_W = 32, FT_H = 32, FT_PX = FT_W * FT_H, FT_PPT = 4;
constexpr int FT_THREADS = FT_PX / FT_PPT;
constexpr int FUSED_CAP = 6144;          // RGBX stage capacity in pixels (24 KB); a job whose box is larger is split by rows
constexpr int FUSED_MAXITEMS = 2 * FT_THREADS;   // conversion items (4 px x 2 rows of the source box) per job
// conversion item descriptor (u16): row pair (7 bits) | 4-px group << 7 (7 bits) | class << 14
enum { FITEM_ZERO = 0, FITEM_FAST = 1, FITEM_SLOW = 2 };     // outside the source (-> 0), inside, straddles the border
constexpr int FJOB_LAST = (int)0x80000000u;                  // cam field: last job of its output tile
constexpr int FJOB_SYNC = 0x40000000;                        // cam field: this job's box overlaps the previous job's box in the stage:
                                                             //   a CTA barrier must separate the previous gather from this conversion
struct FJob {                            // 32 B, one per (tile, camera[, row range]), in the order the CTA consumes them
    int cam;                             // | FJOB_LAST
    int bx0, by0;                        // top-left of the source box (bx0 % 8 == 0, by0 % 2 == 0, may be negative)
    int bw;                              // box width in px (bw % 8 == 0); box area <= FUSED_CAP
    int nitems;                          // conversion items: only the 4x2 blocks some bilinear tap touches
    uint32_t items_off;                  // (unused) descriptors of job j are at FusedParams::items[j * FUSED_MAXITEMS ...]; item i -> thread i % FT_THREADS
    uint32_t tile_xy;                    // output tile: x | y << 16 (tile units)
    int stage_off;                       // where the box lives in the RGBX stage (px, multiple of 8); consecutive jobs alternate
};
static_assert(sizeof(FJob) == 32, "FJob is read as two uint4");
struct FBin { int start, end; };         // a CTA's jobs [start, end)
struct FusedParams {
    CamSrc cam[MAX_CAMS];
    int n;
    const FJob* jobs;                    // [jobs] in schedule order, grouped by CTA
    const FBin* bins;                    // [grid]
    const uint16_t* items;
    const uint2* entries;                // [jobs * FT_PX]: {byte offset in the stage | fy << 16 | fx << 24, f32 weight}
    int out_w, out_h;
    uint8_t* oy; uint8_t* ou; uint8_t* ov;
    uint32_t oy_pitch, ou_pitch, ov_pitch;
    int uv_step;
    uint8_t* rgb_out; uint32_t rgb_pitch;
    const float* gain_f32; const int* gain_flag; const uint8_t* gain_lut;
    int use_gain;
    float inv_n;
};
int fused_ctas_per_sm();                 // occupancy of k_stitch_fused (sizes the persistent grid)
void launch_stitch_fused(const FusedParams& p, int grid, cudaStream_t s);

}  // namespace ob
