/*
 * octvr_b200.h -- C ABI of the B200-native octvr stitch path (liboctvr_b200.so).
 *
 * Drop-in boundary for the per-frame stitching hot path of blahgeek/OpenCV-octVR's
 * modules/octvr.  Plain pointers and sizes only; no OpenCV, no torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the reference tree).
 * All functions return OCTVR_OK (0) or a negative octvr_status; octvr_last_error() gives the
 * message of the calling thread's last failure.  The C++ shim include/octvr.hpp turns these
 * into the exception types the reference throws.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device (sm_100a) and
 * fails with OCTVR_ERR_CUDA otherwise.
 *
 * Conventions that differ between the reference's GPU and CPU paths follow the CPU path
 * (SURVEY.md section 8c): pixel index u*W (not the texture's u*W-0.5), cv::remap's 1/32-px
 * fixed-point bilinear, limited-range BT.601 integer colour, 16-bit CPU MultiBandBlender.
 */
#ifndef OCTVR_B200_H
#define OCTVR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int octvr_status;
enum {
    OCTVR_OK = 0,
    OCTVR_ERR_INVALID = -1,   /* bad argument / shape (reference: CV_Assert -> cv::Exception) */
    OCTVR_ERR_FORMAT = -2,    /* bad .dat magic / JSON (reference: throw std::string, template.cpp:30,33,53,262) */
    OCTVR_ERR_CUDA = -3,      /* CUDA runtime failure or no device (reference: assert(false) without HAVE_CUDA, mapper.cpp:188,321) */
    OCTVR_ERR_UNSUPPORTED = -4/* camera model used in an unsupported direction (reference: throw NotImplemented, camera.hpp:92-103) */
};
const char* octvr_last_error(void);
const char* octvr_version(void);

/* ------------------------------------------------------------------ templates
 * vr::MapperTemplate (include/octvr.hpp:48-91, src/template.cpp). */
typedef struct octvr_template octvr_template;

/* explicit MapperTemplate(std::ifstream&) -- "VRv11" loader, template.cpp:258-314 */
octvr_status octvr_template_load_dat(const void* bytes, size_t n, octvr_template** out);
octvr_status octvr_template_load_file(const char* path, octvr_template** out);
/* MapperTemplate::dump, template.cpp:206-256 (creates seam masks first if absent, like :207-208) */
octvr_status octvr_template_dump_file(octvr_template* t, const char* path);

/* MapperTemplate(to, to_opts, width, height) + add_input(...) per "inputs"/"overlays" entry
 * (+ create_masks() when with_seam_masks) from the JSON schema apps/octvr/dump.cpp:71-127 reads.
 * Map generation (template.cpp:46-153, camera.cpp:189-315, cameras/ *) runs as a CUDA kernel on
 * `device` in f64.  width or height <= 0 -> derived from the output model's aspect ratio. */
octvr_status octvr_template_build_json(const char* json, int width, int height, int use_roi,
                                       int with_seam_masks, int device, octvr_template** out);

/* The same, incrementally, as the reference class is used (octvr.hpp:72-79):
 *   MapperTemplate(const std::string& to, const rapidjson::Value& to_opts, int width, int height)   -> octvr_template_create
 *   void add_input(const std::string& from, const rapidjson::Value& from_opts, bool overlay, bool use_roi) -> octvr_template_add_input
 * followed by octvr_template_create_masks().  Options are passed as the text of the JSON object (NULL / "" = {});
 * width or height <= 0 is derived from the output model's aspect ratio (template.cpp:23-44). */
octvr_status octvr_template_create(const char* to_type, const char* to_opts_json, int width, int height, int device,
                                   octvr_template** out);
octvr_status octvr_template_add_input(octvr_template* t, const char* from_type, const char* from_opts_json,
                                      int overlay, int use_roi);

/* Assemble a template from caller-made tables (what a reference-side shim holding a
 * vr::MapperTemplate passes in): per input i roi = rois_xywh[4i..], map1/map2 (f32, roi_h x roi_w,
 * normalised [0,1), -1 = no source), mask (u8), optional seam mask (u8, may be NULL array or NULL
 * entries), optional vignette (f32 vig_h x vig_w, NULL = none).  Arrays are copied. */
octvr_status octvr_template_from_arrays(int out_w, int out_h, int n_inputs, const int* rois_xywh,
                                        const float* const* map1, const float* const* map2,
                                        const uint8_t* const* mask, const uint8_t* const* seam_mask,
                                        const float* const* vignette, int vig_w, int vig_h,
                                        octvr_template** out);
/* Append one entry of MapperTemplate::overlay_inputs (octvr.hpp:63, filled by add_input(..., overlay = true),
 * template.cpp:46-153) to a template made by octvr_template_from_arrays.  Same array conventions; copied. */
octvr_status octvr_template_add_overlay(octvr_template* t, const int* roi_xywh, const float* map1, const float* map2,
                                        const uint8_t* mask, const float* vignette, int vig_w, int vig_h);
/* MapperTemplate::create_masks() with no images (DistanceSeamFinder), template.cpp:155-204 */
octvr_status octvr_template_create_masks(octvr_template* t);
/* Diagnostics: which code made the most recent set of seam masks -- 1 = the CUDA kernels of csrc/seam.cu (always, when a
 * device is present), 0 = the host routine (processes without a device), -1 = none made yet. */
int          octvr_debug_seam_backend(void);

octvr_status octvr_template_out_size(const octvr_template* t, int* w, int* h);
int          octvr_template_num_inputs(const octvr_template* t);
int          octvr_template_num_overlays(const octvr_template* t);
/* index < num_inputs: inputs[index]; else overlay_inputs[index - num_inputs]. Host pointers, owned by t.
 * seam_mask / vignette may come back NULL. */
octvr_status octvr_template_input(const octvr_template* t, int index, int roi_xywh[4],
                                  const float** map1, const float** map2, const uint8_t** mask,
                                  const uint8_t** seam_mask, const float** vignette, int vig_wh[2]);
void         octvr_template_destroy(octvr_template* t);

/* --------------------------------------------------------------------- frames */
/* 8-bit 4:2:0 frame as three plane pointers.  uv_pixel_stride 1 = planar (I420, and octvr's
 * packed "U|V side by side under Y" layout of mapper.hpp:75-83: u = base + H*pitch, v = u + W/2,
 * all pitches = W); 2 = semi-planar NV12 (v = u + 1).  Width and height must be even
 * (async.cpp:44-46). */
#define OCTVR_FMT_RGB24 3   /* uv_pixel_stride value: y = packed 8UC3 R,G,B rows (y_pitch >= 3 W), u / v ignored */
#define OCTVR_FMT_BGR24 4   /* same with B,G,R byte order (cv::Mat CV_8UC3 as cv::imread / cv::VideoCapture deliver it) */
typedef struct octvr_frame {
    uint8_t* y; uint8_t* u; uint8_t* v;
    size_t y_pitch, u_pitch, v_pitch;
    int uv_pixel_stride;
} octvr_frame;

/* --------------------------------------------------------------------- mapper
 * vr::Mapper (src/mapper.hpp:29-95, src/mapper.cpp:47-323). */
typedef struct octvr_mapper octvr_mapper;

/* Mapper(const MapperTemplate&, std::vector<cv::Size> in_sizes, int blend, bool enable_gain_compensator,
 *        cv::Size scale_output):  blend > 0 multiband width, < 0 feather border, 0 none (mapper.hpp:69-71);
 * in_sizes_wh = {w0,h0,w1,h1,...} for inputs then overlays; scale_w/h = 0 keeps the template size. */
octvr_status octvr_mapper_create(const octvr_template* t, const int* in_sizes_wh, int n_in,
                                 int blend, int enable_gain, int scale_w, int scale_h,
                                 int device, octvr_mapper** out);
/* Row-band mapper for the multi-GPU partition of ONE large frame (no counterpart in the reference, which
 * runs a frame on one GPU): like octvr_mapper_create, but the mapper only holds the tables of, and only writes, output
 * rows [band_y0, band_y1) (multiples of 32, or the frame height).  The frames passed to stitch are still full size; every
 * rank reads all inputs (an NCCL broadcast in sharding.py) and computes the same gains.  Multiband mappers keep a halo of
 * 4 * 2^bands rows of every pyramid around the band, so the band's rows equal the full-frame result bit for bit.  Not
 * available together with overlays or scale_output. */
octvr_status octvr_mapper_create_band(const octvr_template* t, const int* in_sizes_wh, int n_in,
                                      int blend, int enable_gain, int band_y0, int band_y1,
                                      int device, octvr_mapper** out);
/* The same with a window of output columns [x0, x1) x rows [y0, y1) (each range a multiple of 32 or the frame edge; 0, 0 =
 * the whole axis).  Column windows are a multiband (blend > 0) partition: the halo of 4 * 2^bands columns either side is a
 * far smaller share of a 7680-wide eye than a row halo is of its 1920 rows (BASELINE config C4).  Feather / no-blend
 * mappers split by rows only (OCTVR_ERR_UNSUPPORTED). */
octvr_status octvr_mapper_create_window(const octvr_template* t, const int* in_sizes_wh, int n_in,
                                        int blend, int enable_gain, int x0, int x1, int y0, int y1,
                                        int device, octvr_mapper** out);
/* void Mapper::stitch(std::vector<GpuMat>& inputs, GpuMat& output, GpuMat& preview, std::vector<double> gains)
 * mapper.cpp:193-323.  DEVICE pointers; asynchronous on `stream` (a cudaStream_t, NULL = default).
 * gains = NULL computes gains from this frame (when enabled); otherwise n_gains predefined gains
 * (async.cpp:75-86).  preview_rgb may be NULL. */
octvr_status octvr_mapper_stitch(octvr_mapper* m, const octvr_frame* d_inputs, int n_inputs,
                                 const octvr_frame* d_output, uint8_t* d_preview_rgb, size_t preview_pitch,
                                 int preview_w, int preview_h, const double* gains, int n_gains, void* stream);
/* Same with Mapper::stitch's packed single-plane layout (W x 1.5H, mapper.hpp:75-83). */
octvr_status octvr_mapper_stitch_packed(octvr_mapper* m, const uint8_t* const* d_inputs, const size_t* in_pitch,
                                        int n_inputs, uint8_t* d_output, size_t out_pitch,
                                        const double* gains, int n_gains, void* stream);
/* Row-band mode without a collection step: the rank that collects the frame allocates it with octvr_shared_alloc and
 * hands the 64-byte handle to the other ranks of the node (any channel; the tests use torch.distributed), they map it with
 * octvr_shared_open and pass pointers into it as the output frame of their band mapper: the blend kernels' stores then go
 * straight to the collecting GPU over NVLink peer memory.  A frame is complete on the collecting rank once every rank's
 * stitch has finished (the caller orders that, e.g. with one small NCCL all-reduce on the stitch streams).
 * octvr_shared_close(ptr, opened): opened = 1 for a pointer from octvr_shared_open, 0 to free the owner's allocation. */
octvr_status octvr_shared_alloc(size_t bytes, int device, void** d_ptr, unsigned char handle64[64]);
octvr_status octvr_shared_open(const unsigned char handle64[64], int device, void** d_ptr);
octvr_status octvr_shared_close(void* d_ptr, int opened);
/* Keep the RGB888 result at template size (Mapper::result, mapper.hpp:66) in addition to / instead of the YUV output. */
octvr_status octvr_mapper_set_keep_rgb(octvr_mapper* m, int on);
/* RGB888 result of the last stitch, device->host copy, synchronous.  Needs octvr_mapper_set_keep_rgb(m, 1). */
octvr_status octvr_mapper_result_rgb(octvr_mapper* m, uint8_t* h_rgb, size_t pitch);
/* std::vector<double> Mapper::gains() const, mapper.hpp:85-87.  Synchronises the last stitch's stream. */
octvr_status octvr_mapper_gains(octvr_mapper* m, double* out, int n);
/* Bookkeeping for the roofline: P = covered (pixel,camera) pairs, sum of ROI areas, bytes of the
 * packed device tables one stitch reads, kernels launched per stitch. */
octvr_status octvr_mapper_stats(const octvr_mapper* m, int64_t* pairs, int64_t* roi_area,
                                int64_t* table_bytes, int* launches_per_stitch);
/* Per blended input the source rows [lo, hi) this mapper converts and reads (the whole image unless it is a row-band
 * mapper); rows_lo_hi = {lo0, hi0, lo1, hi1, ...}, n = number of blended inputs. */
octvr_status octvr_mapper_source_rows(const octvr_mapper* m, int* rows_lo_hi, int n);
/* The same for source columns: [lo, hi) of every blended input that some table entry reads (a fisheye circle in a 16:9
 * frame leaves the sides unused); only those columns are converted.  The whole width for the non-default blend layouts. */
octvr_status octvr_mapper_source_cols(const octvr_mapper* m, int* cols_lo_hi, int n);
/* Declares that the frames of blended input `cam` passed to stitch hold source columns [col0, col0 + width) only (multiples of
 * 8 that cover octvr_mapper_source_cols): plane pointers address column col0 of each row, pitches need only cover `width`.
 * For callers that move frames between GPUs (row-band mode) or over PCIe and want to move only what is read. */
octvr_status octvr_mapper_set_input_window(octvr_mapper* m, int cam, int col0, int width);
/* The ingest side of input windows: source columns [col0[i], col0[i] + width[i]) of n packed (1.5 h x w) DEVICE frames (Mapper's
 * W x 1.5H layout, mapper.hpp:75-83; wh = {w0, h0, w1, h1, ...}) copied into packed (1.5 h x width) frames, U and V halves
 * cropped alike -- one launch.  Windows, widths and frame widths multiples of 32, pitches and pointers of 16. */
octvr_status octvr_crop_packed_frames(int n, const uint8_t* const* d_src, const size_t* src_pitch, const int* wh, const int* col0,
                                      const int* width, uint8_t* const* d_dst, const size_t* dst_pitch, void* stream);
/* time (ms, CUDA events on the stitch stream) the named stage of the LAST stitch took; stage =
 * "convert" | "gain" | "blend" | "total".  Only valid when octvr_mapper_set_profiling(m, 1). */
octvr_status octvr_mapper_set_profiling(octvr_mapper* m, int on);
/* %globaltimer stamps (ns) of the gain kernel's last CTA: start, ticket, reduced, solved, done, and the start of the
 * first gain CTA (diagnostics; SIX values). */
octvr_status octvr_mapper_debug_gain_ns(octvr_mapper* m, unsigned long long* out6);
/* The whole stamp buffer (diagnostics): with OCTVR_GAIN_TRACE set [6], [7] = first start / last end of the conversion CTAs and
 * [8 + 6c ..] = phase stamps of gain CTA c < 1024.  n <= 8 + 2 * 4096 words. */
octvr_status octvr_mapper_debug_gain_trace(octvr_mapper* m, unsigned long long* out, int n);
/* Counters of the feather kernel's TMA ring (diagnostics; all zero unless the library is built with -DRING_DEBUG=1):
 * jobs, ns waited for data, ns from TMA issue to first use, jobs waited for, ns issue -> complete over those; EIGHT values. */
octvr_status octvr_mapper_debug_ring(octvr_mapper* m, unsigned long long* out8);
/* Diagnostics / tests (host only, no device): the library's restatement of cv::fillPoly(img, {pts}, val) -- one contour,
 * 8UC1, lineType 8 (imgproc/src/drawing.cpp:1195-1404) -- that draws the camera masks of a JSON config
 * (octvr/src/camera.cpp:96-167).  pts_xy = {x0, y0, x1, y1, ...}. */
octvr_status octvr_debug_fill_poly(uint8_t* h_img, int w, int h, const int* pts_xy, int npts, int val);
octvr_status octvr_mapper_stage_ms(octvr_mapper* m, const char* stage, float* ms);
void         octvr_mapper_destroy(octvr_mapper* m);

/* ----------------------------------------------------------------- FastMapper
 * vr::FastMapper (include/octvr.hpp:123-144, src/mapper_fast.cpp): the NV12 path that never leaves 4:2:0 -- u8 feather
 * weights (border 5, no gain), full-resolution tables for luma and half-resolution tables for chroma, 16-bit accumulation,
 * / 255 (cv::remap_weighted, imgproc/src/opencl/remap_weighted.cl:20-77).  One CUDA launch per frame (csrc/fast.cu). */
typedef struct octvr_fast octvr_fast;
/* FastMapper(mt, in_sizes), mapper_fast.cpp:27-109.  The template must have no overlay inputs (:31) and full-frame inputs
 * (:50-51: build it with use_roi = 0, `octvr_dump -n`), else OCTVR_ERR_INVALID / OCTVR_ERR_UNSUPPORTED. */
octvr_status octvr_fast_create(const octvr_template* t, const int* in_sizes_wh, int n_inputs, int device, octvr_fast** out);
/* stitch_nv12(inputs, output), mapper_fast.cpp:153-195.  d_inputs[i]: device pointer to an (in_h + in_h / 2) x pitch NV12
 * frame (luma rows, then interleaved chroma rows -- the cv::UMat layout the reference asserts at :156-160); d_output:
 * (H + H / 2) x out_pitch, same layout.  As in the reference, output chroma byte 0 comes from input chroma byte 1 and
 * vice versa (:179-180).  Asynchronous on `stream`. */
octvr_status octvr_fast_stitch_nv12(octvr_fast* f, const uint8_t* const* d_inputs, const size_t* pitches, int n_inputs,
                                    uint8_t* d_output, size_t out_pitch, void* stream);
/* output size, contributing (pixel, camera) pairs of the luma / chroma tables, bytes of device tables; any pointer may be NULL */
octvr_status octvr_fast_info(const octvr_fast* f, int* out_w, int* out_h, long long* pairs_luma, long long* pairs_chroma, long long* table_bytes);
/* Tests: host copy of one of the constructor's tables for camera `cam`.  which: 0 map1s (W x H x 2 int16), 1 half_map1s,
 * 2 map2s (W x H uint16), 3 half_map2s, 4 feather_masks (u8), 5 half_feather_masks (mapper_fast.cpp:41-101). */
octvr_status octvr_fast_debug_table(const octvr_fast* f, int cam, int which, void* h_out);
void         octvr_fast_destroy(octvr_fast* f);

/* ---------------------------------------------------------------------- async
 * vr::AsyncMultiMapper (include/octvr.hpp:103-121, src/async.cpp). Host planes in, host planes out. */
typedef struct octvr_async octvr_async;

/* AsyncMultiMapper::New(mts, in_sizes, out_size, blend_modes, gain_modes, output_regions, preview_size)
 * async.cpp:195-350.  regions_xywh are fractions of the output frame (async.cpp:181-185). */
octvr_status octvr_async_create(const octvr_template* const* tmpls, int n_out,
                                const int* in_sizes_wh, int n_in, int out_w, int out_h,
                                const int* blend_modes, const int* gain_modes, const double* regions_xywh,
                                int preview_w, int preview_h, int device, octvr_async** out);
/* push(inputs, output): HOST frames; caller keeps them alive and untouched until the matching pop
 * returns (async.cpp:174-189). */
octvr_status octvr_async_push(octvr_async* a, const octvr_frame* h_inputs, int n_inputs, const octvr_frame* h_output);
/* pop(): blocks until the oldest pushed frame is complete (async.cpp:191-193). */
octvr_status octvr_async_pop(octvr_async* a);
octvr_status octvr_async_fps(octvr_async* a, double* fps);
/* The preview frame (preview_size, RGB888; every output region resized into its share of it, async.cpp:76-84,
 * mapper.cpp:308-312) of the frame popped last.  The reference hands this buffer to its GUI through Qt shared memory
 * (async.cpp:119-137, out of scope); here the caller reads it.  Valid until the next-but-two pop(). */
octvr_status octvr_async_preview(octvr_async* a, const uint8_t** h_rgb, size_t* pitch, int* w, int* h);
void         octvr_async_destroy(octvr_async* a);

#ifdef __cplusplus
}
#endif
#endif
