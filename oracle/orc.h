/*
 * oracle/orc.h -- CPU ORACLE for the octvr stitch hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * CPU algorithms (blahgeek/OpenCV-octVR; every function cites the reference
 * file:line it follows).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it.  The product
 * (opencv-octvr_b200/) never links, imports or executes anything in oracle/.
 *
 * Parity status: PINNED -- every function here is checked byte-for-byte (or
 * to the stated tolerance for f64 projection math) against golden vectors
 * produced by the unmodified reference CPU build (tests/golden/, generator
 * oracle/refgen/).
 */
#ifndef OCTVR_ORACLE_H
#define OCTVR_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- colour (modules/imgproc/src/color.cpp:6087-6169, 6430-6481) ---- */
/* planar 4:2:0 -> packed RGB888.  u/v given as pointer + pixel stride + row
 * stride so I420 (pix 1), NV12 (pix 2) and octvr's side-by-side U|V layout
 * (mapper.hpp:75-83) are all expressible. */
void orc_yuv420_to_rgb(const uint8_t* y, ptrdiff_t y_step,
                       const uint8_t* u, ptrdiff_t u_pix, ptrdiff_t u_step,
                       const uint8_t* v, ptrdiff_t v_pix, ptrdiff_t v_step,
                       int w, int h, uint8_t* rgb, ptrdiff_t rgb_step);
/* packed RGB888 -> planar 4:2:0 (chroma from the top-left pixel of each 2x2) */
void orc_rgb_to_yuv420(const uint8_t* rgb, ptrdiff_t rgb_step, int w, int h,
                       uint8_t* y, ptrdiff_t y_step,
                       uint8_t* u, ptrdiff_t u_pix, ptrdiff_t u_step,
                       uint8_t* v, ptrdiff_t v_pix, ptrdiff_t v_step);

/* ---- remap (modules/imgproc/src/imgwarp.cpp:120-300,3496-3560,3812-4020,4246-4480) ---- */
/* mapx/mapy are PIXEL coordinates (already map*W, map*H in f32, template.cpp:175-176).
 * interp: 0 = INTER_NEAREST, 1 = INTER_LINEAR.  BORDER_CONSTANT 0.  cn = 1,3 or 4. */
void orc_remap_u8(const uint8_t* src, ptrdiff_t src_step, int sw, int sh, int cn,
                  const float* mapx, const float* mapy, ptrdiff_t map_step_elems,
                  int dw, int dh, uint8_t* dst, ptrdiff_t dst_step, int interp);
/* out = fl32(map * scale) exactly as Mat*double -> convertTo does (template.cpp:175) */
void orc_scale_map(const float* in, size_t n, int scale, float* out);

/* ---- resize (modules/imgproc/src/imgwarp.cpp:3224-3500, 1387-1500) ---- */
void orc_resize_nn_u8(const uint8_t* src, ptrdiff_t src_step, int sw, int sh, int cn,
                      uint8_t* dst, ptrdiff_t dst_step, int dw, int dh);
void orc_resize_linear_u8(const uint8_t* src, ptrdiff_t src_step, int sw, int sh, int cn,
                          uint8_t* dst, ptrdiff_t dst_step, int dw, int dh);
void orc_resize_linear_f32(const float* src, ptrdiff_t src_step_elems, int sw, int sh,
                           float* dst, ptrdiff_t dst_step_elems, int dw, int dh);

/* ---- distance transform DIST_L2 3x3 (modules/imgproc/src/distransform.cpp:47-139) ---- */
void orc_dist_l2_3x3(const uint8_t* mask, ptrdiff_t mask_step, int w, int h,
                     float* dist, ptrdiff_t dist_step_elems);

/* ---- gain compensation (modules/stitching/src/exposure_compensate.cpp:82-156,329-332) ---- */
/* images: n packed RGB888 working-scale images, masks u8, corners/sizes per image.
 * Writes n gains.  Returns 0 on success. */
int orc_gain_feed(int n, const uint8_t* const* imgs, const ptrdiff_t* img_steps,
                  const uint8_t* const* masks, const ptrdiff_t* mask_steps,
                  const int* corners_xy, const int* sizes_wh, double* gains);
/* img = saturate_u8(rint(img * g)) in f64 (cv::multiply by scalar, arithm.cpp:578-760) */
void orc_mul_scalar_u8(uint8_t* img, ptrdiff_t step, int w_bytes, int h, double g);

/* ---- feather blend, octvr semantics (modules/stitching/src/blenders.cpp:531-586,
 *      src/cuda/blender.cu:73-98; CPU twin of the weight recipe: apps/octvr/monkey_gen.cpp:44-65) ---- */
/* weights[i] (f32, roi_i size) from masks[i]; rois = x,y,w,h per image. */
void orc_feather_weights(int n, const uint8_t* const* masks, const int* rois_xywh,
                         int border, float* const* weights);
/* imgs: RGB888 per ROI; out: RGB888 over the result ROI (union), zero where uncovered. */
void orc_feather_blend(int n, const uint8_t* const* imgs, const float* const* weights,
                       const int* rois_xywh, uint8_t* out, ptrdiff_t out_step,
                       int out_x, int out_y, int out_w, int out_h);

/* ---- pyramids (modules/imgproc/src/pyramids.cpp:849-1060) ---- */
void orc_pyrdown_s16(const int16_t* src, int sw, int sh, int cn, int16_t* dst); /* dst ((sw+1)/2,(sh+1)/2) */
void orc_pyrup_s16(const int16_t* src, int sw, int sh, int cn, int16_t* dst);   /* dst (2sw,2sh) */
void orc_pyrdown_f32(const float* src, int sw, int sh, float* dst);

/* ---- CPU MultiBandBlender (modules/stitching/src/blenders.cpp:221-477,764-933), weight_type CV_32F ---- */
/* imgs: RGB888 per ROI (converted to 16S inside, as test_blenders.cpp:53-55 feeds 16S);
 * masks u8 per ROI (octvr passes seam_masks, mapper.cpp:163).  out RGB888 over the union ROI. */
int orc_multiband_blend(int n, const uint8_t* const* imgs, const uint8_t* const* masks,
                        const int* rois_xywh, int num_bands,
                        uint8_t* out, ptrdiff_t out_step, uint8_t* out_mask, ptrdiff_t out_mask_step);

/* ---- seam masks (modules/stitching/src/seam_finders.cpp:86-133, octvr template.cpp:155-204) ---- */
/* masks: full-res u8 masks per ROI; writes seam masks (same sizes). */
void orc_seam_masks(int n, const uint8_t* const* masks, const int* rois_xywh,
                    int out_w, int out_h, uint8_t* const* seam_masks);

/* ---- camera models + template map generation (modules/octvr/src/camera.cpp, cameras/ *, template.cpp:46-153) ---- */
typedef struct orc_camera {
    int type;                /* ORC_CAM_* */
    double rot[9];           /* rotate_matrix (row-major), camera.cpp:49-73 */
    double min_lon, max_lon; /* longitude_selection */
    /* model parameters */
    double p[16];
    int    ip[8];
    /* ocam */
    double pol[64], invpol[64];
    int n_pol, n_invpol;
    /* pinhole/fisheye */
    double dist[14];
    int n_dist;
    /* optional exclude / include masks (camera.cpp:75-135) */
    const uint8_t* exclude_mask; int ex_w, ex_h;
    const uint8_t* include_mask; int in_w, in_h;
} orc_camera;

enum { ORC_CAM_NORMAL = 0, ORC_CAM_PERSPECTIVE, ORC_CAM_PINHOLE, ORC_CAM_FISHEYE,
       ORC_CAM_EQUIRECT, ORC_CAM_FULLFRAME_FISHEYE, ORC_CAM_OCAM, ORC_CAM_STUPIDOVAL,
       ORC_CAM_CUBIC, ORC_CAM_EQAREA_NORTH, ORC_CAM_EQAREA_SOUTH };

/* camera.cpp:52-64 : rotation {roll,yaw,pitch} -> matrix */
void orc_rotation_matrix(double roll, double yaw, double pitch, double* R9);
/* fullframe_fisheye_cam.cpp:89-103 */
double orc_fisheye_correction_radius(const double* coeff4);
double orc_camera_aspect_ratio(const orc_camera* cam);

/* One add_input (template.cpp:46-153): full-size map1,map2 (f32, normalised, -1 masked),
 * mask u8, visible (u8 W*H, in/out: template.cpp's visible_mask), roi_xywh out.
 * prior_masks/prior_rois: masks of previously added inputs (cleared where include mask hits).
 * Returns 0, or -1 when the model cannot be used in that direction. */
int orc_template_add_input(const orc_camera* out_cam, const orc_camera* in_cam,
                           int W, int H, float* map1, float* map2, uint8_t* mask,
                           uint8_t* visible, int use_roi, int* roi_xywh,
                           int n_prior, uint8_t* const* prior_masks, const int* prior_rois_xywh);

/* vignette.cpp:39-54 */
void orc_vignette_map(const float abcd[4], int width, int height, float* out);

int orc_num_threads(void);
void orc_set_num_threads(int n);

/* cv::fillPoly, one contour, 8UC1, lineType 8 (drawing.cpp:80-265,1195-1404); pts = {x0,y0,x1,y1,...} */
void orc_fill_poly(uint8_t* img, ptrdiff_t step, int w, int h, const int* pts, int npts, int val);

#ifdef __cplusplus
}
#endif
#endif
