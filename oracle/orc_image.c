/*
 * oracle/orc_image.c -- CPU ORACLE (test infrastructure, never shipped):
 * colour conversion, remap, resize, distance transform.
 * Restates the reference's CPU arithmetic; citations are to /root/reference.
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <omp.h>

static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
static inline int16_t sat_s16(int v) { return (int16_t)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* ------------------------------------------------------------------ colour */
/* modules/imgproc/src/color.cpp:6087-6094 (YUV->RGB) and :6096-6104 (RGB->YUV) */
enum { CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527, SHIFT = 20 };
enum { CRY = 269484, CGY = 528482, CBY = 102760, CRU = -155188, CGU = -305135, CBU = 460324,
       CGV = -385875, CBV = -74448 };

/* color.cpp:6121-6169 (semi-planar) and the planar twin that follows it: one chroma
 * sample per 2x2 block, luma floor at 16, +2^19 rounding, >>20, saturate. */
void orc_yuv420_to_rgb(const uint8_t* y, ptrdiff_t y_step,
                       const uint8_t* u, ptrdiff_t u_pix, ptrdiff_t u_step,
                       const uint8_t* v, ptrdiff_t v_pix, ptrdiff_t v_step,
                       int w, int h, uint8_t* rgb, ptrdiff_t rgb_step)
{
    #pragma omp parallel for schedule(static)
    for (int j = 0; j < h; j += 2) {
        const uint8_t* y1 = y + (ptrdiff_t)j * y_step;
        const uint8_t* y2 = y1 + y_step;
        const uint8_t* ur = u + (ptrdiff_t)(j / 2) * u_step;
        const uint8_t* vr = v + (ptrdiff_t)(j / 2) * v_step;
        uint8_t* r1 = rgb + (ptrdiff_t)j * rgb_step;
        uint8_t* r2 = r1 + rgb_step;
        for (int i = 0; i < w; i += 2, r1 += 6, r2 += 6) {
            int uu = (int)ur[(i / 2) * u_pix] - 128;
            int vv = (int)vr[(i / 2) * v_pix] - 128;
            int ruv = (1 << (SHIFT - 1)) + CVR * vv;
            int guv = (1 << (SHIFT - 1)) + CVG * vv + CUG * uu;
            int buv = (1 << (SHIFT - 1)) + CUB * uu;
            int yy;
            yy = (y1[i] > 16 ? y1[i] - 16 : 0) * CY;
            r1[0] = sat_u8((yy + ruv) >> SHIFT); r1[1] = sat_u8((yy + guv) >> SHIFT); r1[2] = sat_u8((yy + buv) >> SHIFT);
            yy = (y1[i + 1] > 16 ? y1[i + 1] - 16 : 0) * CY;
            r1[3] = sat_u8((yy + ruv) >> SHIFT); r1[4] = sat_u8((yy + guv) >> SHIFT); r1[5] = sat_u8((yy + buv) >> SHIFT);
            yy = (y2[i] > 16 ? y2[i] - 16 : 0) * CY;
            r2[0] = sat_u8((yy + ruv) >> SHIFT); r2[1] = sat_u8((yy + guv) >> SHIFT); r2[2] = sat_u8((yy + buv) >> SHIFT);
            yy = (y2[i + 1] > 16 ? y2[i + 1] - 16 : 0) * CY;
            r2[3] = sat_u8((yy + ruv) >> SHIFT); r2[4] = sat_u8((yy + guv) >> SHIFT); r2[5] = sat_u8((yy + buv) >> SHIFT);
        }
    }
}

/* color.cpp:6442-6481: note the V row uses CBU as its R coefficient. */
void orc_rgb_to_yuv420(const uint8_t* rgb, ptrdiff_t rgb_step, int w, int h,
                       uint8_t* y, ptrdiff_t y_step,
                       uint8_t* u, ptrdiff_t u_pix, ptrdiff_t u_step,
                       uint8_t* v, ptrdiff_t v_pix, ptrdiff_t v_step)
{
    const int half = 1 << (SHIFT - 1), s16 = 16 << SHIFT, s128 = 128 << SHIFT;
    #pragma omp parallel for schedule(static)
    for (int i = 0; i < h / 2; i++) {
        const uint8_t* row0 = rgb + (ptrdiff_t)(2 * i) * rgb_step;
        const uint8_t* row1 = row0 + rgb_step;
        uint8_t* yr0 = y + (ptrdiff_t)(2 * i) * y_step;
        uint8_t* yr1 = yr0 + y_step;
        uint8_t* ur = u + (ptrdiff_t)i * u_step;
        uint8_t* vr = v + (ptrdiff_t)i * v_step;
        for (int k = 0; k < w / 2; k++) {
            const uint8_t* p00 = row0 + 6 * k, *p01 = p00 + 3, *p10 = row1 + 6 * k, *p11 = p10 + 3;
            yr0[2 * k]     = sat_u8((CRY * p00[0] + CGY * p00[1] + CBY * p00[2] + half + s16) >> SHIFT);
            yr0[2 * k + 1] = sat_u8((CRY * p01[0] + CGY * p01[1] + CBY * p01[2] + half + s16) >> SHIFT);
            yr1[2 * k]     = sat_u8((CRY * p10[0] + CGY * p10[1] + CBY * p10[2] + half + s16) >> SHIFT);
            yr1[2 * k + 1] = sat_u8((CRY * p11[0] + CGY * p11[1] + CBY * p11[2] + half + s16) >> SHIFT);
            ur[k * u_pix] = sat_u8((CRU * p00[0] + CGU * p00[1] + CBU * p00[2] + half + s128) >> SHIFT);
            vr[k * v_pix] = sat_u8((CBU * p00[0] + CGV * p00[1] + CBV * p00[2] + half + s128) >> SHIFT);
        }
    }
}

/* ------------------------------------------------------------------- remap */
/* imgwarp.cpp:131-135 (linear 1-D taps), :210-262 (2-D table, forced to sum to 32768) */
static int16_t g_bilin[32 * 32][4];
static int g_bilin_ready = 0;
static void init_bilin(void)
{
    if (g_bilin_ready) return;
    float t1[32][2];
    for (int i = 0; i < 32; i++) { float x = i * (1.f / 32); t1[i][0] = 1.f - x; t1[i][1] = x; }
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 32; j++) {
            int16_t* it = g_bilin[i * 32 + j];
            int isum = 0;
            for (int k1 = 0; k1 < 2; k1++)
                for (int k2 = 0; k2 < 2; k2++) {
                    float vv = t1[i][k1] * t1[j][k2];
                    it[k1 * 2 + k2] = sat_s16((int)lrintf(vv * 32768.f));
                    isum += it[k1 * 2 + k2];
                }
            /* imgwarp.cpp:238-262: when the four entries do not sum to 32768 the difference goes to
             * the largest entry found scanning from (ksize/2, ksize/2) = index 3 (for ksize = 2 that scan
             * only ever sees index 3 and the still-zero start of the next cell).  This happens exactly
             * once: cell (0,0), where 1.0*32768 saturates to 32767 -> {32767, 0, 0, 1}.  The pixel
             * result is unchanged: (32767*a + d + 16384) >> 15 == a for 8-bit a, d. */
            if (isum != 32768) {
                int diff = isum - 32768;
                if (diff > 0) abort();
                it[3] = (int16_t)(it[3] - diff);
            }
        }
    g_bilin_ready = 1;
}

const int16_t* orc_bilinear_table(void) { init_bilin(); return &g_bilin[0][0]; }

void orc_scale_map(const float* in, size_t n, int scale, float* out)
{
    /* Mat * double -> convertTo(f32, alpha): dst = src*(float)alpha + 0.f (convert.cpp cvtScale_<float,float,float>) */
    const float a = (float)(double)scale;
    for (size_t i = 0; i < n; i++) out[i] = in[i] * a + 0.f;
}

void orc_remap_u8(const uint8_t* src, ptrdiff_t src_step, int sw, int sh, int cn,
                  const float* mapx, const float* mapy, ptrdiff_t map_step,
                  int dw, int dh, uint8_t* dst, ptrdiff_t dst_step, int interp)
{
    init_bilin();
    #pragma omp parallel for schedule(static)
    for (int yy = 0; yy < dh; yy++) {
        const float* mx = mapx + (ptrdiff_t)yy * map_step;
        const float* my = mapy + (ptrdiff_t)yy * map_step;
        uint8_t* d = dst + (ptrdiff_t)yy * dst_step;
        for (int xx = 0; xx < dw; xx++, d += cn) {
            if (interp == 0) {
                /* imgwarp.cpp:4307-4343 (cvRound + saturate to short), :3496-3560 */
                int sx = sat_s16((int)lrintf(mx[xx])), sy = sat_s16((int)lrintf(my[xx]));
                if ((unsigned)sx < (unsigned)sw && (unsigned)sy < (unsigned)sh) {
                    const uint8_t* s = src + (ptrdiff_t)sy * src_step + sx * cn;
                    for (int c = 0; c < cn; c++) d[c] = s[c];
                } else
                    for (int c = 0; c < cn; c++) d[c] = 0;
                continue;
            }
            /* imgwarp.cpp:4383-4442: 1/32-px fixed point coordinates */
            int fsx = (int)lrintf(mx[xx] * 32.f), fsy = (int)lrintf(my[xx] * 32.f);
            int sx = sat_s16(fsx >> 5), sy = sat_s16(fsy >> 5);
            const int16_t* w = g_bilin[(fsy & 31) * 32 + (fsx & 31)];
            /* imgwarp.cpp:3812-4020 : taps outside the image contribute cval = 0 */
            int in_x0 = (unsigned)sx < (unsigned)sw, in_x1 = (unsigned)(sx + 1) < (unsigned)sw;
            int in_y0 = (unsigned)sy < (unsigned)sh, in_y1 = (unsigned)(sy + 1) < (unsigned)sh;
            const uint8_t* s0 = src + (ptrdiff_t)sy * src_step + sx * cn;
            const uint8_t* s1 = s0 + src_step;
            for (int c = 0; c < cn; c++) {
                int v00 = (in_x0 && in_y0) ? s0[c] : 0, v01 = (in_x1 && in_y0) ? s0[c + cn] : 0;
                int v10 = (in_x0 && in_y1) ? s1[c] : 0, v11 = (in_x1 && in_y1) ? s1[c + cn] : 0;
                d[c] = sat_u8((v00 * w[0] + v01 * w[1] + v10 * w[2] + v11 * w[3] + (1 << 14)) >> 15);
            }
        }
    }
}

/* ------------------------------------------------------------------ resize */
/* imgwarp.cpp resizeNN: x_ofs = min(cvFloor(x*ifx), sw-1), ifx = 1/inv_scale_x */
void orc_resize_nn_u8(const uint8_t* src, ptrdiff_t src_step, int sw, int sh, int cn,
                      uint8_t* dst, ptrdiff_t dst_step, int dw, int dh)
{
    double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    double ifx = 1. / inv_x, ify = 1. / inv_y;
    for (int y = 0; y < dh; y++) {
        int sy = (int)floor(y * ify); if (sy > sh - 1) sy = sh - 1;
        const uint8_t* s = src + (ptrdiff_t)sy * src_step;
        uint8_t* d = dst + (ptrdiff_t)y * dst_step;
        for (int x = 0; x < dw; x++) {
            int sx = (int)floor(x * ifx); if (sx > sw - 1) sx = sw - 1;
            for (int c = 0; c < cn; c++) d[x * cn + c] = s[sx * cn + c];
        }
    }
}

/* imgwarp.cpp:3224-3500 coefficient setup; :1387-1419 horizontal; :1477-1500 vertical.
 * 11-bit coefficients; vertical pass uses the odd ((b*(S>>4))>>16 ... +2)>>2 form. */
void orc_resize_linear_u8(const uint8_t* src, ptrdiff_t src_step, int sw, int sh, int cn,
                          uint8_t* dst, ptrdiff_t dst_step, int dw, int dh)
{
    double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
    int* xofs = (int*)malloc(sizeof(int) * dw);
    short* ialpha = (short*)malloc(sizeof(short) * 2 * dw);
    int* yofs = (int*)malloc(sizeof(int) * dh);
    short* ibeta = (short*)malloc(sizeof(short) * 2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = sat_s16((int)lrintf((1.f - fx) * 2048));
        ialpha[2 * dx + 1] = sat_s16((int)lrintf(fx * 2048));
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[2 * dy] = sat_s16((int)lrintf((1.f - fy) * 2048));
        ibeta[2 * dy + 1] = sat_s16((int)lrintf(fy * 2048));
    }
    int* r0 = (int*)malloc(sizeof(int) * dw * cn), *r1 = (int*)malloc(sizeof(int) * dw * cn);
    for (int dy = 0; dy < dh; dy++) {
        int sy0 = yofs[dy], sy1 = sy0 + 1;
        if (sy0 < 0) sy0 = 0; if (sy0 > sh - 1) sy0 = sh - 1;
        if (sy1 < 0) sy1 = 0; if (sy1 > sh - 1) sy1 = sh - 1;
        const uint8_t* S0 = src + (ptrdiff_t)sy0 * src_step, *S1 = src + (ptrdiff_t)sy1 * src_step;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
            int sx1 = sx + 1 < sw ? sx + 1 : sx; /* a1 == 0 whenever sx is the last column */
            for (int c = 0; c < cn; c++) {
                r0[dx * cn + c] = S0[sx * cn + c] * a0 + S0[sx1 * cn + c] * a1;
                r1[dx * cn + c] = S1[sx * cn + c] * a0 + S1[sx1 * cn + c] * a1;
            }
        }
        int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
        uint8_t* d = dst + (ptrdiff_t)dy * dst_step;
        for (int x = 0; x < dw * cn; x++)
            d[x] = (uint8_t)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(ialpha); free(yofs); free(ibeta); free(r0); free(r1);
}

/* f32 bilinear resize (same coefficient setup, float taps; HResizeLinear<float,float,float,1> +
 * VResizeLinear<float,...,Cast<float,float>>): used for the vignette map (mapper.cpp:108-112). */
void orc_resize_linear_f32(const float* src, ptrdiff_t src_step, int sw, int sh,
                           float* dst, ptrdiff_t dst_step, int dw, int dh)
{
    double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
    int* xofs = (int*)malloc(sizeof(int) * dw);
    float* alpha = (float*)malloc(sizeof(float) * 2 * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx; alpha[2 * dx] = 1.f - fx; alpha[2 * dx + 1] = fx;
    }
    float* r0 = (float*)malloc(sizeof(float) * dw), *r1 = (float*)malloc(sizeof(float) * dw);
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        int sy0 = sy, sy1 = sy + 1;
        if (sy0 < 0) sy0 = 0; if (sy0 > sh - 1) sy0 = sh - 1;
        if (sy1 < 0) sy1 = 0; if (sy1 > sh - 1) sy1 = sh - 1;
        const float* S0 = src + (ptrdiff_t)sy0 * src_step, *S1 = src + (ptrdiff_t)sy1 * src_step;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sx;
            r0[dx] = S0[sx] * alpha[2 * dx] + S0[sx1] * alpha[2 * dx + 1];
            r1[dx] = S1[sx] * alpha[2 * dx] + S1[sx1] * alpha[2 * dx + 1];
        }
        float b0 = 1.f - fy, b1 = fy;
        float* d = dst + (ptrdiff_t)dy * dst_step;
        for (int x = 0; x < dw; x++) d[x] = r0[x] * b0 + r1[x] * b1;
    }
    free(xofs); free(alpha); free(r0); free(r1);
}

/* ------------------------------------------------------ distance transform */
/* distransform.cpp:69-139: two-pass 3x3 chamfer, 16.16 fixed point,
 * HV = cvRound(0.955f*65536), DIAG = cvRound(1.3693f*65536), outside = INT_MAX>>2. */
void orc_dist_l2_3x3(const uint8_t* mask, ptrdiff_t mask_step, int w, int h,
                     float* dist, ptrdiff_t dist_step)
{
    const int HV = (int)lrint(0.955f * (1 << 16)), DG = (int)lrint(1.3693f * (1 << 16));
    const int INIT = INT_MAX >> 2;
    const float scale = 1.f / (1 << 16);
    ptrdiff_t step = w + 2;
    int* temp = (int*)malloc(sizeof(int) * step * (h + 2));
    for (ptrdiff_t j = 0; j < step; j++) { temp[j] = INIT; temp[(h + 1) * step + j] = INIT; }
    for (int i = 0; i < h; i++) {
        const uint8_t* s = mask + (ptrdiff_t)i * mask_step;
        int* tmp = temp + (i + 1) * step + 1;
        tmp[-1] = tmp[w] = INIT;
        for (int j = 0; j < w; j++) {
            if (!s[j]) tmp[j] = 0;
            else {
                int t0 = tmp[j - step - 1] + DG, t = tmp[j - step] + HV;
                if (t0 > t) t0 = t;
                t = tmp[j - step + 1] + DG; if (t0 > t) t0 = t;
                t = tmp[j - 1] + HV; if (t0 > t) t0 = t;
                tmp[j] = t0;
            }
        }
    }
    for (int i = h - 1; i >= 0; i--) {
        float* d = dist + (ptrdiff_t)i * dist_step;
        int* tmp = temp + (i + 1) * step + 1;
        for (int j = w - 1; j >= 0; j--) {
            int t0 = tmp[j];
            if (t0 > HV) {
                int t = tmp[j + step + 1] + DG; if (t0 > t) t0 = t;
                t = tmp[j + step] + HV; if (t0 > t) t0 = t;
                t = tmp[j + step - 1] + DG; if (t0 > t) t0 = t;
                t = tmp[j + 1] + HV; if (t0 > t) t0 = t;
                tmp[j] = t0;
            }
            d[j] = (float)(t0 * scale);
        }
    }
    free(temp);
}


/* ------------------------------------------------------------------------------------------------
 * cv::fillPoly(img, {pts}, val): one contour, 8UC1, lineType 8, shift 0 (imgproc/src/drawing.cpp).
 *   outline : CollectPolyEdges :1195-1248 -> Line :238-265 -> LineIterator :142-236 (8-connected, left to
 *             right) after clipLine :80-136
 *   interior: FillEdgeCollection :1261-1404, 16.16 fixed-point edge table, active list kept sorted by x
 * Camera::Camera draws the selection rectangle and the polygonal exclude / include masks with it
 * (modules/octvr/src/camera.cpp:96-167).  Pinned by tests/golden/fillpoly.npz (reference build).
 * ------------------------------------------------------------------------------------------------ */
static int orc_clip_line(int w, int h, int* ax, int* ay, int* bx, int* by)
{
    long long right = w - 1, bottom = h - 1, x1 = *ax, y1 = *ay, x2 = *bx, y2 = *by, a;
    int c1, c2;
    if (w <= 0 || h <= 0) return 0;
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        if (c1 & 12) { a = c1 < 8 ? 0 : bottom; x1 += (a - y1) * (x2 - x1) / (y2 - y1); y1 = a; c1 = (x1 < 0) + (x1 > right) * 2; }
        if (c2 & 12) { a = c2 < 8 ? 0 : bottom; x2 += (a - y2) * (x2 - x1) / (y2 - y1); y2 = a; c2 = (x2 < 0) + (x2 > right) * 2; }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) { a = c1 == 1 ? 0 : right; y1 += (a - x1) * (y2 - y1) / (x2 - x1); x1 = a; c1 = 0; }
            if (c2) { a = c2 == 1 ? 0 : right; y2 += (a - x2) * (y2 - y1) / (x2 - x1); x2 = a; c2 = 0; }
        }
        *ax = (int)x1; *ay = (int)y1; *bx = (int)x2; *by = (int)y2;
    }
    return (c1 | c2) == 0;
}

static void orc_line8(uint8_t* img, ptrdiff_t step, int w, int h, int ax, int ay, int bx, int by, uint8_t val)
{
    int dx, dy, sy, steep, major, minor, err, i, x, y;
    if ((unsigned)ax >= (unsigned)w || (unsigned)bx >= (unsigned)w || (unsigned)ay >= (unsigned)h || (unsigned)by >= (unsigned)h)
        if (!orc_clip_line(w, h, &ax, &ay, &bx, &by)) return;
    dx = bx - ax; dy = by - ay;
    if (dx < 0) { dx = -dx; dy = -dy; ax = bx; ay = by; }
    sy = dy < 0 ? -1 : 1;
    if (dy < 0) dy = -dy;
    steep = dy > dx;
    major = steep ? dy : dx; minor = steep ? dx : dy;
    err = major - 2 * minor;
    x = ax; y = ay;
    for (i = 0; i <= major; i++) {
        int both = err < 0;
        img[(ptrdiff_t)y * step + x] = val;
        err += -2 * minor + (both ? 2 * major : 0);
        if (steep) { y += sy; x += both; } else { x += 1; if (both) y += sy; }
    }
}

typedef struct OrcEdge { int y0, y1, x, dx; struct OrcEdge* next; } OrcEdge;
static int orc_edge_cmp(const void* pa, const void* pb)
{
    const OrcEdge* a = (const OrcEdge*)pa; const OrcEdge* b = (const OrcEdge*)pb;
    if (a->y0 != b->y0) return a->y0 < b->y0 ? -1 : 1;
    if (a->x != b->x) return a->x < b->x ? -1 : 1;
    if (a->dx != b->dx) return a->dx < b->dx ? -1 : 1;
    return 0;
}

void orc_fill_poly(uint8_t* img, ptrdiff_t step, int w, int h, const int* pts, int npts, int val)
{
    OrcEdge* ed; OrcEdge head; OrcEdge* e;
    int total = 0, i, y, px, py, y_max = INT_MIN, x_max = INT_MIN, y_min = INT_MAX, x_min = INT_MAX;
    if (npts <= 0) return;
    ed = (OrcEdge*)malloc(sizeof(OrcEdge) * ((size_t)npts + 1));
    px = pts[2 * (npts - 1)] << 16; py = pts[2 * (npts - 1) + 1];
    for (i = 0; i < npts; i++) {
        int qx = pts[2 * i] << 16, qy = pts[2 * i + 1];
        orc_line8(img, step, w, h, (px + 32768) >> 16, py, (qx + 32768) >> 16, qy, (uint8_t)val);
        if (py != qy) {
            OrcEdge* n = &ed[total++];
            if (py < qy) { n->y0 = py; n->y1 = qy; n->x = px; } else { n->y0 = qy; n->y1 = py; n->x = qx; }
            n->dx = (qx - px) / (qy - py);
            n->next = 0;
        }
        px = qx; py = qy;
    }
    if (total < 2) { free(ed); return; }
    for (i = 0; i < total; i++) {
        int x1 = ed[i].x + (ed[i].y1 - ed[i].y0) * ed[i].dx;
        if (ed[i].y0 < y_min) y_min = ed[i].y0;
        if (ed[i].y1 > y_max) y_max = ed[i].y1;
        if (ed[i].x < x_min) x_min = ed[i].x;
        if (ed[i].x > x_max) x_max = ed[i].x;
        if (x1 < x_min) x_min = x1;
        if (x1 > x_max) x_max = x1;
    }
    if (y_max < 0 || y_min >= h || x_max < 0 || x_min >= (w << 16)) { free(ed); return; }
    /* std::sort with CmpEdges is not stable, but elements that compare equal are identical edges */
    qsort(ed, (size_t)total, sizeof(OrcEdge), orc_edge_cmp);
    head.y0 = INT_MAX; head.y1 = 0; head.x = 0; head.dx = 0; head.next = 0;
    ed[total] = head;
    i = 0; e = &ed[0];
    if (y_max > h) y_max = h;
    for (y = e->y0; y < y_max; y++) {
        OrcEdge *last = head.next, *prelast = &head, *keep;
        int sorted_something = 0, draw = 0, clip = y < 0;
        while (last || e->y0 == y) {
            if (last && last->y1 == y) { prelast->next = last->next; last = last->next; continue; }
            keep = prelast;
            if (last && (e->y0 > y || last->x < e->x)) { prelast = last; last = last->next; }
            else if (i < total) { prelast->next = e; e->next = last; prelast = e; e = &ed[++i]; }
            else break;
            if (draw) {
                if (!clip) {
                    int x1 = keep->x, x2 = prelast->x, x;
                    if (x1 > x2) { int t = x1; x1 = x2; x2 = t; }
                    x1 = (x1 + 65535) >> 16; x2 >>= 16;
                    if (x1 < w && x2 >= 0) {
                        if (x1 < 0) x1 = 0;
                        if (x2 >= w) x2 = w - 1;
                        for (x = x1; x <= x2; x++) img[(ptrdiff_t)y * step + x] = (uint8_t)val;
                    }
                }
                keep->x += keep->dx;
                prelast->x += prelast->dx;
            }
            draw ^= 1;
        }
        keep = 0;
        do {
            prelast = &head; last = head.next;
            while (last != keep && last->next != 0) {
                OrcEdge* te = last->next;
                if (last->x > te->x) { prelast->next = te; last->next = te->next; te->next = last; prelast = te; sorted_something = 1; }
                else { prelast = last; last = te; }
            }
            keep = prelast;
        } while (sorted_something && keep != head.next && keep != &head);
    }
    free(ed);
}
