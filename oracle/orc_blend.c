/*
 * oracle/orc_blend.c -- CPU ORACLE (test infrastructure, never shipped):
 * gain compensation, octvr feather blend, Gaussian/Laplacian pyramids,
 * CPU MultiBandBlender, DistanceSeamFinder seam masks.
 * Restates the reference's CPU arithmetic; citations are to /root/reference.
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
static inline int16_t sat_s16(int v) { return (int16_t)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------- gain */
/* core/src/matrix_decomp.cpp:52-109 (LU64f, eps = DBL_EPSILON*100) and lapack.cpp:1050-1170
 * (closed forms for n <= 3). A is n x n row-major, b length n; solution left in x. */
static int solve_like_cv(int n, double* A, double* b, double* x)
{
    if (n == 1) { if (A[0] == 0.) return -1; x[0] = b[0] / A[0]; return 0; } /* falls through to LU in cv; same value */
    if (n == 2) {
        double d = A[0] * A[3] - A[1] * A[2];
        if (d == 0.) return -1;
        d = 1. / d;
        double t = (b[0] * A[3] - b[1] * A[1]) * d;
        x[1] = (b[1] * A[0] - b[0] * A[2]) * d;
        x[0] = t;
        return 0;
    }
    if (n == 3) {
        #define S(i, j) A[(i) * 3 + (j)]
        double d = S(0,0) * (S(1,1) * S(2,2) - S(1,2) * S(2,1)) - S(0,1) * (S(1,0) * S(2,2) - S(1,2) * S(2,0)) +
                   S(0,2) * (S(1,0) * S(2,1) - S(1,1) * S(2,0));
        if (d == 0.) return -1;
        d = 1. / d;
        x[0] = ((S(1,1) * S(2,2) - S(1,2) * S(2,1)) * b[0] + (S(0,2) * S(2,1) - S(0,1) * S(2,2)) * b[1] +
                (S(0,1) * S(1,2) - S(0,2) * S(1,1)) * b[2]) * d;
        x[1] = ((S(1,2) * S(2,0) - S(1,0) * S(2,2)) * b[0] + (S(0,0) * S(2,2) - S(0,2) * S(2,0)) * b[1] +
                (S(0,2) * S(1,0) - S(0,0) * S(1,2)) * b[2]) * d;
        x[2] = ((S(1,0) * S(2,1) - S(1,1) * S(2,0)) * b[0] + (S(0,1) * S(2,0) - S(0,0) * S(2,1)) * b[1] +
                (S(0,0) * S(1,1) - S(0,1) * S(1,0)) * b[2]) * d;
        #undef S
        return 0;
    }
    const double eps = DBL_EPSILON * 100;
    for (int i = 0; i < n; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++)
            if (fabs(A[j * n + i]) > fabs(A[k * n + i])) k = j;
        if (fabs(A[k * n + i]) < eps) return -1;
        if (k != i) {
            for (int j = i; j < n; j++) { double t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = -1 / A[i * n + i];
        for (int j = i + 1; j < n; j++) {
            double alpha = A[j * n + i] * d;
            for (int kk = i + 1; kk < n; kk++) A[j * n + kk] += alpha * A[i * n + kk];
            b[j] += alpha * b[i];
        }
        A[i * n + i] = -d;
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= A[i * n + k] * b[k];
        b[i] = s * A[i * n + i];
    }
    for (int i = 0; i < n; i++) x[i] = b[i];
    return 0;
}

/* exposure_compensate.cpp:82-156.  Intersection rule: mask == 255 on both (ExposureCompensator::feed :71-78). */
int orc_gain_feed(int n, const uint8_t* const* imgs, const ptrdiff_t* img_steps,
                  const uint8_t* const* masks, const ptrdiff_t* mask_steps,
                  const int* corners, const int* sizes, double* gains)
{
    int* N = (int*)calloc((size_t)n * n, sizeof(int));
    double* I = (double*)calloc((size_t)n * n, sizeof(double));
    for (int i = 0; i < n; i++)
        for (int j = i; j < n; j++) {
            /* util.cpp overlapRoi */
            int x_tl = imax(corners[2 * i], corners[2 * j]), y_tl = imax(corners[2 * i + 1], corners[2 * j + 1]);
            int x_br = imin(corners[2 * i] + sizes[2 * i], corners[2 * j] + sizes[2 * j]);
            int y_br = imin(corners[2 * i + 1] + sizes[2 * i + 1], corners[2 * j + 1] + sizes[2 * j + 1]);
            if (!(x_tl < x_br && y_tl < y_br)) continue;
            int cnt = 0; double s1 = 0, s2 = 0;
            for (int y = y_tl; y < y_br; y++) {
                const uint8_t* r1 = imgs[i] + (ptrdiff_t)(y - corners[2 * i + 1]) * img_steps[i] + 3 * (x_tl - corners[2 * i]);
                const uint8_t* r2 = imgs[j] + (ptrdiff_t)(y - corners[2 * j + 1]) * img_steps[j] + 3 * (x_tl - corners[2 * j]);
                const uint8_t* m1 = masks[i] + (ptrdiff_t)(y - corners[2 * i + 1]) * mask_steps[i] + (x_tl - corners[2 * i]);
                const uint8_t* m2 = masks[j] + (ptrdiff_t)(y - corners[2 * j + 1]) * mask_steps[j] + (x_tl - corners[2 * j]);
                for (int x = 0; x < x_br - x_tl; x++)
                    if (m1[x] == 255 && m2[x] == 255) {
                        cnt++;
                        const uint8_t* a = r1 + 3 * x, *b = r2 + 3 * x;
                        s1 += sqrt((double)(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]));
                        s2 += sqrt((double)(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]));
                    }
            }
            N[i * n + j] = N[j * n + i] = imax(1, cnt);
            I[i * n + j] = s1 / N[i * n + j];
            I[j * n + i] = s2 / N[i * n + j];
        }
    const double alpha = 0.01, beta = 100;
    double* A = (double*)calloc((size_t)n * n, sizeof(double));
    double* b = (double*)calloc((size_t)n, sizeof(double));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            b[i] += beta * N[i * n + j];
            A[i * n + i] += beta * N[i * n + j];
            if (j == i) continue;
            A[i * n + i] += 2 * alpha * I[i * n + j] * I[i * n + j] * N[i * n + j];
            A[i * n + j] -= 2 * alpha * I[i * n + j] * I[j * n + i] * N[i * n + j];
        }
    int rc = solve_like_cv(n, A, b, gains);
    free(N); free(I); free(A); free(b);
    return rc;
}

/* GainCompensator::apply -> cv::multiply(image, double): working type f64, result cvRound + saturate
 * (exposure_compensate.cpp:329-332, core/src/arithm.cpp:578-760 with muldiv=true). */
void orc_mul_scalar_u8(uint8_t* img, ptrdiff_t step, int w_bytes, int h, double g)
{
    uint8_t lut[256];
    for (int v = 0; v < 256; v++) lut[v] = sat_u8((int)lrint((double)v * g));
    for (int y = 0; y < h; y++) {
        uint8_t* r = img + (ptrdiff_t)y * step;
        for (int x = 0; x < w_bytes; x++) r[x] = lut[r[x]];
    }
}

/* ---------------------------------------------------------------- feather */
static void union_roi(int n, const int* rois, int* out)
{
    int x0 = rois[0], y0 = rois[1], x1 = rois[0] + rois[2], y1 = rois[1] + rois[3];
    for (int i = 1; i < n; i++) {
        x0 = imin(x0, rois[4 * i]); y0 = imin(y0, rois[4 * i + 1]);
        x1 = imax(x1, rois[4 * i] + rois[4 * i + 2]); y1 = imax(y1, rois[4 * i + 1] + rois[4 * i + 3]);
    }
    out[0] = x0; out[1] = y0; out[2] = x1 - x0; out[3] = y1 - y0;
}

/* blenders.cpp:531-572 (FeatherGPUBlender ctor; CPU twins: apps/octvr/monkey_gen.cpp:44-65,
 * modules/octvr/src/mapper_fast.cpp:75-94):  w = max(DT - border, 0);  S = 1e-5f + sum_i w_i
 * (f32 adds in camera order);  W_i = (N * w_i) / S  (cudaarithm DivScaleOp: scale*a/b in f32). */
void orc_feather_weights(int n, const uint8_t* const* masks, const int* rois, int border,
                         float* const* weights)
{
    int R[4]; union_roi(n, rois, R);
    size_t area = (size_t)R[2] * R[3];
    float* S = (float*)malloc(sizeof(float) * area);
    for (size_t k = 0; k < area; k++) S[k] = 1e-5f;
    for (int i = 0; i < n; i++) {
        int x = rois[4 * i], y = rois[4 * i + 1], w = rois[4 * i + 2], h = rois[4 * i + 3];
        orc_dist_l2_3x3(masks[i], w, w, h, weights[i], w);
        for (int r = 0; r < h; r++) {
            float* wr = weights[i] + (size_t)r * w;
            float* sr = S + (size_t)(y - R[1] + r) * R[2] + (x - R[0]);
            for (int c = 0; c < w; c++) {
                float t = wr[c] - (float)border;
                wr[c] = t > 0.f ? t : 0.f;
                sr[c] = wr[c] + sr[c];
            }
        }
    }
    const float scale = (float)n;
    for (int i = 0; i < n; i++) {
        int x = rois[4 * i], y = rois[4 * i + 1], w = rois[4 * i + 2], h = rois[4 * i + 3];
        for (int r = 0; r < h; r++) {
            float* wr = weights[i] + (size_t)r * w;
            const float* sr = S + (size_t)(y - R[1] + r) * R[2] + (x - R[0]);
            for (int c = 0; c < w; c++) wr[c] = sr[c] != 0 ? scale * wr[c] / sr[c] : 0;
        }
    }
    free(S);
}

/* blenders.cpp:574-586 + cuda/blender.cu:73-98 (CPU twin: FeatherBlender::feed blenders.cpp:165-179):
 * acc(s16) += (short)(u8 * W) truncating, skipped where W == 0; out = sat_u8(rint(acc * (float)(1/N))). */
void orc_feather_blend(int n, const uint8_t* const* imgs, const float* const* weights,
                       const int* rois, uint8_t* out, ptrdiff_t out_step,
                       int ox, int oy, int ow, int oh)
{
    int16_t* acc = (int16_t*)calloc((size_t)ow * oh * 3, sizeof(int16_t));
    for (int i = 0; i < n; i++) {
        int x = rois[4 * i], y = rois[4 * i + 1], w = rois[4 * i + 2], h = rois[4 * i + 3];
        #pragma omp parallel for schedule(static)
        for (int r = 0; r < h; r++) {
            const uint8_t* s = imgs[i] + (size_t)r * w * 3;
            const float* wr = weights[i] + (size_t)r * w;
            int16_t* a = acc + ((size_t)(y - oy + r) * ow + (x - ox)) * 3;
            for (int c = 0; c < w; c++) {
                float ww = wr[c];
                if (ww == 0) continue;
                a[3 * c]     = (int16_t)(a[3 * c]     + (int16_t)(s[3 * c] * ww));
                a[3 * c + 1] = (int16_t)(a[3 * c + 1] + (int16_t)(s[3 * c + 1] * ww));
                a[3 * c + 2] = (int16_t)(a[3 * c + 2] + (int16_t)(s[3 * c + 2] * ww));
            }
        }
    }
    const float alpha = (float)(1.0 / n);
    #pragma omp parallel for schedule(static)
    for (int r = 0; r < oh; r++) {
        const int16_t* a = acc + (size_t)r * ow * 3;
        uint8_t* o = out + (ptrdiff_t)r * out_step;
        for (int c = 0; c < ow * 3; c++) o[c] = sat_u8((int)lrintf(a[c] * alpha + 0.f));
    }
    free(acc);
}

/* --------------------------------------------------------------- pyramids */
static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}
static inline int reflect(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p - 1; else p = 2 * len - 1 - p; }
    return p;
}

/* pyramids.cpp:849-964 with FixPtCast<short,8>: [1 4 6 4 1] both ways, (v+128)>>8, BORDER_REFLECT_101 */
void orc_pyrdown_s16(const int16_t* src, int sw, int sh, int cn, int16_t* dst)
{
    int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    #pragma omp parallel
    {
        int* rows = (int*)malloc(sizeof(int) * 5 * dw * cn);
        #pragma omp for schedule(static)
        for (int y = 0; y < dh; y++) {
            for (int k = 0; k < 5; k++) {
                int sy = reflect101(2 * y - 2 + k, sh);
                const int16_t* s = src + (size_t)sy * sw * cn;
                int* row = rows + (size_t)k * dw * cn;
                for (int x = 0; x < dw; x++) {
                    int x0 = reflect101(2 * x - 2, sw) * cn, x1 = reflect101(2 * x - 1, sw) * cn, x2 = 2 * x * cn,
                        x3 = reflect101(2 * x + 1, sw) * cn, x4 = reflect101(2 * x + 2, sw) * cn;
                    for (int c = 0; c < cn; c++)
                        row[x * cn + c] = s[x2 + c] * 6 + (s[x1 + c] + s[x3 + c]) * 4 + s[x0 + c] + s[x4 + c];
                }
            }
            int16_t* d = dst + (size_t)y * dw * cn;
            const int* r0 = rows, *r1 = rows + dw * cn, *r2 = r1 + dw * cn, *r3 = r2 + dw * cn, *r4 = r3 + dw * cn;
            for (int x = 0; x < dw * cn; x++)
                d[x] = sat_s16((r2[x] * 6 + (r1[x] + r3[x]) * 4 + r0[x] + r4[x] + 128) >> 8);
        }
        free(rows);
    }
}

/* pyramids.cpp:967-1060 with FixPtCast<short,6>.  dst is exactly (2sw, 2sh). */
static void pyrup_row_s16(const int16_t* s, int sw, int cn, int* row)
{
    if (sw == 1) { for (int c = 0; c < cn; c++) row[c] = row[c + cn] = s[c] * 8; return; }
    for (int c = 0; c < cn; c++) {
        row[c] = s[c] * 6 + s[c + cn] * 2;
        row[c + cn] = (s[c] + s[c + cn]) * 4;
        int sx = (sw - 1) * cn + c, dx = (sw - 1) * 2 * cn + c;
        row[dx] = s[sx - cn] + s[sx] * 7;
        row[dx + cn] = s[sx] * 8;
    }
    for (int x = 1; x < sw - 1; x++)
        for (int c = 0; c < cn; c++) {
            int sx = x * cn + c, dx = 2 * x * cn + c;
            row[dx] = s[sx - cn] + s[sx] * 6 + s[sx + cn];
            row[dx + cn] = (s[sx] + s[sx + cn]) * 4;
        }
}

void orc_pyrup_s16(const int16_t* src, int sw, int sh, int cn, int16_t* dst)
{
    int dw = 2 * sw, dh = 2 * sh;
    #pragma omp parallel
    {
        int* rows = (int*)malloc(sizeof(int) * 3 * dw * cn);
        #pragma omp for schedule(static)
        for (int y = 0; y < sh; y++) {
            for (int k = 0; k < 3; k++) {
                int sy = reflect101((y - 1 + k) * 2, dh) / 2;
                pyrup_row_s16(src + (size_t)sy * sw * cn, sw, cn, rows + (size_t)k * dw * cn);
            }
            const int* r0 = rows, *r1 = rows + dw * cn, *r2 = r1 + dw * cn;
            int16_t* d0 = dst + (size_t)(2 * y) * dw * cn, *d1 = d0 + (size_t)dw * cn;
            for (int x = 0; x < dw * cn; x++) {
                d1[x] = sat_s16(((r1[x] + r2[x]) * 4 + 32) >> 6);
                d0[x] = sat_s16((r0[x] + r1[x] * 6 + r2[x] + 32) >> 6);
            }
        }
        free(rows);
    }
}

/* pyramids.cpp:849-964 with FltCast<float,8>; the SSE vertical pass (:143-185, used for x < width&~7)
 * associates differently from the scalar tail -- both are reproduced. */
void orc_pyrdown_f32(const float* src, int sw, int sh, float* dst)
{
    int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    float* rows = (float*)malloc(sizeof(float) * 5 * dw);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            int sy = reflect101(2 * y - 2 + k, sh);
            const float* s = src + (size_t)sy * sw;
            float* row = rows + (size_t)k * dw;
            for (int x = 0; x < dw; x++) {
                int x0 = reflect101(2 * x - 2, sw), x1 = reflect101(2 * x - 1, sw), x2 = 2 * x,
                    x3 = reflect101(2 * x + 1, sw), x4 = reflect101(2 * x + 2, sw);
                row[x] = s[x2] * 6 + (s[x1] + s[x3]) * 4 + s[x0] + s[x4];
            }
        }
        const float* r0 = rows, *r1 = rows + dw, *r2 = r1 + dw, *r3 = r2 + dw, *r4 = r3 + dw;
        float* d = dst + (size_t)y * dw;
        int x = 0;
        for (; x <= dw - 8; x += 8)
            for (int k = 0; k < 8; k++) {
                float a = r0[x + k] + r4[x + k];
                float b = (r1[x + k] + r3[x + k]) + r2[x + k];
                a = a + (r2[x + k] + r2[x + k]);
                d[x + k] = (a + b * 4.f) * (1.f / 256);
            }
        for (; x < dw; x++)
            d[x] = (r2[x] * 6 + (r1[x] + r3[x]) * 4 + r0[x] + r4[x]) * (float)(1. / 256);
    }
    free(rows);
}

/* -------------------------------------------------------------- multiband */
/* blenders.cpp:237-477 (prepare/feed/blend), :764-825 (normalizeUsingWeightMap), :881-892
 * (createLaplacePyr 16S branch), :923-933 (restoreImageFromLaplacePyr).  weight_type CV_32F. */
int orc_multiband_blend(int n, const uint8_t* const* imgs, const uint8_t* const* masks,
                        const int* rois, int num_bands_req,
                        uint8_t* out, ptrdiff_t out_step, uint8_t* out_mask, ptrdiff_t out_mask_step)
{
    int Rf[4]; union_roi(n, rois, Rf);               /* dst_roi_final_ */
    double max_len = (double)imax(Rf[2], Rf[3]);
    int nb = imin(num_bands_req, (int)ceil(log(max_len) / log(2.0)));
    int al = 1 << nb;
    int R[4] = { Rf[0], Rf[1], Rf[2] + (al - Rf[2] % al) % al, Rf[3] + (al - Rf[3] % al) % al };

    int lw[32], lh[32];
    int16_t* dl[32]; float* dwt[32];
    lw[0] = R[2]; lh[0] = R[3];
    for (int i = 1; i <= nb; i++) { lw[i] = (lw[i - 1] + 1) / 2; lh[i] = (lh[i - 1] + 1) / 2; }
    for (int i = 0; i <= nb; i++) {
        dl[i] = (int16_t*)calloc((size_t)lw[i] * lh[i] * 3, sizeof(int16_t));
        dwt[i] = (float*)calloc((size_t)lw[i] * lh[i], sizeof(float));
    }

    for (int im = 0; im < n; im++) {
        int tlx = rois[4 * im], tly = rois[4 * im + 1], iw = rois[4 * im + 2], ih = rois[4 * im + 3];
        int gap = 3 * (1 << nb);
        int tnx = imax(R[0], tlx - gap), tny = imax(R[1], tly - gap);
        int bnx = imin(R[0] + R[2], tlx + iw + gap), bny = imin(R[1] + R[3], tly + ih + gap);
        tnx = R[0] + (((tnx - R[0]) >> nb) << nb);
        tny = R[1] + (((tny - R[1]) >> nb) << nb);
        int width = bnx - tnx, height = bny - tny;
        width += (al - width % al) % al;
        height += (al - height % al) % al;
        bnx = tnx + width; bny = tny + height;
        int dy = imax(bny - (R[1] + R[3]), 0), dx = imax(bnx - (R[0] + R[2]), 0);
        tnx -= dx; bnx -= dx; tny -= dy; bny -= dy;
        int top = tly - tny, left = tlx - tnx;

        /* bordered 16S image (BORDER_REFLECT) and f32 weight (BORDER_CONSTANT 0) */
        int16_t* pyr[32]; float* wp[32]; int pw[32], ph[32];
        pw[0] = width; ph[0] = height;
        for (int i = 1; i <= nb; i++) { pw[i] = (pw[i - 1] + 1) / 2; ph[i] = (ph[i - 1] + 1) / 2; }
        for (int i = 0; i <= nb; i++) {
            pyr[i] = (int16_t*)malloc(sizeof(int16_t) * (size_t)pw[i] * ph[i] * 3);
            wp[i] = (float*)malloc(sizeof(float) * (size_t)pw[i] * ph[i]);
        }
        const float inv255 = (float)(1. / 255.);
        #pragma omp parallel for schedule(static)
        for (int y = 0; y < height; y++) {
            int sy = reflect(y - top, ih);
            int inside_y = (y - top) >= 0 && (y - top) < ih;
            for (int x = 0; x < width; x++) {
                int sx = reflect(x - left, iw);
                const uint8_t* s = imgs[im] + ((size_t)sy * iw + sx) * 3;
                int16_t* d = pyr[0] + ((size_t)y * width + x) * 3;
                d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
                int inside = inside_y && (x - left) >= 0 && (x - left) < iw;
                wp[0][(size_t)y * width + x] = inside ? masks[im][(size_t)(y - top) * iw + (x - left)] * inv255 + 0.f : 0.f;
            }
        }
        for (int i = 0; i < nb; i++) {
            orc_pyrdown_s16(pyr[i], pw[i], ph[i], 3, pyr[i + 1]);
            orc_pyrdown_f32(wp[i], pw[i], ph[i], wp[i + 1]);
        }
        int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * (size_t)width * height * 3);
        for (int i = 0; i < nb; i++) {
            orc_pyrup_s16(pyr[i + 1], pw[i + 1], ph[i + 1], 3, tmp);
            size_t cnt = (size_t)pw[i] * ph[i] * 3;
            for (size_t k = 0; k < cnt; k++) pyr[i][k] = sat_s16(pyr[i][k] - tmp[k]);
        }
        free(tmp);

        int y_tl = tny - R[1], y_br = bny - R[1], x_tl = tnx - R[0], x_br = bnx - R[0];
        for (int i = 0; i <= nb; i++) {
            int rw = x_br - x_tl, rh = y_br - y_tl;
            for (int y = 0; y < rh; y++) {
                const int16_t* s = pyr[i] + (size_t)y * pw[i] * 3;
                const float* w = wp[i] + (size_t)y * pw[i];
                int16_t* d = dl[i] + ((size_t)(y_tl + y) * lw[i] + x_tl) * 3;
                float* dwr = dwt[i] + (size_t)(y_tl + y) * lw[i] + x_tl;
                for (int x = 0; x < rw; x++) {
                    d[3 * x]     = (int16_t)(d[3 * x]     + (int16_t)(s[3 * x] * w[x]));
                    d[3 * x + 1] = (int16_t)(d[3 * x + 1] + (int16_t)(s[3 * x + 1] * w[x]));
                    d[3 * x + 2] = (int16_t)(d[3 * x + 2] + (int16_t)(s[3 * x + 2] * w[x]));
                    dwr[x] += w[x];
                }
            }
            x_tl /= 2; y_tl /= 2; x_br /= 2; y_br /= 2;
        }
        for (int i = 0; i <= nb; i++) { free(pyr[i]); free(wp[i]); }
    }

    const float WEIGHT_EPS = 1e-5f;
    for (int i = 0; i <= nb; i++) {
        size_t cnt = (size_t)lw[i] * lh[i];
        for (size_t k = 0; k < cnt; k++) {
            float den = dwt[i][k] + WEIGHT_EPS;
            dl[i][3 * k]     = (int16_t)(dl[i][3 * k] / den);
            dl[i][3 * k + 1] = (int16_t)(dl[i][3 * k + 1] / den);
            dl[i][3 * k + 2] = (int16_t)(dl[i][3 * k + 2] / den);
        }
    }
    {
        int16_t* tmp = (int16_t*)malloc(sizeof(int16_t) * (size_t)lw[0] * lh[0] * 3);
        for (int i = nb; i > 0; i--) {
            orc_pyrup_s16(dl[i], lw[i], lh[i], 3, tmp);
            size_t cnt = (size_t)lw[i - 1] * lh[i - 1] * 3;
            for (size_t k = 0; k < cnt; k++) dl[i - 1][k] = sat_s16(tmp[k] + dl[i - 1][k]);
        }
        free(tmp);
    }
    for (int y = 0; y < Rf[3]; y++) {
        const int16_t* s = dl[0] + (size_t)y * lw[0] * 3;
        const float* w = dwt[0] + (size_t)y * lw[0];
        uint8_t* o = out + (ptrdiff_t)y * out_step;
        for (int x = 0; x < Rf[2]; x++) {
            int on = w[x] > WEIGHT_EPS;
            if (out_mask) out_mask[(ptrdiff_t)y * out_mask_step + x] = on ? 255 : 0;
            o[3 * x]     = on ? sat_u8(s[3 * x]) : 0;
            o[3 * x + 1] = on ? sat_u8(s[3 * x + 1]) : 0;
            o[3 * x + 2] = on ? sat_u8(s[3 * x + 2]) : 0;
        }
    }
    for (int i = 0; i <= nb; i++) { free(dl[i]); free(dwt[i]); }
    return nb;
}

/* ------------------------------------------------------------- seam masks */
/* octvr template.cpp:155-204 (create_masks, no images) + seam_finders.cpp:86-133
 * (DistanceSeamFinder, max_n = 1; full-width masks use the 3x-tiled wrap-around DT). */
void orc_seam_masks(int n, const uint8_t* const* masks, const int* rois, int out_w, int out_h,
                    uint8_t* const* seam_masks)
{
    (void)out_h;
    double scale = 960.0 / out_w; if (scale > 1.0) scale = 1.0;
    int* sr = (int*)malloc(sizeof(int) * 4 * n);
    uint8_t** um = (uint8_t**)malloc(sizeof(uint8_t*) * n);
    float** dist = (float**)malloc(sizeof(float*) * n);
    for (int i = 0; i < n; i++) {
        sr[4 * i] = (int)(rois[4 * i] * scale); sr[4 * i + 1] = (int)(rois[4 * i + 1] * scale);
        sr[4 * i + 2] = (int)(rois[4 * i + 2] * scale); sr[4 * i + 3] = (int)(rois[4 * i + 3] * scale);
        um[i] = (uint8_t*)malloc((size_t)sr[4 * i + 2] * sr[4 * i + 3]);
        orc_resize_linear_u8(masks[i], rois[4 * i + 2], rois[4 * i + 2], rois[4 * i + 3], 1,
                             um[i], sr[4 * i + 2], sr[4 * i + 2], sr[4 * i + 3]);
    }
    int R[4]; union_roi(n, sr, R);
    for (int i = 0; i < n; i++) {
        int w = sr[4 * i + 2], h = sr[4 * i + 3];
        dist[i] = (float*)malloc(sizeof(float) * (size_t)w * h);
        if (sr[4 * i] == 0 && w == R[2]) {
            uint8_t* t = (uint8_t*)malloc((size_t)3 * w * h);
            float* td = (float*)malloc(sizeof(float) * (size_t)3 * w * h);
            for (int y = 0; y < h; y++)
                for (int k = 0; k < 3; k++) memcpy(t + (size_t)y * 3 * w + (size_t)k * w, um[i] + (size_t)y * w, w);
            orc_dist_l2_3x3(t, 3 * w, 3 * w, h, td, 3 * w);
            for (int y = 0; y < h; y++) memcpy(dist[i] + (size_t)y * w, td + (size_t)y * 3 * w + w, sizeof(float) * w);
            free(t); free(td);
        } else
            orc_dist_l2_3x3(um[i], w, w, h, dist[i], w);
    }
    for (int y = R[1]; y < R[1] + R[3]; y++)
        for (int x = R[0]; x < R[0] + R[2]; x++) {
            /* winner = largest distance; libstdc++ std::sort on <= 16 elements is an insertion sort,
             * so ties keep the lower camera index first (seam_finders.cpp:126). */
            int best = -1; float bd = 0;
            for (int k = 0; k < n; k++) {
                int lx = x - sr[4 * k], ly = y - sr[4 * k + 1];
                float d = (lx >= 0 && ly >= 0 && ly < sr[4 * k + 3] && lx < sr[4 * k + 2]) ? dist[k][(size_t)ly * sr[4 * k + 2] + lx] : -1.f;
                if (best < 0 || d > bd) { best = k; bd = d; }
            }
            for (int k = 0; k < n; k++) {
                if (k == best) continue;
                int lx = x - sr[4 * k], ly = y - sr[4 * k + 1];
                if (lx >= 0 && ly >= 0 && ly < sr[4 * k + 3] && lx < sr[4 * k + 2])
                    um[k][(size_t)ly * sr[4 * k + 2] + lx] = 0;
            }
        }
    for (int i = 0; i < n; i++) {
        orc_resize_linear_u8(um[i], sr[4 * i + 2], sr[4 * i + 2], sr[4 * i + 3], 1,
                             seam_masks[i], rois[4 * i + 2], rois[4 * i + 2], rois[4 * i + 3]);
        free(um[i]); free(dist[i]);
    }
    free(sr); free(um); free(dist);
}
