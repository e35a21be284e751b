/*
 * oracle/orc_camera.c -- CPU ORACLE (test infrastructure, never shipped):
 * octvr camera projection models and MapperTemplate::add_input map generation.
 * Restates modules/octvr/src/camera.cpp, cameras/ *, template.cpp:46-153, vignette.cpp
 * in f64 exactly as the reference evaluates them; citations are to /root/reference.
 */
#include "orc.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y; } P2;
typedef struct { double x, y, z; } P3;
static const P2 NANP = { NAN, NAN };

/* ---- rotation (camera.cpp:49-73; calib3d Rodrigues: R = c*I + (1-c)*r*rT + s*[r]x) ---- */
static void rodrigues(const double v[3], double R[9])
{
    double theta = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    static const double I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    if (theta < 2.220446049250313e-16) { memcpy(R, I, sizeof(I)); return; }
    double c = cos(theta), s = sin(theta), c1 = 1. - c, it = theta ? 1. / theta : 0.;
    double rx = v[0] * it, ry = v[1] * it, rz = v[2] * it;
    double rrt[9] = { rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz };
    double rxm[9] = { 0, -rz, ry, rz, 0, -rx, -ry, rx, 0 };
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * rxm[k];
}
static void mat3_mul(const double* a, const double* b, double* d)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            d[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
void orc_rotation_matrix(double roll, double yaw, double pitch, double* R)
{
    double rv[3] = { roll, -yaw, -pitch }, v[3], Rx[9], Ry[9], Rz[9], t[9];
    v[0] = rv[0]; v[1] = 0; v[2] = 0; rodrigues(v, Rx);
    v[0] = 0; v[1] = rv[1]; v[2] = 0; rodrigues(v, Ry);
    v[0] = 0; v[1] = 0; v[2] = rv[2]; rodrigues(v, Rz);
    mat3_mul(Rx, Rz, t);
    mat3_mul(t, Ry, R);
}
/* core/src/lapack.cpp cv::invert, 3x3 double closed form */
static void mat3_inv(const double* S, double* t)
{
    #define Sd(i, j) S[(i) * 3 + (j)]
    double d = Sd(0,0) * (Sd(1,1) * Sd(2,2) - Sd(1,2) * Sd(2,1)) - Sd(0,1) * (Sd(1,0) * Sd(2,2) - Sd(1,2) * Sd(2,0)) +
               Sd(0,2) * (Sd(1,0) * Sd(2,1) - Sd(1,1) * Sd(2,0));
    d = 1. / d;
    t[0] = (Sd(1,1) * Sd(2,2) - Sd(1,2) * Sd(2,1)) * d;
    t[1] = (Sd(0,2) * Sd(2,1) - Sd(0,1) * Sd(2,2)) * d;
    t[2] = (Sd(0,1) * Sd(1,2) - Sd(0,2) * Sd(1,1)) * d;
    t[3] = (Sd(1,2) * Sd(2,0) - Sd(1,0) * Sd(2,2)) * d;
    t[4] = (Sd(0,0) * Sd(2,2) - Sd(0,2) * Sd(2,0)) * d;
    t[5] = (Sd(0,2) * Sd(1,0) - Sd(0,0) * Sd(1,2)) * d;
    t[6] = (Sd(1,0) * Sd(2,1) - Sd(1,1) * Sd(2,0)) * d;
    t[7] = (Sd(0,1) * Sd(2,0) - Sd(0,0) * Sd(2,1)) * d;
    t[8] = (Sd(0,0) * Sd(1,1) - Sd(0,1) * Sd(1,0)) * d;
    #undef Sd
}

/* ---- sphere helpers (camera.cpp:189-210) ---- */
static inline P2 xyz_to_lonlat(P3 q)
{
    double s = 1.0 / sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    P3 p = { q.x * s, q.y * s, q.z * s };
    P2 r = { atan2(-p.z, p.x), asin(p.y) };
    return r;
}
static inline P3 lonlat_to_xyz(P2 ll)
{
    P3 r = { cos(ll.x) * cos(ll.y), sin(ll.y), -sin(ll.x) * cos(ll.y) };
    return r;
}
/* m * r.t(): out_j = sum_k m_k * r[j][k], accumulated in k order */
static inline P3 rot_apply(const double* r, P3 m)
{
    P3 o = { m.x * r[0] + m.y * r[1] + m.z * r[2],
             m.x * r[3] + m.y * r[4] + m.z * r[5],
             m.x * r[6] + m.y * r[7] + m.z * r[8] };
    return o;
}
static inline int valid_longitude(const orc_camera* c, double lon)
{
    #define BETWEEN(x) ((x) >= c->min_lon && (x) <= c->max_lon)
    return BETWEEN(lon) || BETWEEN(lon + 2 * M_PI) || BETWEEN(lon - 2 * M_PI) ||
           BETWEEN(lon + 4 * M_PI) || BETWEEN(lon - 4 * M_PI);
    #undef BETWEEN
}

/* ---- fullframe_fisheye (cameras/fullframe_fisheye_cam.cpp) ---- */
static double cube_root(double x) { return x == 0.0 ? 0.0 : x > 0.0 ? pow(x, 1.0 / 3.0) : -pow(-x, 1.0 / 3.0); }
static void square_zero(const double* a, int* n, double* root)
{
    if (a[2] == 0.0) {
        if (a[1] == 0.0) { if (a[0] == 0.0) { *n = 1; root[0] = 0.0; } else *n = 0; }
        else { *n = 1; root[0] = -a[0] / a[1]; }
    } else if (4.0 * a[2] * a[0] > a[1] * a[1]) *n = 0;
    else {
        *n = 2;
        root[0] = (-a[1] + sqrt(a[1] * a[1] - 4.0 * a[2] * a[0])) / (2.0 * a[2]);
        root[1] = (-a[1] - sqrt(a[1] * a[1] - 4.0 * a[2] * a[0])) / (2.0 * a[2]);
    }
}
static void cube_zero(const double* a, int* n, double* root)
{
    if (a[3] == 0.0) { square_zero(a, n, root); return; }
    double p = ((-1.0 / 3.0) * (a[2] / a[3]) * (a[2] / a[3]) + a[1] / a[3]) / 3.0;
    double q = ((2.0 / 27.0) * (a[2] / a[3]) * (a[2] / a[3]) * (a[2] / a[3]) - (1.0 / 3.0) * (a[2] / a[3]) * (a[1] / a[3]) + a[0] / a[3]) / 2.0;
    if (q * q + p * p * p >= 0.0) {
        *n = 1;
        root[0] = cube_root(-q + sqrt(q * q + p * p * p)) + cube_root(-q - sqrt(q * q + p * p * p)) - a[2] / (3.0 * a[3]);
    } else {
        double phi = acos(-q / sqrt(-p * p * p));
        *n = 3;
        root[0] = 2.0 * sqrt(-p) * cos(phi / 3.0) - a[2] / (3.0 * a[3]);
        root[1] = -2.0 * sqrt(-p) * cos(phi / 3.0 + M_PI / 3.0) - a[2] / (3.0 * a[3]);
        root[2] = -2.0 * sqrt(-p) * cos(phi / 3.0 - M_PI / 3.0) - a[2] / (3.0 * a[3]);
    }
}
/* fullframe_fisheye_cam.cpp:72-103: smallest positive root of d/dr of the radial polynomial */
double orc_fisheye_correction_radius(const double* coeff)
{
    double a[4], root[3], sroot = 1000.0; int n = 0;
    for (int k = 0; k < 4; k++) { a[k] = 0.0; if (coeff[k] != 0.0) a[k] = (k + 1) * coeff[k]; }
    cube_zero(a, &n, root);
    for (int i = 0; i < n; i++) if (root[i] > 0.0 && root[i] < sroot) sroot = root[i];
    return sroot;
}
#define FF_W(c) ((c)->ip[0])
#define FF_H(c) ((c)->ip[1])
#define FF_CX(c) ((c)->ip[2])
#define FF_CY(c) ((c)->ip[3])
#define FF_CW(c) ((c)->ip[4])
#define FF_CH(c) ((c)->ip[5])
#define FF_CIRC(c) ((c)->ip[6])
#define FF_VAR(c, n) ((c)->p[3 + (n)])
static P2 ff_obj_to_image(const orc_camera* c, P2 ll)
{
    double lon = ll.x, lat = ll.y;
    double s = cos(lat) * cos(lon), v1 = sin(lat), v0 = -cos(lat) * sin(lon);
    double r = sqrt(v0 * v0 + v1 * v1);
    double theta = atan2(r, s);
    double distance = (double)FF_CW(c) / c->p[0];
    double x = -(theta * v0 / r) * distance, y = -(theta * v1 / r) * distance;
    if (fabs(ll.x) < 1e-5 && fabs(ll.y) < 1e-5) x = y = 0;
    /* do_radial_distort (:148-158) */
    double rr = sqrt(x * x + y * y) / FF_VAR(c, 4), scale;
    if (rr < FF_VAR(c, 5)) scale = ((FF_VAR(c, 3) * rr + FF_VAR(c, 2)) * rr + FF_VAR(c, 1)) * rr + FF_VAR(c, 0);
    else scale = 1000.0;
    P2 ret = { x * scale, y * scale };
    ret.x += c->p[1]; ret.y += c->p[2];
    ret.x /= (double)FF_CW(c); ret.y /= (double)FF_CH(c);
    ret.x += 0.5; ret.y += 0.5;
    if (FF_CIRC(c) && (ret.x - 0.5) * (ret.x - 0.5) + (ret.y - 0.5) * (ret.y - 0.5) > 0.25) return NANP;
    ret.x = (ret.x * FF_CW(c)) + FF_CX(c);
    ret.y = (ret.y * FF_CH(c)) + FF_CY(c);
    ret.x /= (double)FF_W(c); ret.y /= (double)FF_H(c);
    return ret;
}

/* image_to_obj_single (cameras/fullframe_fisheye_cam.cpp:223-253) with do_reverse_radial_distort (:160-184).
 * The reference finds the radius with cv::solvePoly (Durand-Kerner on the quartic V3 r^4 + V2 r^3 + V1 r^2 + V0 r - s/V4,
 * keeping the smallest positive real root, then rejecting it unless it is below the correction radius).  On [0, r_corr)
 * the polynomial P(r) = (((V3 r + V2) r + V1) r + V0) r rises monotonically from 0 (r_corr is the first positive root of
 * P'), so "smallest positive root, accepted only below r_corr" == "the root of P(r) = s/V4 in (0, r_corr) if
 * s/V4 < P(r_corr), else none".  That root is computed here by safeguarded Newton to full f64 precision; solvePoly's
 * own answer differs from it by its iteration tolerance only, hence this model as OUTPUT is pinned to the reference with
 * the contract tolerance (1e-4 px) rather than bit for bit. */
static double ff_reverse_radius(const orc_camera* c, double t)
{
    const double V0 = FF_VAR(c, 0), V1 = FF_VAR(c, 1), V2 = FF_VAR(c, 2), V3 = FF_VAR(c, 3), rc = FF_VAR(c, 5);
    const double prc = (((V3 * rc + V2) * rc + V1) * rc + V0) * rc;
    if (!(t > 0) || !(t < prc)) return -1;
    double lo = 0, hi = rc, r = t < rc ? t : 0.5 * rc;
    for (int it = 0; it < 100; it++) {
        const double f = (((V3 * r + V2) * r + V1) * r + V0) * r - t;
        if (f == 0) break;
        if (f > 0) hi = r; else lo = r;
        const double d = ((4 * V3 * r + 3 * V2) * r + 2 * V1) * r + V0;
        double rn = r - f / d;
        if (!(rn > lo && rn < hi)) rn = 0.5 * (lo + hi);
        if (rn == r) break;
        r = rn;
    }
    return r;
}
static P2 ff_image_to_obj(const orc_camera* c, P2 xy, int* ok)
{
    if (!(FF_CW(c) == FF_W(c) && FF_CH(c) == FF_H(c) && FF_CX(c) == 0 && FF_CY(c) == 0)) { *ok = 0; return NANP; }   /* CV_Assert */
    xy.x -= 0.5; xy.y -= 0.5;
    xy.x *= (double)FF_CW(c); xy.y *= (double)FF_CH(c);
    xy.x -= c->p[1]; xy.y -= c->p[2];
    if (fabs(xy.x) < 1e-5 && fabs(xy.y) < 1e-5) { P2 z = { 0, 0 }; return z; }
    {
        const double s = sqrt(xy.x * xy.x + xy.y * xy.y);
        const double r = ff_reverse_radius(c, s / FF_VAR(c, 4));
        const double scale = (r > 0 && r < FF_VAR(c, 5)) ? s / FF_VAR(c, 4) / r : 1000.0;
        xy.x = xy.x / scale; xy.y = xy.y / scale;
    }
    {
        const double distance = (double)FF_CW(c) / c->p[0];
        const double alpha = atan2(-xy.y, xy.x);
        double theta = -xy.y / distance / sin(alpha);
        if (fabs(sin(alpha)) < 1e-3) theta = -xy.x / distance / cos(alpha);
        const double lon = atan2(sin(theta) * cos(alpha), cos(theta));
        const double lat = atan(tan(alpha) * sin(lon));
        P2 r = { lon, lat };
        return r;
    }
}

/* ---- ocam (cameras/ocam_fisheye.cpp:135-244) ---- */
static P2 ocam_obj_to_image(const orc_camera* c, P2 ll)
{
    P3 q = lonlat_to_xyz(ll);
    double p3[3] = { -q.y, -q.z, -q.x }, p2[2];
    double norm = sqrt(p3[0] * p3[0] + p3[1] * p3[1]);
    double theta = atan(p3[2] / norm);
    if (norm != 0) {
        double invnorm = 1 / norm, t = theta, rho = c->invpol[0], t_i = 1;
        for (int i = 1; i < c->n_invpol; i++) { t_i *= t; rho += t_i * c->invpol[i]; }
        double x = p3[0] * invnorm * rho, y = p3[1] * invnorm * rho;
        p2[0] = x * c->p[2] + y * c->p[3] + c->p[0];
        p2[1] = x * c->p[4] + y + c->p[1];
    } else { p2[0] = c->p[0]; p2[1] = c->p[1]; }
    P2 r = { p2[1] / c->ip[0], p2[0] / c->ip[1] };
    return r;
}
static P2 ocam_image_to_obj(const orc_camera* c, P2 xy)
{
    double p2[2] = { xy.y * c->ip[1], xy.x * c->ip[0] };
    double cc = c->p[2], d = c->p[3], e = c->p[4];
    double invdet = 1 / (cc - d * e);
    double xp = invdet * ((p2[0] - c->p[0]) - d * (p2[1] - c->p[1]));
    double yp = invdet * (-e * (p2[0] - c->p[0]) + cc * (p2[1] - c->p[1]));
    double r = sqrt(xp * xp + yp * yp), zp = c->pol[0], r_i = 1;
    for (int i = 1; i < c->n_pol; i++) { r_i *= r; zp += r_i * c->pol[i]; }
    double invnorm = 1 / sqrt(xp * xp + yp * yp + zp * zp);
    P3 q = { -(invnorm * zp), -(invnorm * xp), -(invnorm * yp) };
    return xyz_to_lonlat(q);
}

/* ---- cubic (cameras/cubic.hpp) ---- */
static P2 cubic_face_to_img(int index, double x, double y)
{
    P2 r = { (index % 3) * 1.0 / 3.0, (index / 3) * 1.0 / 2.0 };
    r.x += (x + 1.0) / 2.0 / 3.0;
    r.y += (y + 1.0) / 2.0 / 2.0;
    return r;
}
static inline int within(double a, double b) { return a >= -1.0 && a <= 1.0 && b >= -1.0 && b <= 1.0; }
static P2 cubic_obj_to_image(P2 ll)
{
    P3 p = lonlat_to_xyz(ll), s;
    if (fabs(p.x) > 1e-2) {
        double f = fabs(p.x); s.x = p.x / f; s.y = p.y / f; s.z = p.z / f;
        if (within(s.y, s.z)) return s.x < 0 ? cubic_face_to_img(1, -s.z, s.y) : cubic_face_to_img(0, s.z, s.y);
    }
    if (fabs(p.z) > 1e-2) {
        double f = fabs(p.z); s.x = p.x / f; s.y = p.y / f; s.z = p.z / f;
        if (within(s.x, s.y)) return s.z < 0 ? cubic_face_to_img(4, s.x, s.y) : cubic_face_to_img(5, -s.x, s.y);
    }
    if (fabs(p.y) > 1e-2) {
        double f = fabs(p.y); s.x = p.x / f; s.y = p.y / f; s.z = p.z / f;
        if (within(s.x, s.z)) return s.y < 0 ? cubic_face_to_img(2, s.x, -s.z) : cubic_face_to_img(3, s.x, s.z);
    }
    return NANP;
}
static P2 cubic_image_to_obj(P2 xy)
{
    int ix = 0, iy = 0;
    if (xy.y >= 0.5) iy = 1;
    if (xy.x >= 2.0 / 3.0) ix = 2; else if (xy.x >= 1.0 / 3.0) ix = 1;
    double fx = (xy.x - ix * 1.0 / 3.0) * 3.0 * 2.0 - 1.0, fy = (xy.y - iy * 1.0 / 2.0) * 2.0 * 2.0 - 1.0;
    P3 q;
    switch (iy * 3 + ix) {
    case 0: q.x = 1.0; q.y = fy; q.z = fx; break;
    case 1: q.x = -1.; q.y = fy; q.z = -fx; break;
    case 2: q.x = fx; q.y = -1.; q.z = -fy; break;
    case 3: q.x = fx; q.y = 1.0; q.z = fy; break;
    case 4: q.x = fx; q.y = fy; q.z = -1.0; break;
    default: q.x = -fx; q.y = fy; q.z = 1.0; break;
    }
    return xyz_to_lonlat(q);
}

/* ---- per-model dispatch ---- */
static void normal_cam(const orc_camera* c, double* cx, double* cy, double* cz)
{
    double ar = c->p[0];
    *cx = c->p[1];
    *cz = sqrt((1.0 - *cx * *cx) / (1.0 + 1.0 / ar / ar));
    *cy = *cz / ar;
}
/* returns 0 if the model has no obj_to_image_single (pinhole/fisheye override the batch) */
static P2 obj_to_image_single(const orc_camera* c, P2 ll)
{
    switch (c->type) {
    case ORC_CAM_NORMAL: {                         /* cameras/normal.cpp:31-39 */
        double cx, cy, cz; normal_cam(c, &cx, &cy, &cz);
        P3 q = lonlat_to_xyz(ll);
        if (q.x < 0) return NANP;
        double f = q.x / cx; q.x /= f; q.y /= f; q.z /= f;
        P2 r = { (cz - q.z) / 2.0 / cz, (cy - q.y) / 2.0 / cy };
        return r; }
    case ORC_CAM_PERSPECTIVE: {                    /* cameras/perspective.cpp:28-33 */
        P3 q = lonlat_to_xyz(ll);
        double y_ = q.y * (1.0 / c->p[1] / q.x), z_ = q.z * (1.0 / c->p[1] / q.x);
        P2 r = { 0.5 - z_ / c->p[0], 0.5 - y_ };
        return r; }
    case ORC_CAM_EQUIRECT: {                       /* cameras/equirectangular.cpp:25-29 */
        P2 r = { ll.x / (M_PI * 2.0) + 0.5, (ll.y - c->p[1]) / (c->p[0] - c->p[1]) };
        return r; }
    case ORC_CAM_FULLFRAME_FISHEYE: return ff_obj_to_image(c, ll);
    case ORC_CAM_OCAM: return ocam_obj_to_image(c, ll);
    case ORC_CAM_STUPIDOVAL: {                     /* cameras/stupidoval.hpp:23-28 */
        P2 r = { cos(ll.y) * ll.x / (M_PI * 2.0) + 0.5, -ll.y / M_PI + 0.5 };
        return r; }
    case ORC_CAM_CUBIC: return cubic_obj_to_image(ll);
    case ORC_CAM_EQAREA_NORTH: {                   /* cameras/eqareanorthpole.hpp:24-33 */
        if (ll.y < c->p[0]) return NANP;
        double rho = (M_PI / 2 - ll.y) / (M_PI / 2 - c->p[0]);
        P2 r = { -rho * sin(ll.x) / 2 + 0.5, -rho * cos(ll.x) / 2 + 0.5 };
        return r; }
    case ORC_CAM_EQAREA_SOUTH: {                   /* cameras/eqareasouthpole.hpp:23-32 */
        if (ll.y > c->p[0]) return NANP;
        double rho = (ll.y + M_PI / 2) / (c->p[0] + M_PI / 2);
        P2 r = { rho * sin(ll.x) / 2 + 0.5, -rho * cos(ll.x) / 2 + 0.5 };
        return r; }
    default: return NANP;
    }
}
/* ok=0 -> NotImplemented in that direction (camera.hpp:92-103) */
static P2 image_to_obj_single(const orc_camera* c, P2 xy, int* ok)
{
    *ok = 1;
    switch (c->type) {
    case ORC_CAM_NORMAL: {                         /* cameras/normal.cpp:23-29 */
        double cx, cy, cz; normal_cam(c, &cx, &cy, &cz);
        P3 q = { cx, cy - xy.y * 2.0 * cy, cz - xy.x * 2.0 * cz };
        return xyz_to_lonlat(q); }
    case ORC_CAM_PERSPECTIVE: {                    /* cameras/perspective.cpp:21-26 */
        P3 q = { 1.0 / c->p[1], 0.5 - xy.y, (0.5 - xy.x) * c->p[0] };
        return xyz_to_lonlat(q); }
    case ORC_CAM_FULLFRAME_FISHEYE: return ff_image_to_obj(c, xy, ok);
    case ORC_CAM_EQUIRECT: {                       /* cameras/equirectangular.cpp:31-35 */
        P2 r = { (xy.x - 0.5) * M_PI * 2.0, (c->p[0] - c->p[1]) * xy.y + c->p[1] };
        return r; }
    case ORC_CAM_OCAM: return ocam_image_to_obj(c, xy);
    case ORC_CAM_STUPIDOVAL: {                     /* cameras/stupidoval.hpp:29-35 */
        double lat = (0.5 - xy.y) * M_PI, lon = (xy.x - 0.5) * M_PI * 2.0 / cos(lat);
        if (lon < -M_PI || lon > M_PI) return NANP;
        P2 r = { lon, lat };
        return r; }
    case ORC_CAM_CUBIC: return cubic_image_to_obj(xy);
    case ORC_CAM_EQAREA_NORTH: {                   /* cameras/eqareanorthpole.hpp:35-41 */
        double dx = xy.x - 0.5, dy = xy.y - 0.5;
        double rho = sqrt(dx * dx + dy * dy) * 2;
        P2 r = { atan2(-dx, -dy), M_PI / 2 - (M_PI / 2 - c->p[0]) * rho };
        return r; }
    case ORC_CAM_EQAREA_SOUTH: {                   /* cameras/eqareasouthpole.hpp:34-40 */
        double dx = xy.x - 0.5, dy = xy.y - 0.5;
        double rho = sqrt(dx * dx + dy * dy) * 2;
        P2 r = { atan2(dx, -dy), -M_PI / 2 + (c->p[0] + M_PI / 2) * rho };
        return r; }
    default: *ok = 0; return NANP; /* pinhole/fisheye: NotImplemented; fullframe_fisheye inverse needs cv::solvePoly (not restated) */
    }
}

/* pinhole / fisheye batch override (cameras/pinhole_cam.cpp:32-57, fisheye_cam.cpp:13-18;
 * calib3d/src/calibration.cpp:759-793, fisheye.cpp:120-148 with rvec = tvec = 0) */
static P2 pinhole_project(const orc_camera* c, P3 q)
{
    double k[14] = { 0 };
    for (int i = 0; i < c->n_dist && i < 14; i++) k[i] = c->dist[i];
    double fx = c->p[0], fy = c->p[1], cx = c->p[2], cy = c->p[3];
    double x = q.x, y = q.y, z = q.z;
    if (c->type == ORC_CAM_PINHOLE) {
        z = z ? 1. / z : 1; x *= z; y *= z;
        double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2, a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
        double cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
        double icdist2 = 1. / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6);
        double xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
        double yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
        P2 r = { xd * fx + cx, yd * fy + cy };
        return r;
    }
    double xx = x / z, yy = y / z;
    double r2 = xx * xx + yy * yy, r = sqrt(r2), theta = atan(r);
    double t2 = theta * theta, t3 = t2 * theta, t4 = t2 * t2, t5 = t4 * theta, t6 = t3 * t3, t7 = t6 * theta, t8 = t4 * t4, t9 = t8 * theta;
    double theta_d = theta + k[0] * t3 + k[1] * t5 + k[2] * t7 + k[3] * t9;
    double inv_r = r > 1e-8 ? 1.0 / r : 1, cdist = r > 1e-8 ? theta_d * inv_r : 1;
    double x1 = xx * cdist, y1 = yy * cdist;
    P2 o = { (x1 + 0 * y1) * fx + cx, y1 * fy + cy };
    return o;
}

double orc_camera_aspect_ratio(const orc_camera* c)
{
    switch (c->type) {
    case ORC_CAM_NORMAL: case ORC_CAM_PERSPECTIVE: return c->p[0];
    case ORC_CAM_PINHOLE: case ORC_CAM_FISHEYE: return (double)c->ip[0] / (double)c->ip[1];
    case ORC_CAM_EQUIRECT: return (2.0f * c->p[2]) / ((c->p[1] - c->p[0]) / M_PI);
    case ORC_CAM_FULLFRAME_FISHEYE: return (double)c->ip[0] / c->ip[1];
    case ORC_CAM_OCAM: return (double)c->ip[0] / c->ip[1];
    case ORC_CAM_STUPIDOVAL: return 2.0;
    case ORC_CAM_CUBIC: return 3.0 / 2.0;
    default: return 1.0;
    }
}

/* template.cpp:46-153 (one add_input) with camera.cpp:212-253 (obj_to_image),
 * :255-294 (get_include_mask), :296-315 (image_to_obj). */
int orc_template_add_input(const orc_camera* oc, const orc_camera* ic, int W, int H,
                           float* map1, float* map2, uint8_t* mask, uint8_t* visible,
                           int use_roi, int* roi, int n_prior, uint8_t* const* prior_masks,
                           const int* prior_rois)
{
    double rinv[9];
    mat3_inv(oc->rot, rinv);
    int fail = 0;
    const int batch_override = ic->type == ORC_CAM_PINHOLE || ic->type == ORC_CAM_FISHEYE;
    uint8_t* vis_tmp = ic->include_mask && !batch_override ? (uint8_t*)malloc((size_t)W * H) : NULL;

    #pragma omp parallel for schedule(static)
    for (int j = 0; j < H; j++) {
        for (int i = 0; i < W; i++) {
            size_t idx = (size_t)j * W + i;
            P2 xy = { (double)i / W, (double)j / H };
            int ok;
            P2 ll0 = image_to_obj_single(oc, xy, &ok);
            if (!ok) { fail = 1; continue; }
            P2 ll = xyz_to_lonlat(rot_apply(rinv, lonlat_to_xyz(ll0)));   /* output lon/lat (world) */
            P3 q = rot_apply(ic->rot, lonlat_to_xyz(ll));
            P2 p;
            if (batch_override) {
                if (q.z <= 0) q.x = q.y = q.z = NAN;
                P2 ip = pinhole_project(ic, q);
                p.x = ip.x / ic->ip[0]; p.y = 1.0 - ip.y / ic->ip[1];
            } else {
                P2 lli = xyz_to_lonlat(q);
                p = NANP;
                if (valid_longitude(ic, ll.x)) p = obj_to_image_single(ic, lli);
                if (p.x >= 0 && p.x < 1 && p.y >= 0 && p.y < 1 && ic->exclude_mask) {
                    int ex = (int)(p.x * ic->ex_w), ey = (int)(p.y * ic->ex_h);
                    if (ic->exclude_mask[(size_t)ey * ic->ex_w + ex]) p = NANP;
                }
                if (vis_tmp) {
                    /* get_include_mask re-projects without the longitude test (camera.cpp:275-289);
                     * note it guards on exclude_mask but indexes include_mask. */
                    P2 pv = obj_to_image_single(ic, lli);
                    uint8_t v = 0;
                    if (pv.x >= 0 && pv.x < 1 && pv.y >= 0 && pv.y < 1 && ic->exclude_mask) {
                        int ex = (int)(pv.x * ic->ex_w), ey = (int)(pv.y * ic->ex_h);
                        if (ic->include_mask[(size_t)ey * ic->in_w + ex]) v = 1;
                    }
                    vis_tmp[idx] = v;
                }
            }
            /* template.cpp:80-94: narrow to f32 FIRST, then the validity test */
            float x = (float)p.x, y = (float)p.y;
            if (isnan(x) || isnan(y) || x < 0 || x >= 1.0f || y < 0 || y >= 1.0f || visible[idx]) {
                mask[idx] = 0; map1[idx] = map2[idx] = -1.0f;
            } else {
                mask[idx] = 255; map1[idx] = x; map2[idx] = y;
            }
        }
    }
    if (fail) { free(vis_tmp); return -1; }
    /* template.cpp:104-116: newly visible points knock out earlier inputs' masks */
    if (vis_tmp) {
        for (int j = 0; j < H; j++)
            for (int i = 0; i < W; i++) {
                size_t idx = (size_t)j * W + i;
                if (!visible[idx] && vis_tmp[idx])
                    for (int k = 0; k < n_prior; k++) {
                        const int* pr = prior_rois + 4 * k;
                        if (j < pr[1] || j >= pr[1] + pr[3] || i < pr[0] || i >= pr[0] + pr[2]) continue;
                        prior_masks[k][(size_t)(j - pr[1]) * pr[2] + (i - pr[0])] = 0;
                    }
                visible[idx] = visible[idx] || vis_tmp[idx];
            }
        free(vis_tmp);
    }
    int min_h = H, max_h = 0, min_w = W, max_w = 0;
    for (int j = 0; j < H; j++)
        for (int i = 0; i < W; i++)
            if (mask[(size_t)j * W + i]) {
                if (j < min_h) min_h = j; if (j > max_h) max_h = j;
                if (i < min_w) min_w = i; if (i > max_w) max_w = i;
            }
    if (!(min_h <= max_h && min_w <= max_w)) return -2;
    min_w = min_w - 8 > 0 ? min_w - 8 : 0; min_h = min_h - 8 > 0 ? min_h - 8 : 0;
    max_w = max_w + 8 < W - 1 ? max_w + 8 : W - 1; max_h = max_h + 8 < H - 1 ? max_h + 8 : H - 1;
    roi[0] = min_w; roi[1] = min_h; roi[2] = max_w + 1 - min_w; roi[3] = max_h + 1 - min_h;
    if (!use_roi) { roi[0] = 0; roi[1] = 0; roi[2] = W; roi[3] = H; }
    return 0;
}

/* vignette.cpp:39-54 (f32 arithmetic; 1.0/(...) evaluated in double then narrowed) */
void orc_vignette_map(const float abcd[4], int width, int height, float* out)
{
    float a = abcd[0], b = abcd[1], c = abcd[2], d = abcd[3];
    for (int j = 0; j < height; j++)
        for (int i = 0; i < width; i++) {
            #define P(X) ((float)(X) * (float)(X))
            float r = sqrtf(P(i - width / 2) + P(j - height / 2)) / sqrtf(P(width / 2) + P(height / 2));
            #undef P
            out[(size_t)j * width + i] = (float)(1.0 / (a + r * r * (b + r * r * (c + d * r * r))));
        }
}
