// csrc/mapgen.cu -- template / map generation on the GPU (north_star subsystem 1).
//
// Replaces the CPU loops of MapperTemplate::add_input (modules/octvr/src/template.cpp:46-153) and
// Camera::{image_to_obj, obj_to_image} (src/camera.cpp:189-315) with one kernel: a thread owns one output
// pixel, evaluates the output model's inverse projection, both rotations and the input model's forward
// projection in f64 (same operation order as the reference), narrows to f32 BEFORE the [0,1) validity test
// (template.cpp:80-94) and writes map1/map2/mask; the ROI bounding box comes from warp-reduced atomics.
#include "camera.h"
#include "prep.h"
#include "template.h"
#include <cmath>
#include <memory>

namespace ob {

struct D2 { double x, y; };
struct D3 { double x, y, z; };

__device__ __forceinline__ D2 nan2() { const double n = __longlong_as_double(0x7ff8000000000000LL); return D2{ n, n }; }

// camera.cpp:189-200
__device__ __forceinline__ D2 xyz_to_lonlat(D3 q)
{
    const double s = 1.0 / sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    const double px = q.x * s, py = q.y * s, pz = q.z * s;
    return D2{ atan2(-pz, px), asin(py) };
}
__device__ __forceinline__ D3 lonlat_to_xyz(D2 ll)
{
    return D3{ cos(ll.x) * cos(ll.y), sin(ll.y), -sin(ll.x) * cos(ll.y) };
}
// camera.cpp:202-210 : row vector times R^T, accumulated in k order, no FMA contraction
__device__ __forceinline__ D3 rotate(const double* r, D3 m)
{
    D3 o;
    o.x = __dadd_rn(__dadd_rn(__dmul_rn(m.x, r[0]), __dmul_rn(m.y, r[1])), __dmul_rn(m.z, r[2]));
    o.y = __dadd_rn(__dadd_rn(__dmul_rn(m.x, r[3]), __dmul_rn(m.y, r[4])), __dmul_rn(m.z, r[5]));
    o.z = __dadd_rn(__dadd_rn(__dmul_rn(m.x, r[6]), __dmul_rn(m.y, r[7])), __dmul_rn(m.z, r[8]));
    return o;
}
__device__ __forceinline__ bool valid_longitude(const CamModel& c, double lon)
{
    const double PI = 3.14159265358979323846;
    #define OB_BETWEEN(x) ((x) >= c.min_lon && (x) <= c.max_lon)
    return OB_BETWEEN(lon) || OB_BETWEEN(lon + 2 * PI) || OB_BETWEEN(lon - 2 * PI) || OB_BETWEEN(lon + 4 * PI) || OB_BETWEEN(lon - 4 * PI);
    #undef OB_BETWEEN
}

// ---- per-model forward projections (lon/lat -> normalised image coordinates) ----
#define PI_D 3.14159265358979323846

// cameras/fullframe_fisheye_cam.cpp:148-158,188-221
__device__ D2 fullframe_fwd(const CamModel& c, D2 ll)
{
    const double lon = ll.x, lat = ll.y;
    const double s = cos(lat) * cos(lon), v1 = sin(lat), v0 = -cos(lat) * sin(lon);
    const double r = sqrt(v0 * v0 + v1 * v1);
    const double theta = atan2(r, s);
    const double cw = (double)c.ip[4], chh = (double)c.ip[5];
    const double distance = cw / c.p[0];
    double x = -(theta * v0 / r) * distance, y = -(theta * v1 / r) * distance;
    if (fabs(ll.x) < 1e-5 && fabs(ll.y) < 1e-5) x = y = 0;
    const double rr = sqrt(x * x + y * y) / c.p[7];
    const double scale = rr < c.p[8] ? ((c.p[6] * rr + c.p[5]) * rr + c.p[4]) * rr + c.p[3] : 1000.0;
    double rx = x * scale, ry = y * scale;
    rx += c.p[1]; ry += c.p[2];
    rx /= cw; ry /= chh;
    rx += 0.5; ry += 0.5;
    if (c.ip[6] && (rx - 0.5) * (rx - 0.5) + (ry - 0.5) * (ry - 0.5) > 0.25) return nan2();
    rx = (rx * c.ip[4]) + c.ip[2];
    ry = (ry * c.ip[5]) + c.ip[3];
    rx /= (double)c.ip[0]; ry /= (double)c.ip[1];
    return D2{ rx, ry };
}

// image_to_obj_single (cameras/fullframe_fisheye_cam.cpp:223-253) with do_reverse_radial_distort (:160-184).  The reference
// takes the smallest positive real root of the quartic from cv::solvePoly and accepts it only below the correction radius;
// on [0, r_corr) the polynomial P(r) = (((V3 r + V2) r + V1) r + V0) r rises monotonically from 0, so that root is the
// solution of P(r) = s / V4 in (0, r_corr) when s / V4 < P(r_corr), and there is none otherwise.  Safeguarded Newton in
// f64 (same steps as oracle/orc_camera.c ff_reverse_radius).  The crop must be the full image (CV_Assert, checked on the host).
__device__ D2 fullframe_inv(const CamModel& c, D2 xy)
{
    xy.x -= 0.5; xy.y -= 0.5;
    xy.x *= (double)c.ip[4]; xy.y *= (double)c.ip[5];
    xy.x -= c.p[1]; xy.y -= c.p[2];
    if (fabs(xy.x) < 1e-5 && fabs(xy.y) < 1e-5) return D2{ 0, 0 };
    const double V0 = c.p[3], V1 = c.p[4], V2 = c.p[5], V3 = c.p[6], rc = c.p[8];
    const double s = sqrt(xy.x * xy.x + xy.y * xy.y), t = s / c.p[7];
    double r = -1;
    const double prc = (((V3 * rc + V2) * rc + V1) * rc + V0) * rc;
    if (t > 0 && t < prc) {
        double lo = 0, hi = rc;
        r = t < rc ? t : 0.5 * rc;
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
    }
    const double scale = (r > 0 && r < rc) ? s / c.p[7] / r : 1000.0;
    xy.x = xy.x / scale; xy.y = xy.y / scale;
    const double distance = (double)c.ip[4] / c.p[0];
    const double alpha = atan2(-xy.y, xy.x);
    double theta = -xy.y / distance / sin(alpha);
    if (fabs(sin(alpha)) < 1e-3) theta = -xy.x / distance / cos(alpha);
    const double lon = atan2(sin(theta) * cos(alpha), cos(theta));
    const double lat = atan(tan(alpha) * sin(lon));
    return D2{ lon, lat };
}

// cameras/ocam_fisheye.cpp:183-244 (world2cam) and :135-166 (cam2world)
__device__ D2 ocam_fwd(const CamModel& c, D2 ll)
{
    const D3 q = lonlat_to_xyz(ll);
    const double p0 = -q.y, p1 = -q.z, p2 = -q.x;
    const double norm = sqrt(p0 * p0 + p1 * p1);
    const double theta = atan(p2 / norm);
    double u, v;
    if (norm != 0) {
        const double invnorm = 1 / norm;
        double rho = c.invpol[0], t_i = 1;
        for (int i = 1; i < c.n_invpol; i++) { t_i *= theta; rho += t_i * c.invpol[i]; }
        const double x = p0 * invnorm * rho, y = p1 * invnorm * rho;
        u = x * c.p[2] + y * c.p[3] + c.p[0];
        v = x * c.p[4] + y + c.p[1];
    } else { u = c.p[0]; v = c.p[1]; }
    return D2{ v / c.ip[0], u / c.ip[1] };
}
__device__ D2 ocam_inv(const CamModel& c, D2 xy)
{
    const double a0 = xy.y * c.ip[1], a1 = xy.x * c.ip[0];
    const double cc = c.p[2], d = c.p[3], e = c.p[4];
    const double invdet = 1 / (cc - d * e);
    const double xp = invdet * ((a0 - c.p[0]) - d * (a1 - c.p[1]));
    const double yp = invdet * (-e * (a0 - c.p[0]) + cc * (a1 - c.p[1]));
    const double r = sqrt(xp * xp + yp * yp);
    double zp = c.pol[0], r_i = 1;
    for (int i = 1; i < c.n_pol; i++) { r_i *= r; zp += r_i * c.pol[i]; }
    const double invnorm = 1 / sqrt(xp * xp + yp * yp + zp * zp);
    return xyz_to_lonlat(D3{ -(invnorm * zp), -(invnorm * xp), -(invnorm * yp) });
}

// cameras/cubic.hpp
__device__ __forceinline__ D2 cubic_face(int index, double x, double y)
{
    D2 r{ (index % 3) * 1.0 / 3.0, (index / 3) * 1.0 / 2.0 };
    r.x += (x + 1.0) / 2.0 / 3.0;
    r.y += (y + 1.0) / 2.0 / 2.0;
    return r;
}
__device__ __forceinline__ bool in_face(double a, double b) { return a >= -1.0 && a <= 1.0 && b >= -1.0 && b <= 1.0; }
__device__ D2 cubic_fwd(D2 ll)
{
    const D3 p = lonlat_to_xyz(ll);
    if (fabs(p.x) > 1e-2) {
        const double f = fabs(p.x), sx = p.x / f, sy = p.y / f, sz = p.z / f;
        if (in_face(sy, sz)) return sx < 0 ? cubic_face(1, -sz, sy) : cubic_face(0, sz, sy);
    }
    if (fabs(p.z) > 1e-2) {
        const double f = fabs(p.z), sx = p.x / f, sy = p.y / f, sz = p.z / f;
        if (in_face(sx, sy)) return sz < 0 ? cubic_face(4, sx, sy) : cubic_face(5, -sx, sy);
    }
    if (fabs(p.y) > 1e-2) {
        const double f = fabs(p.y), sx = p.x / f, sy = p.y / f, sz = p.z / f;
        if (in_face(sx, sz)) return sy < 0 ? cubic_face(2, sx, -sz) : cubic_face(3, sx, sz);
    }
    return nan2();
}
__device__ D2 cubic_inv(D2 xy)
{
    int ix = 0, iy = 0;
    if (xy.y >= 0.5) iy = 1;
    if (xy.x >= 2.0 / 3.0) ix = 2; else if (xy.x >= 1.0 / 3.0) ix = 1;
    const double fx = (xy.x - ix * 1.0 / 3.0) * 3.0 * 2.0 - 1.0, fy = (xy.y - iy * 1.0 / 2.0) * 2.0 * 2.0 - 1.0;
    D3 q;
    switch (iy * 3 + ix) {
    case 0: q = D3{ 1.0, fy, fx }; break;
    case 1: q = D3{ -1., fy, -fx }; break;
    case 2: q = D3{ fx, -1., -fy }; break;
    case 3: q = D3{ fx, 1.0, fy }; break;
    case 4: q = D3{ fx, fy, -1.0 }; break;
    default: q = D3{ -fx, fy, 1.0 }; break;
    }
    return xyz_to_lonlat(q);
}

__device__ D2 model_fwd(const CamModel& c, D2 ll)
{
    switch (c.type) {
    case CAM_NORMAL: {                         // cameras/normal.cpp:31-39 ; p[1..3] = cam_x, cam_y, cam_z
        D3 q = lonlat_to_xyz(ll);
        if (q.x < 0) return nan2();
        const double f = q.x / c.p[1];
        q.y /= f; q.z /= f;
        return D2{ (c.p[3] - q.z) / 2.0 / c.p[3], (c.p[2] - q.y) / 2.0 / c.p[2] }; }
    case CAM_PERSPECTIVE: {                    // cameras/perspective.cpp:28-33 (no behind-camera test)
        const D3 q = lonlat_to_xyz(ll);
        const double y_ = q.y * (1.0 / c.p[1] / q.x), z_ = q.z * (1.0 / c.p[1] / q.x);
        return D2{ 0.5 - z_ / c.p[0], 0.5 - y_ }; }
    case CAM_EQUIRECT:                         // cameras/equirectangular.cpp:25-29
        return D2{ ll.x / (PI_D * 2.0) + 0.5, (ll.y - c.p[1]) / (c.p[0] - c.p[1]) };
    case CAM_FULLFRAME_FISHEYE: return fullframe_fwd(c, ll);
    case CAM_OCAM: return ocam_fwd(c, ll);
    case CAM_STUPIDOVAL:                       // cameras/stupidoval.hpp:23-28
        return D2{ cos(ll.y) * ll.x / (PI_D * 2.0) + 0.5, -ll.y / PI_D + 0.5 };
    case CAM_CUBIC: return cubic_fwd(ll);
    case CAM_EQAREA_NORTH: {                   // cameras/eqareanorthpole.hpp:24-33
        if (ll.y < c.p[0]) return nan2();
        const double rho = (PI_D / 2 - ll.y) / (PI_D / 2 - c.p[0]);
        return D2{ -rho * sin(ll.x) / 2 + 0.5, -rho * cos(ll.x) / 2 + 0.5 }; }
    case CAM_EQAREA_SOUTH: {                   // cameras/eqareasouthpole.hpp:23-32
        if (ll.y > c.p[0]) return nan2();
        const double rho = (ll.y + PI_D / 2) / (c.p[0] + PI_D / 2);
        return D2{ rho * sin(ll.x) / 2 + 0.5, -rho * cos(ll.x) / 2 + 0.5 }; }
    default: return nan2();
    }
}

// image (normalised) -> lon/lat for OUTPUT models
__device__ D2 model_inv(const CamModel& c, D2 xy)
{
    switch (c.type) {
    case CAM_NORMAL:                           // cameras/normal.cpp:23-29
        return xyz_to_lonlat(D3{ c.p[1], c.p[2] - xy.y * 2.0 * c.p[2], c.p[3] - xy.x * 2.0 * c.p[3] });
    case CAM_PERSPECTIVE:                      // cameras/perspective.cpp:21-26
        return xyz_to_lonlat(D3{ 1.0 / c.p[1], 0.5 - xy.y, (0.5 - xy.x) * c.p[0] });
    case CAM_EQUIRECT:                         // cameras/equirectangular.cpp:31-35
        return D2{ (xy.x - 0.5) * PI_D * 2.0, (c.p[0] - c.p[1]) * xy.y + c.p[1] };
    case CAM_OCAM: return ocam_inv(c, xy);
    case CAM_FULLFRAME_FISHEYE: return fullframe_inv(c, xy);
    case CAM_STUPIDOVAL: {                     // cameras/stupidoval.hpp:29-35
        const double lat = (0.5 - xy.y) * PI_D, lon = (xy.x - 0.5) * PI_D * 2.0 / cos(lat);
        if (lon < -PI_D || lon > PI_D) return nan2();
        return D2{ lon, lat }; }
    case CAM_CUBIC: return cubic_inv(xy);
    case CAM_EQAREA_NORTH: {                   // cameras/eqareanorthpole.hpp:35-41
        const double dx = xy.x - 0.5, dy = xy.y - 0.5, rho = sqrt(dx * dx + dy * dy) * 2;
        return D2{ atan2(-dx, -dy), PI_D / 2 - (PI_D / 2 - c.p[0]) * rho }; }
    case CAM_EQAREA_SOUTH: {                   // cameras/eqareasouthpole.hpp:34-40
        const double dx = xy.x - 0.5, dy = xy.y - 0.5, rho = sqrt(dx * dx + dy * dy) * 2;
        return D2{ atan2(dx, -dy), -PI_D / 2 + (c.p[0] + PI_D / 2) * rho }; }
    default: return nan2();
    }
}

// cv::projectPoints / cv::fisheye::projectPoints with rvec = tvec = 0 (calib3d/src/calibration.cpp:759-793,
// fisheye.cpp:120-148), as called from cameras/pinhole_cam.cpp:52-57 and fisheye_cam.cpp:13-18
__device__ D2 pinhole_project(const CamModel& c, D3 q)
{
    double k[12];
    #pragma unroll
    for (int i = 0; i < 12; i++) k[i] = i < c.n_dist ? c.dist[i] : 0.0;
    const double fx = c.p[0], fy = c.p[1], cx = c.p[2], cy = c.p[3];
    double x = q.x, y = q.y, z = q.z;
    if (c.type == CAM_PINHOLE) {
        z = z ? 1. / z : 1; x *= z; y *= z;
        const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2, a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
        const double cdist = 1 + k[0] * r2 + k[1] * r4 + k[4] * r6;
        const double icdist2 = 1. / (1 + k[5] * r2 + k[6] * r4 + k[7] * r6);
        const double xd = x * cdist * icdist2 + k[2] * a1 + k[3] * a2 + k[8] * r2 + k[9] * r4;
        const double yd = y * cdist * icdist2 + k[2] * a3 + k[3] * a1 + k[10] * r2 + k[11] * r4;
        return D2{ xd * fx + cx, yd * fy + cy };
    }
    const double xx = x / z, yy = y / z;
    const double r2 = xx * xx + yy * yy, r = sqrt(r2), theta = atan(r);
    const double t2 = theta * theta, t3 = t2 * theta, t4 = t2 * t2, t5 = t4 * theta, t6 = t3 * t3, t7 = t6 * theta, t8 = t4 * t4, t9 = t8 * theta;
    const double theta_d = theta + k[0] * t3 + k[1] * t5 + k[2] * t7 + k[3] * t9;
    const double inv_r = r > 1e-8 ? 1.0 / r : 1, cdist = r > 1e-8 ? theta_d * inv_r : 1;
    const double x1 = xx * cdist, y1 = yy * cdist;
    return D2{ (x1 + 0 * y1) * fx + cx, y1 * fy + cy };
}

struct MapgenParams {
    CamModel out, in;
    int W, H;
    float* map1; float* map2; uint8_t* mask;
    const uint8_t* visible;      // W*H or null: pixels already claimed by an earlier include mask (template.cpp:86)
    uint8_t* vis;                // W*H or null: Camera::get_include_mask of this input (camera.cpp:255-293)
    int* bbox;                   // min_w, min_h, max_w, max_h
};

__global__ void __launch_bounds__(256) k_mapgen(const __grid_constant__ MapgenParams p)
{
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    bool on = false;
    if (i < p.W && j < p.H) {
        const size_t idx = (size_t)j * p.W + i;
        const D2 xy{ (double)i / p.W, (double)j / p.H };
        // Camera::image_to_obj of the OUTPUT model (camera.cpp:296-315)
        const D2 ll0 = model_inv(p.out, xy);
        const D2 ll = xyz_to_lonlat(rotate(p.out.rot_inv, lonlat_to_xyz(ll0)));
        // Camera::obj_to_image of the INPUT model (camera.cpp:212-253)
        D3 q = rotate(p.in.rot, lonlat_to_xyz(ll));
        D2 pt;
        if (p.in.type == CAM_PINHOLE || p.in.type == CAM_FISHEYE) {     // batch override, +z forward, y flipped
            if (q.z <= 0) { const D2 n = nan2(); q = D3{ n.x, n.x, n.x }; }
            const D2 ip = pinhole_project(p.in, q);
            pt = D2{ ip.x / p.in.ip[0], 1.0 - ip.y / p.in.ip[1] };
        } else {
            const D2 lli = xyz_to_lonlat(q);
            pt = nan2();
            if (valid_longitude(p.in, ll.x)) pt = model_fwd(p.in, lli);
            if (pt.x >= 0 && pt.x < 1 && pt.y >= 0 && pt.y < 1 && p.in.exclude_mask) {
                const int ex = (int)(pt.x * p.in.ex_w), ey = (int)(pt.y * p.in.ex_h);
                if (p.in.exclude_mask[(size_t)ey * p.in.ex_w + ex]) pt = nan2();
            }
            if (p.vis) {
                // get_include_mask re-projects WITHOUT the longitude test, guards on the exclude mask being present and
                // indexes the include mask with the exclude mask's size (camera.cpp:275-289) -- reproduced as is
                const D2 pv = model_fwd(p.in, lli);
                uint8_t v = 0;
                if (pv.x >= 0 && pv.x < 1 && pv.y >= 0 && pv.y < 1 && p.in.exclude_mask) {
                    const int ex = (int)(pv.x * p.in.ex_w), ey = (int)(pv.y * p.in.ex_h);
                    if (p.in.include_mask[(size_t)ey * p.in.in_w + ex]) v = 1;
                }
                p.vis[idx] = v;
            }
        }
        const float x = (float)pt.x, y = (float)pt.y;           // narrow FIRST (template.cpp:82-83)
        const bool bad = isnan(x) || isnan(y) || x < 0 || x >= 1.0f || y < 0 || y >= 1.0f || (p.visible && p.visible[idx]);
        on = !bad;
        p.mask[idx] = on ? 255 : 0;
        p.map1[idx] = on ? x : -1.0f;
        p.map2[idx] = on ? y : -1.0f;
    }
    // ROI bounding box (template.cpp:96-101): warp-level min/max, then four atomics per warp that saw a valid pixel
    const unsigned any = __ballot_sync(0xffffffffu, on);
    if (any) {
        const int BIG = 0x3fffffff;
        int mnx = on ? i : BIG, mxx = on ? i : -1, mny = on ? j : BIG, mxy = on ? j : -1;
        mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
        mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(p.bbox + 0, mnx); atomicMin(p.bbox + 1, mny);
            atomicMax(p.bbox + 2, mxx); atomicMax(p.bbox + 3, mxy);
        }
    }
}

namespace {
template <class T> struct DevBuf {
    T* p = nullptr;
    explicit DevBuf(size_t n) { OB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); }
    ~DevBuf() { cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
};
bool output_supported(int type)
{
    return type != CAM_PINHOLE && type != CAM_FISHEYE;      // these two have no image_to_obj (camera.hpp:92-103: NotImplemented)
}
}  // namespace

namespace {
const Json& empty_object()
{
    static const Json e = [] { Json j; j.kind = Json::Obj; return j; }();
    return e;
}
octvr_template* create_from(const std::string& type, const Json& opts, int width, int height, int device)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count)
        fail(OCTVR_ERR_CUDA, "no usable CUDA device (map generation has no CPU fallback)");
    OB_CUDA(cudaSetDevice(device));
    std::shared_ptr<CamHost> oc = std::make_shared<CamHost>(camera_from_json(type, opts));
    // pinhole / fisheye have no image_to_obj (throws NotImplemented, camera.hpp:92-103)
    if (!output_supported(oc->m.type)) fail(OCTVR_ERR_UNSUPPORTED, "this camera model cannot be used as the output model");
    if (oc->m.type == CAM_FULLFRAME_FISHEYE && !(oc->m.ip[4] == oc->m.ip[0] && oc->m.ip[5] == oc->m.ip[1] && oc->m.ip[2] == 0 && oc->m.ip[3] == 0))
        fail(OCTVR_ERR_INVALID, "fullframe_fisheye as the output model must not be cropped (fullframe_fisheye_cam.cpp:224)");
    // MapperTemplate::MapperTemplate (template.cpp:23-44)
    if (height <= 0 && width <= 0) fail(OCTVR_ERR_FORMAT, "Output width/height invalid");
    const double ar = camera_aspect_ratio(oc->m);
    if (height <= 0) height = int(double(width) / ar);
    if (width <= 0) width = int(double(height) * ar);
    OB_CHECK(width > 0 && height > 0 && (int64_t)width * height < ((int64_t)1 << 31), "output size");
    std::unique_ptr<octvr_template> t(new octvr_template);
    t->out_w = width; t->out_h = height; t->out_cam = oc; t->device = device;
    return t.release();
}
// MapperTemplate::add_input (template.cpp:46-153): the projection of every output pixel runs as a CUDA kernel
void add_input_from(octvr_template& tt, const std::string& type, const Json& opts, bool overlay, bool use_roi)
{
    octvr_template* t = &tt;
    if (!t->out_cam) fail(OCTVR_ERR_INVALID, "add_input needs a template made by MapperTemplate(to, to_opts, width, height)");
    const CamHost& oc = *static_cast<const CamHost*>(t->out_cam.get());
    OB_CUDA(cudaSetDevice(t->device));
    InitTrace tr("add_input");
    const int width = t->out_w, height = t->out_h;
    const size_t area = (size_t)width * height;
    DevBuf<float> d_m1(area), d_m2(area);
    DevBuf<uint8_t> d_mask(area);
    DevBuf<int> d_bbox(4);
    // MapperTemplate::visible_mask (template.cpp:41): output pixels an include mask has claimed so far
    std::vector<uint8_t>& visible = t->visible;
    std::unique_ptr<DevBuf<uint8_t>> d_visible, d_vis;
    if (!visible.empty()) {
        d_visible.reset(new DevBuf<uint8_t>(area));
        OB_CUDA(cudaMemcpy(d_visible->p, visible.data(), area, cudaMemcpyHostToDevice));
    }
    {
        tr.lap("alloc");
        CamHost ic = camera_from_json(type, opts);
        tr.lap("camera_from_json");
        std::unique_ptr<DevBuf<uint8_t>> d_ex;
        if (!ic.exclude.empty()) {
            d_ex.reset(new DevBuf<uint8_t>(ic.exclude.size()));
            OB_CUDA(cudaMemcpy(d_ex->p, ic.exclude.data(), ic.exclude.size(), cudaMemcpyHostToDevice));
            ic.m.exclude_mask = d_ex->p;
        }
        std::unique_ptr<DevBuf<uint8_t>> d_in;
        const bool batch = ic.m.type == CAM_PINHOLE || ic.m.type == CAM_FISHEYE;      // no obj_to_image_single: get_include_mask would throw
        if (!ic.include.empty()) {
            if (batch) fail(OCTVR_ERR_UNSUPPORTED, "include masks on a pinhole / fisheye input (NotImplemented in the reference, camera.hpp:92-103)");
            d_in.reset(new DevBuf<uint8_t>(ic.include.size()));
            OB_CUDA(cudaMemcpy(d_in->p, ic.include.data(), ic.include.size(), cudaMemcpyHostToDevice));
            ic.m.include_mask = d_in->p;
            d_vis.reset(new DevBuf<uint8_t>(area));
        }
        MapgenParams p;
        p.out = oc.m; p.in = ic.m; p.W = width; p.H = height;
        p.map1 = d_m1.p; p.map2 = d_m2.p; p.mask = d_mask.p; p.bbox = d_bbox.p;
        p.visible = d_visible ? d_visible->p : nullptr;
        p.vis = !ic.include.empty() ? d_vis->p : nullptr;
        const int init[4] = { 0x3fffffff, 0x3fffffff, -1, -1 };
        OB_CUDA(cudaMemcpy(d_bbox.p, init, sizeof(init), cudaMemcpyHostToDevice));
        k_mapgen<<<dim3((width + 31) / 32, (height + 7) / 8), dim3(32, 8)>>>(p);
        OB_CUDA(cudaGetLastError());
        int bb[4];
        OB_CUDA(cudaMemcpy(bb, d_bbox.p, sizeof(bb), cudaMemcpyDeviceToHost));
        tr.lap("k_mapgen");
        // CV_Assert(min_h <= max_h && min_w <= max_w), template.cpp:123
        if (bb[2] < 0 || bb[3] < 0) fail(OCTVR_ERR_INVALID, "input does not cover any output pixel (min_h <= max_h && min_w <= max_w)");
        int min_w = std::max(0, bb[0] - 8), min_h = std::max(0, bb[1] - 8);
        int max_w = std::min(width - 1, bb[2] + 8), max_h = std::min(height - 1, bb[3] + 8);
        Rect roi{ min_w, min_h, max_w + 1 - min_w, max_h + 1 - min_h };
        if (!use_roi) roi = Rect{ 0, 0, width, height };
        TInput in;
        in.roi = roi;
        in.map1 = Img<float>(roi.w, roi.h); in.map2 = Img<float>(roi.w, roi.h); in.mask = Img<uint8_t>(roi.w, roi.h);
        const size_t off = (size_t)roi.y * width + roi.x;
        OB_CUDA(cudaMemcpy2D(in.map1.d.data(), (size_t)roi.w * 4, d_m1.p + off, (size_t)width * 4, (size_t)roi.w * 4, roi.h, cudaMemcpyDeviceToHost));
        OB_CUDA(cudaMemcpy2D(in.map2.d.data(), (size_t)roi.w * 4, d_m2.p + off, (size_t)width * 4, (size_t)roi.w * 4, roi.h, cudaMemcpyDeviceToHost));
        OB_CUDA(cudaMemcpy2D(in.mask.d.data(), (size_t)roi.w, d_mask.p + off, (size_t)width, (size_t)roi.w, roi.h, cudaMemcpyDeviceToHost));
        tr.lap("copy tables to host");
        if (ic.has_vignette) in.vignette = vignette_map(ic.vig, 512, 512);          // template.cpp:18-19,135-136
        tr.lap("vignette");
        if (p.vis) {
            // template.cpp:102-116: points this input's include mask makes visible for the first time are knocked out of
            // every EARLIER input's mask (their ROIs stay as they were), then join the visible set
            std::vector<uint8_t> vis(area);
            OB_CUDA(cudaMemcpy(vis.data(), d_vis->p, area, cudaMemcpyDeviceToHost));
            if (visible.empty()) visible.assign(area, 0);
            for (int h = 0; h < height; h++)
                for (int w = 0; w < width; w++) {
                    const size_t idx = (size_t)h * width + w;
                    if (!visible[idx] && vis[idx])
                        for (auto& prior : t->inputs) {
                            const Rect& r = prior.roi;
                            if (h < r.y || h >= r.y + r.h || w < r.x || w >= r.x + r.w) continue;
                            prior.mask.row(h - r.y)[w - r.x] = 0;
                        }
                    visible[idx] = visible[idx] || vis[idx];
                }
        }
        (overlay ? t->overlays : t->inputs).push_back(std::move(in));
    }
    t->seam_masks.clear();                     // stale once the set of inputs changes; create_masks() rebuilds them
}
}  // namespace

octvr_template* template_create(const std::string& to, const std::string& to_opts_json, int width, int height, int device)
{
    const Json opts = to_opts_json.empty() ? empty_object() : JsonParser::parse(to_opts_json);
    return create_from(to, opts, width, height, device);
}

void template_add_input(octvr_template& t, const std::string& from, const std::string& from_opts_json, bool overlay, bool use_roi)
{
    const Json opts = from_opts_json.empty() ? empty_object() : JsonParser::parse(from_opts_json);
    add_input_from(t, from, opts, overlay, use_roi);
}

// the whole config at once (apps/octvr/dump.cpp:71-127): MapperTemplate(to, ...) + add_input per entry + create_masks()
octvr_template* template_from_json(const std::string& json, int width, int height, bool use_roi, bool with_seams, int device)
{
    const Json cfg = JsonParser::parse(json);
    const Json& jo = cfg.at("output");
    std::unique_ptr<octvr_template> t(create_from(jo.at("type").string(), jo.has("options") ? jo.at("options") : empty_object(), width, height, device));
    const Json& ins = cfg.at("inputs");
    OB_CHECK(ins.kind == Json::Arr && ins.size() >= 1, "config: \"inputs\" must be a non-empty array");
    for (size_t i = 0; i < ins.size(); i++) add_input_from(*t, ins.at(i).at("type").string(), ins.at(i).at("options"), false, use_roi);
    if (cfg.has("overlays"))
        for (size_t i = 0; i < cfg.at("overlays").size(); i++) add_input_from(*t, cfg.at("overlays").at(i).at("type").string(), cfg.at("overlays").at(i).at("options"), true, use_roi);
    if (with_seams) t->seam_masks = distance_seam_masks(t->inputs, t->out_w, device);      // create_masks(), template.cpp:155-204
    return t.release();
}

}  // namespace ob
