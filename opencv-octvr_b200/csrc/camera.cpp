// csrc/camera.cpp -- JSON camera options -> CamModel (host side of map generation).
// Follows vr::Camera::Camera (modules/octvr/src/camera.cpp:49-135) and the per-model constructors
// (src/cameras/*.cpp|hpp); all set-up arithmetic is f64 like the reference.
#include "camera.h"
#include <fstream>
#include <charconv>
#include "prep.h"
#include <cmath>
#include <cstring>

namespace ob {

namespace {

// calib3d Rodrigues (vector -> matrix): R = cos(t) I + (1 - cos t) r r^T + sin(t) [r]x
void rodrigues(const double v[3], double R[9])
{
    const double theta = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    static const double I[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    if (theta < 2.220446049250313e-16) { memcpy(R, I, sizeof(I)); return; }
    const double c = std::cos(theta), s = std::sin(theta), c1 = 1. - c, it = theta ? 1. / theta : 0.;
    const double rx = v[0] * it, ry = v[1] * it, rz = v[2] * it;
    const double rrt[9] = { rx * rx, rx * ry, rx * rz, rx * ry, ry * ry, ry * rz, rx * rz, ry * rz, rz * rz };
    const double rxm[9] = { 0, -rz, ry, rz, 0, -rx, -ry, rx, 0 };
    for (int k = 0; k < 9; k++) R[k] = c * I[k] + c1 * rrt[k] + s * rxm[k];
}
void mul3(const double* a, const double* b, double* d)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) d[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
// cv::invert of a 3x3 double matrix (closed form, core/src/lapack.cpp)
void inv3(const double* S, double* t)
{
    auto s = [&](int i, int j) { return S[i * 3 + j]; };
    double d = s(0,0) * (s(1,1) * s(2,2) - s(1,2) * s(2,1)) - s(0,1) * (s(1,0) * s(2,2) - s(1,2) * s(2,0)) + s(0,2) * (s(1,0) * s(2,1) - s(1,1) * s(2,0));
    d = 1. / d;
    t[0] = (s(1,1) * s(2,2) - s(1,2) * s(2,1)) * d; t[1] = (s(0,2) * s(2,1) - s(0,1) * s(2,2)) * d; t[2] = (s(0,1) * s(1,2) - s(0,2) * s(1,1)) * d;
    t[3] = (s(1,2) * s(2,0) - s(1,0) * s(2,2)) * d; t[4] = (s(0,0) * s(2,2) - s(0,2) * s(2,0)) * d; t[5] = (s(0,2) * s(1,0) - s(0,0) * s(1,2)) * d;
    t[6] = (s(1,0) * s(2,1) - s(1,1) * s(2,0)) * d; t[7] = (s(0,1) * s(2,0) - s(0,0) * s(2,1)) * d; t[8] = (s(0,0) * s(1,1) - s(0,1) * s(1,0)) * d;
}

// --- fullframe_fisheye: radius up to which the radial polynomial is monotonic (fullframe_fisheye_cam.cpp:18-103,
//     the panotools cubic solver) ---
double cbrt_signed(double x) { return x == 0.0 ? 0.0 : x > 0.0 ? std::pow(x, 1.0 / 3.0) : -std::pow(-x, 1.0 / 3.0); }
int quadratic_roots(const double* a, double* r)
{
    if (a[2] == 0.0) {
        if (a[1] == 0.0) { if (a[0] == 0.0) { r[0] = 0.0; return 1; } return 0; }
        r[0] = -a[0] / a[1]; return 1;
    }
    if (4.0 * a[2] * a[0] > a[1] * a[1]) return 0;
    r[0] = (-a[1] + std::sqrt(a[1] * a[1] - 4.0 * a[2] * a[0])) / (2.0 * a[2]);
    r[1] = (-a[1] - std::sqrt(a[1] * a[1] - 4.0 * a[2] * a[0])) / (2.0 * a[2]);
    return 2;
}
int cubic_roots(const double* a, double* r)
{
    if (a[3] == 0.0) return quadratic_roots(a, r);
    const double p = ((-1.0 / 3.0) * (a[2] / a[3]) * (a[2] / a[3]) + a[1] / a[3]) / 3.0;
    const double q = ((2.0 / 27.0) * (a[2] / a[3]) * (a[2] / a[3]) * (a[2] / a[3]) - (1.0 / 3.0) * (a[2] / a[3]) * (a[1] / a[3]) + a[0] / a[3]) / 2.0;
    if (q * q + p * p * p >= 0.0) {
        r[0] = cbrt_signed(-q + std::sqrt(q * q + p * p * p)) + cbrt_signed(-q - std::sqrt(q * q + p * p * p)) - a[2] / (3.0 * a[3]);
        return 1;
    }
    const double phi = std::acos(-q / std::sqrt(-p * p * p));
    r[0] = 2.0 * std::sqrt(-p) * std::cos(phi / 3.0) - a[2] / (3.0 * a[3]);
    r[1] = -2.0 * std::sqrt(-p) * std::cos(phi / 3.0 + M_PI / 3.0) - a[2] / (3.0 * a[3]);
    r[2] = -2.0 * std::sqrt(-p) * std::cos(phi / 3.0 - M_PI / 3.0) - a[2] / (3.0 * a[3]);
    return 3;
}
double correction_radius(const double* coeff)
{
    double a[4], r[3], best = 1000.0;
    for (int k = 0; k < 4; k++) { a[k] = 0.0; if (coeff[k] != 0.0) a[k] = (k + 1) * coeff[k]; }
    const int n = cubic_roots(a, r);
    for (int i = 0; i < n; i++) if (r[i] > 0.0 && r[i] < best) best = r[i];
    return best;
}

CamType type_of(const std::string& t)
{
    static const char* names[] = { "normal", "perspective", "pinhole", "fisheye", "equirectangular", "fullframe_fisheye",
                                   "ocam_fisheye", "stupidoval", "cubic", "eqareanorthpole", "eqareasouthpole" };
    for (int i = 0; i < 11; i++) if (t == names[i]) return (CamType)i;
    return CAM_INVALID;
}

}  // namespace

CamHost camera_from_json(const std::string& type, const Json& o)
{
    CamHost ch;
    CamModel& m = ch.m;
    memset(&m, 0, sizeof(m));
    m.type = type_of(type);
    if (m.type == CAM_INVALID) fail(OCTVR_ERR_FORMAT, "Invalid camera type \"" + type + "\"");   // template.cpp:30,53

    double rv[3] = { 0, 0, 0 };
    if (o.has("rotation")) {
        rv[0] = o.at("rotation").at("roll").number();
        rv[1] = -o.at("rotation").at("yaw").number();
        rv[2] = -o.at("rotation").at("pitch").number();
    }
    double v[3], Rx[9], Ry[9], Rz[9], t[9];
    v[0] = rv[0]; v[1] = 0; v[2] = 0; rodrigues(v, Rx);
    v[0] = 0; v[1] = rv[1]; v[2] = 0; rodrigues(v, Ry);
    v[0] = 0; v[1] = 0; v[2] = rv[2]; rodrigues(v, Rz);
    mul3(Rx, Rz, t);
    mul3(t, Ry, m.rot);
    if (o.has("rotation_matrix"))
        for (int k = 0; k < 9; k++) m.rot[k] = o.at("rotation_matrix").at(k).number();
    inv3(m.rot, m.rot_inv);

    // exclude / include masks in the input image's pixel grid (octvr/src/camera.cpp:72-123,146-187), drawn with the
    // reference's cv::fillPoly (prep.cpp fill_poly_u8)
    auto prepare = [&](std::vector<uint8_t>& mk, uint8_t initial, int& mw, int& mh) {
        const int w = o.at("width").integer(), h = o.at("height").integer();
        OB_CHECK(w > 0 && h > 0, "mask: width / height");
        if (mk.empty()) { mk.assign((size_t)w * h, initial); mw = w; mh = h; }
        else OB_CHECK(mw == w && mh == h, "mask size != width x height");
    };
    auto draw = [&](const Json& masks, std::vector<uint8_t>& target, int mw, int mh) {
        for (size_t k = 0; k < masks.size(); k++) {
            const Json& area = masks.at(k);
            const std::string kind = area.at("type").string();
            if (kind == "polygonal") {
                const Json& a = area.at("args");
                std::vector<int> pts;
                for (size_t q = 0; q + 1 < a.size(); q += 2) { pts.push_back((int)a.at(q).number()); pts.push_back((int)a.at(q + 1).number()); }
                fill_poly_u8(target.data(), mw, mh, pts.data(), (int)pts.size() / 2, 255);
            } else if (kind == "png") {
                // camera.cpp:169-187: cv::imdecode(args, 1); red != 0 -> excluded, green != 0 -> included, whichever list the entry is in
                const Json& a = area.at("args");
                std::vector<uint8_t> file(a.size());
                for (size_t q = 0; q < a.size(); q++) file[q] = (uint8_t)a.at(q).integer();
                int pw = 0, ph = 0;
                std::vector<uint8_t> bgr;
                png_decode_bgr(file.data(), file.size(), pw, ph, bgr);
                OB_CHECK(!ch.exclude.empty() && pw == m.ex_w && ph == m.ex_h, "png mask: size != the camera's exclude mask (camera.cpp:175; list it under exclude_masks)");
                OB_CHECK(!ch.include.empty() && m.in_w == pw && m.in_h == ph, "png mask: no include mask of that size");
                for (size_t q = 0; q < (size_t)pw * ph; q++) {
                    if (bgr[3 * q + 2]) ch.exclude[q] = 255;
                    if (bgr[3 * q + 1]) ch.include[q] = 255;
                }
            } else
                fail(OCTVR_ERR_FORMAT, "unknown mask type \"" + kind + "\"");
        }
    };
    if (o.has("selection")) {                                          // camera.cpp:96-112: exclude everything but the rectangle
        prepare(ch.exclude, 255, m.ex_w, m.ex_h);
        const Json& s = o.at("selection");
        const int left = s.at(0).integer(), right = s.at(1).integer(), top = s.at(2).integer(), bottom = s.at(3).integer();
        const int pts[8] = { left, top, left, bottom - 1, right - 1, bottom - 1, right - 1, top };
        fill_poly_u8(ch.exclude.data(), m.ex_w, m.ex_h, pts, 4, 0);
    }
    if (o.has("exclude_masks")) {                                      // camera.cpp:114-118 (PTGui)
        prepare(ch.exclude, 0, m.ex_w, m.ex_h);
        prepare(ch.include, 0, m.in_w, m.in_h);
        draw(o.at("exclude_masks"), ch.exclude, m.ex_w, m.ex_h);
    }
    if (o.has("include_masks")) {                                      // camera.cpp:120-123 (Hugin)
        prepare(ch.include, 0, m.in_w, m.in_h);
        draw(o.at("include_masks"), ch.include, m.in_w, m.in_h);
    }
    if (o.has("longitude_selection")) {
        m.min_lon = o.at("longitude_selection").at(0).number();
        m.max_lon = o.at("longitude_selection").at(1).number();
        OB_CHECK(m.max_lon > m.min_lon, "longitude_selection");
    } else { m.min_lon = -M_PI; m.max_lon = M_PI; }

    switch (m.type) {
    case CAM_NORMAL:          // cameras/normal.cpp:13-21 ; cam_y, cam_z derived on the device side from p[0], p[1]
        m.p[0] = o.at("aspect_ratio").number(); m.p[1] = o.at("cam_opt").number();
        m.p[3] = std::sqrt((1.0 - m.p[1] * m.p[1]) / (1.0 + 1.0 / m.p[0] / m.p[0]));   // cam_z
        m.p[2] = m.p[3] / m.p[0];                                                        // cam_y
        break;
    case CAM_PERSPECTIVE:     // cameras/perspective.cpp:13-19
        m.p[0] = o.at("aspect_ratio").number(); m.p[1] = o.at("sf").number();
        break;
    case CAM_PINHOLE: case CAM_FISHEYE: {   // cameras/pinhole_cam.cpp:13-30
        m.p[0] = o.at("fx").number(); m.p[1] = o.at("fy").number(); m.p[2] = o.at("cx").number(); m.p[3] = o.at("cy").number();
        m.ip[0] = o.at("width").integer(); m.ip[1] = o.at("height").integer();
        const Json& d = o.at("dist_coeffs");
        m.n_dist = (int)d.size();
        OB_CHECK(m.n_dist <= 12, "dist_coeffs: at most 12 coefficients (no tilt model)");
        if (m.type == CAM_FISHEYE) OB_CHECK(m.n_dist == 4, "fisheye needs 4 dist_coeffs (fisheye.cpp:83)");
        for (int i = 0; i < m.n_dist; i++) m.dist[i] = d.at(i).number();
        break; }
    case CAM_EQUIRECT:        // cameras/equirectangular.cpp:13-23
        m.p[0] = o.has("min_lat") ? o.at("min_lat").number() : -M_PI / 2;
        m.p[1] = o.has("max_lat") ? o.at("max_lat").number() : M_PI / 2;
        m.p[2] = o.has("scale_lon") ? o.at("scale_lon").number() : 1.0;
        break;
    case CAM_FULLFRAME_FISHEYE: {   // cameras/fullframe_fisheye_cam.cpp:105-140
        const int w = o.at("width").integer(), h = o.at("height").integer();
        int cx = 0, cy = 0, cw = 0, chh = 0; bool circ = false;
        if (o.has("crop")) {
            const Json& r = o.at("crop").at("rect");
            cx = r.at(0).integer(); cy = r.at(2).integer(); cw = r.at(1).integer() - r.at(0).integer(); chh = r.at(3).integer() - r.at(2).integer();
            circ = o.at("crop").at("is_circular").boolean();
        }
        if ((int64_t)cw * chh == 0) { cx = 0; cy = 0; cw = w; chh = h; circ = false; }
        m.ip[0] = w; m.ip[1] = h; m.ip[2] = cx; m.ip[3] = cy; m.ip[4] = cw; m.ip[5] = chh; m.ip[6] = circ ? 1 : 0;
        m.p[0] = o.at("hfov").number(); m.p[1] = o.at("center_dx").number(); m.p[2] = o.at("center_dy").number();
        const Json& r = o.at("radial");
        const double a = r.at(0).number(), b = r.at(1).number(), c = r.at(2).number();
        double rd[4] = { 1.0 - a - b - c, c, b, a };
        for (int k = 0; k < 4; k++) m.p[3 + k] = rd[k];
        m.p[7] = (cw < chh ? cw : chh) / 2.0;
        m.p[8] = correction_radius(rd);
        break; }
    case CAM_OCAM: {          // cameras/ocam_fisheye.cpp:82-110
        if (o.has("file")) {
            // get_ocam_model (cameras/ocam_fisheye.cpp:19-80), Scaramuzza's calib_results.txt: '#' comment lines, then n + n direct
            // coefficients, m + m inverse coefficients, centre "xc yc", affine "c d e", image "height width"
            std::ifstream f(o.at("file").string());
            if (!f) fail(OCTVR_ERR_INVALID, "ocam_fisheye: cannot open " + o.at("file").string());
            std::vector<double> v;
            std::string line;
            while (std::getline(f, line)) {
                const size_t b = line.find_first_not_of(" \t\r");
                if (b == std::string::npos || line[b] == '#') continue;
                const char* q = line.c_str() + b;
                for (;;) {
                    while (*q == ' ' || *q == '\t' || *q == '\r') q++;
                    if (!*q) break;
                    const char* e = q;
                    while (*e && *e != ' ' && *e != '\t' && *e != '\r') e++;
                    double d = 0;
                    const auto res = std::from_chars(q, e, d);
                    if (res.ec != std::errc() || res.ptr != e) fail(OCTVR_ERR_FORMAT, "ocam_fisheye: bad number in the calibration file");
                    v.push_back(d); q = e;
                }
            }
            size_t k = 0;
            auto next = [&]() { if (k >= v.size()) fail(OCTVR_ERR_FORMAT, "ocam_fisheye: calibration file too short"); return v[k++]; };
            m.n_pol = (int)next();
            OB_CHECK(m.n_pol > 0 && m.n_pol <= 64, "ocam polynomial length");
            for (int i = 0; i < m.n_pol; i++) m.pol[i] = next();
            m.n_invpol = (int)next();
            OB_CHECK(m.n_invpol > 0 && m.n_invpol <= 64, "ocam polynomial length");
            for (int i = 0; i < m.n_invpol; i++) m.invpol[i] = next();
            m.p[0] = next(); m.p[1] = next(); m.p[2] = next(); m.p[3] = next(); m.p[4] = next();
            m.ip[1] = (int)next(); m.ip[0] = (int)next();
            break;
        }
        const Json& pol = o.at("pol"); const Json& inv = o.at("invpol");
        m.n_pol = (int)pol.size(); m.n_invpol = (int)inv.size();
        OB_CHECK(m.n_pol > 0 && m.n_pol <= 64 && m.n_invpol > 0 && m.n_invpol <= 64, "ocam polynomial length");
        for (int i = 0; i < m.n_pol; i++) m.pol[i] = pol.at(i).number();
        for (int i = 0; i < m.n_invpol; i++) m.invpol[i] = inv.at(i).number();
        m.p[0] = o.at("xc").number(); m.p[1] = o.at("yc").number(); m.p[2] = o.at("c").number(); m.p[3] = o.at("d").number(); m.p[4] = o.at("e").number();
        m.ip[0] = o.at("width").integer(); m.ip[1] = o.at("height").integer();
        break; }
    case CAM_EQAREA_NORTH: m.p[0] = o.has("arctic_circle") ? o.at("arctic_circle").number() : M_PI / 3; break;
    case CAM_EQAREA_SOUTH: m.p[0] = o.has("antarctic_circle") ? o.at("antarctic_circle").number() : -M_PI / 3; break;
    default: break;
    }

    if (o.has("vignette")) {            // vignette.cpp:18-37 (f32 members; ev is a float)
        ch.has_vignette = true;
        for (int k = 0; k < 4; k++) ch.vig[k] = (float)o.at("vignette").at(k).number();
        if (o.has("exposure")) {
            const float ev = (float)std::pow(2.0, o.at("exposure").number());
            for (int k = 0; k < 4; k++) ch.vig[k] /= ev;
        }
    }
    return ch;
}

double camera_aspect_ratio(const CamModel& m)
{
    switch (m.type) {
    case CAM_NORMAL: case CAM_PERSPECTIVE: return m.p[0];
    case CAM_PINHOLE: case CAM_FISHEYE: return (double)m.ip[0] / (double)m.ip[1];
    case CAM_EQUIRECT: return (2.0f * m.p[2]) / ((m.p[1] - m.p[0]) / M_PI);      // equirectangular.hpp:32-34
    case CAM_FULLFRAME_FISHEYE: case CAM_OCAM: return (double)m.ip[0] / m.ip[1];
    case CAM_STUPIDOVAL: return 2.0;
    case CAM_CUBIC: return 3.0 / 2.0;
    default: return 1.0;
    }
}

Img<float> vignette_map(const float abcd[4], int width, int height)
{
    Img<float> out(width, height);
    const float a = abcd[0], b = abcd[1], c = abcd[2], d = abcd[3];
    auto sq = [](int v) { return float(v) * float(v); };
    for (int j = 0; j < height; j++)
        for (int i = 0; i < width; i++) {
            const float r = std::sqrt(sq(i - width / 2) + sq(j - height / 2)) / std::sqrt(sq(width / 2) + sq(height / 2));
            out.row(j)[i] = (float)(1.0 / (a + r * r * (b + r * r * (c + d * r * r))));
        }
    return out;
}

}  // namespace ob
