// csrc/template.cpp -- "VRv11" template files and the template half of the C ABI.
// Format: modules/octvr/src/template.cpp:206-314 (every scalar an int64, mats = type,rows,cols,payload).
#include "template.h"
#include "prep.h"
#include <cstring>
#include <fstream>
#include <memory>

namespace ob {

static thread_local std::string g_err;
void set_last_error(const std::string& m) { g_err = m; }

namespace {
struct Reader {
    const uint8_t* p; size_t n, pos = 0;
    int64_t i64()
    {
        if (pos + 8 > n) fail(OCTVR_ERR_FORMAT, "Invalid data file (truncated)");
        int64_t v; memcpy(&v, p + pos, 8); pos += 8; return v;
    }
    void bytes(void* dst, size_t k)
    {
        if (pos + k > n) fail(OCTVR_ERR_FORMAT, "Invalid data file (truncated)");
        memcpy(dst, p + pos, k); pos += k;
    }
    template <class T> Img<T> mat(int cv_depth)
    {
        int64_t type = i64(), rows = i64(), cols = i64();
        if (rows * cols == 0) return Img<T>();
        if ((type & 7) != cv_depth || (type >> 3) != 0) fail(OCTVR_ERR_FORMAT, "Invalid data file (unexpected mat type)");
        if (rows < 0 || cols < 0 || rows > (1 << 20) || cols > (1 << 20)) fail(OCTVR_ERR_FORMAT, "Invalid data file (bad mat size)");
        Img<T> m((int)cols, (int)rows);
        bytes(m.d.data(), m.d.size() * sizeof(T));
        return m;
    }
    TInput input()
    {
        TInput in;
        int64_t v[4];
        for (auto& q : v) { q = i64(); if (q < 0 || q > (1 << 20)) fail(OCTVR_ERR_FORMAT, "Invalid data file (ROI out of range)"); }   // before narrowing
        in.roi.x = (int)v[0]; in.roi.y = (int)v[1]; in.roi.w = (int)v[2]; in.roi.h = (int)v[3];
        in.map1 = mat<float>(5); in.map2 = mat<float>(5); in.mask = mat<uint8_t>(0); in.vignette = mat<float>(5);
        return in;
    }
};
void w64(std::ofstream& f, int64_t v) { f.write(reinterpret_cast<const char*>(&v), 8); }
template <class T> void wmat(std::ofstream& f, const Img<T>& m, int cv_type)
{
    w64(f, m.empty() ? 0 : cv_type); w64(f, m.h); w64(f, m.w);
    if (!m.empty()) f.write(reinterpret_cast<const char*>(m.d.data()), (std::streamsize)(m.d.size() * sizeof(T)));
}
void winput(std::ofstream& f, const TInput& in)
{
    w64(f, in.roi.x); w64(f, in.roi.y); w64(f, in.roi.w); w64(f, in.roi.h);
    wmat(f, in.map1, 5); wmat(f, in.map2, 5); wmat(f, in.mask, 0); wmat(f, in.vignette, 5);
}
void check_input(const TInput& in, int W, int H)
{
    OB_CHECK(in.roi.w > 0 && in.roi.h > 0 && in.roi.x >= 0 && in.roi.y >= 0 && in.roi.x + in.roi.w <= W && in.roi.y + in.roi.h <= H,
             "input ROI outside the output frame");
    OB_CHECK(in.map1.w == in.roi.w && in.map1.h == in.roi.h && in.map2.w == in.roi.w && in.map2.h == in.roi.h &&
             in.mask.w == in.roi.w && in.mask.h == in.roi.h, "map/mask size != ROI size");
}
}  // namespace

octvr_template* template_from_dat(const uint8_t* bytes, size_t n)
{
    if (n < 5 || memcmp(bytes, "VRv11", 5) != 0) fail(OCTVR_ERR_FORMAT, "Invalid data file (version does not match)");
    Reader r{ bytes, n, 5 };
    std::unique_ptr<octvr_template> t(new octvr_template);
    const int64_t ow = r.i64(), oh = r.i64();
    if (ow <= 0 || oh <= 0 || ow > (1 << 20) || oh > (1 << 20)) fail(OCTVR_ERR_FORMAT, "Invalid data file (output size)");
    t->out_w = (int)ow; t->out_h = (int)oh;
    int64_t ni = r.i64();
    if (ni < 0 || ni > 4096) fail(OCTVR_ERR_FORMAT, "Invalid data file (input count)");
    for (int64_t i = 0; i < ni; i++) t->inputs.push_back(r.input());
    for (int64_t i = 0; i < ni; i++) t->seam_masks.push_back(r.mat<uint8_t>(0));
    int64_t no = r.i64();
    if (no < 0 || no > 4096) fail(OCTVR_ERR_FORMAT, "Invalid data file (overlay count)");
    for (int64_t i = 0; i < no; i++) t->overlays.push_back(r.input());
    for (auto& in : t->inputs) check_input(in, t->out_w, t->out_h);
    for (auto& in : t->overlays) check_input(in, t->out_w, t->out_h);
    // a seam mask is indexed with ROI coordinates by the multiband set-up: it must be absent or exactly ROI-sized
    for (size_t i = 0; i < t->inputs.size(); i++) {
        const Img<uint8_t>& sm = t->seam_masks[i];
        if (!sm.empty() && (sm.w != t->inputs[i].roi.w || sm.h != t->inputs[i].roi.h)) fail(OCTVR_ERR_FORMAT, "Invalid data file (seam mask size != ROI size)");
    }
    return t.release();
}

void template_ensure_seams(octvr_template& t)
{
    bool have = t.seam_masks.size() == t.inputs.size();
    for (auto& m : t.seam_masks) have = have && !m.empty();
    if (!have) t.seam_masks = distance_seam_masks(t.inputs, t.out_w, t.device);
}

void template_to_dat(octvr_template& t, const std::string& path)
{
    template_ensure_seams(t);                 // template.cpp:207-208
    std::ofstream f(path, std::ios::binary);
    if (!f) fail(OCTVR_ERR_INVALID, "cannot open " + path);
    f.write("VRv11", 5);
    w64(f, t.out_w); w64(f, t.out_h);
    w64(f, (int64_t)t.inputs.size());
    for (auto& in : t.inputs) winput(f, in);
    for (auto& m : t.seam_masks) wmat(f, m, 0);
    w64(f, (int64_t)t.overlays.size());
    for (auto& in : t.overlays) winput(f, in);
}

}  // namespace ob

using namespace ob;

extern "C" {

const char* octvr_last_error(void) { return g_err.c_str(); }
const char* octvr_version(void) { return "octvr_b200 0.1 (sm_100a)"; }

octvr_status octvr_template_load_dat(const void* bytes, size_t n, octvr_template** out)
{
    return guard([&] { OB_CHECK(bytes && out, "null argument"); *out = template_from_dat((const uint8_t*)bytes, n); });
}

octvr_status octvr_template_load_file(const char* path, octvr_template** out)
{
    return guard([&] {
        OB_CHECK(path && out, "null argument");
        std::ifstream f(path, std::ios::binary | std::ios::ate);
        if (!f) fail(OCTVR_ERR_INVALID, std::string("cannot open ") + path);
        std::vector<uint8_t> buf((size_t)f.tellg());
        f.seekg(0);
        f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size());
        *out = template_from_dat(buf.data(), buf.size());
    });
}

octvr_status octvr_template_dump_file(octvr_template* t, const char* path)
{
    return guard([&] { OB_CHECK(t && path, "null argument"); template_to_dat(*t, path); });
}

octvr_status octvr_template_build_json(const char* json, int width, int height, int use_roi,
                                       int with_seam_masks, int device, octvr_template** out)
{
    return guard([&] {
        OB_CHECK(json && out, "null argument");
        *out = template_from_json(json, width, height, use_roi != 0, with_seam_masks != 0, device);
    });
}

octvr_status octvr_template_create(const char* to_type, const char* to_opts_json, int width, int height, int device, octvr_template** out)
{
    return guard([&] {
        OB_CHECK(to_type && out, "null argument");
        *out = template_create(to_type, to_opts_json ? to_opts_json : "", width, height, device);
    });
}

octvr_status octvr_template_add_input(octvr_template* t, const char* from_type, const char* from_opts_json, int overlay, int use_roi)
{
    return guard([&] {
        OB_CHECK(t && from_type, "null argument");
        template_add_input(*t, from_type, from_opts_json ? from_opts_json : "", overlay != 0, use_roi != 0);
    });
}

octvr_status octvr_template_from_arrays(int out_w, int out_h, int n, const int* rois,
                                        const float* const* map1, const float* const* map2,
                                        const uint8_t* const* mask, const uint8_t* const* seam,
                                        const float* const* vignette, int vig_w, int vig_h,
                                        octvr_template** out)
{
    return guard([&] {
        OB_CHECK(out_w > 0 && out_h > 0 && n >= 1 && rois && map1 && map2 && mask && out, "bad template arrays");
        std::unique_ptr<octvr_template> t(new octvr_template);
        t->out_w = out_w; t->out_h = out_h;
        bool all_seams = seam != nullptr;
        for (int i = 0; i < n; i++) {
            TInput in;
            in.roi = Rect{ rois[4 * i], rois[4 * i + 1], rois[4 * i + 2], rois[4 * i + 3] };
            OB_CHECK(in.roi.w > 0 && in.roi.h > 0, "empty ROI");
            size_t area = (size_t)in.roi.w * in.roi.h;
            in.map1 = Img<float>(in.roi.w, in.roi.h); memcpy(in.map1.d.data(), map1[i], area * 4);
            in.map2 = Img<float>(in.roi.w, in.roi.h); memcpy(in.map2.d.data(), map2[i], area * 4);
            in.mask = Img<uint8_t>(in.roi.w, in.roi.h); memcpy(in.mask.d.data(), mask[i], area);
            if (vignette && vignette[i]) {
                OB_CHECK(vig_w > 0 && vig_h > 0, "vignette size");
                in.vignette = Img<float>(vig_w, vig_h); memcpy(in.vignette.d.data(), vignette[i], (size_t)vig_w * vig_h * 4);
            }
            check_input(in, out_w, out_h);
            t->inputs.push_back(std::move(in));
            all_seams = all_seams && seam[i] != nullptr;
        }
        if (all_seams)
            for (int i = 0; i < n; i++) {
                Img<uint8_t> s(t->inputs[i].roi.w, t->inputs[i].roi.h);
                memcpy(s.d.data(), seam[i], s.d.size());
                t->seam_masks.push_back(std::move(s));
            }
        *out = t.release();
    });
}

octvr_status octvr_template_add_overlay(octvr_template* t, const int* roi, const float* map1, const float* map2,
                                        const uint8_t* mask, const float* vignette, int vig_w, int vig_h)
{
    return guard([&] {
        OB_CHECK(t && roi && map1 && map2 && mask, "null argument");
        TInput in;
        in.roi = Rect{ roi[0], roi[1], roi[2], roi[3] };
        OB_CHECK(in.roi.w > 0 && in.roi.h > 0, "empty ROI");
        const size_t area = (size_t)in.roi.w * in.roi.h;
        in.map1 = Img<float>(in.roi.w, in.roi.h); memcpy(in.map1.d.data(), map1, area * 4);
        in.map2 = Img<float>(in.roi.w, in.roi.h); memcpy(in.map2.d.data(), map2, area * 4);
        in.mask = Img<uint8_t>(in.roi.w, in.roi.h); memcpy(in.mask.d.data(), mask, area);
        if (vignette) {
            OB_CHECK(vig_w > 0 && vig_h > 0, "vignette size");
            in.vignette = Img<float>(vig_w, vig_h); memcpy(in.vignette.d.data(), vignette, (size_t)vig_w * vig_h * 4);
        }
        check_input(in, t->out_w, t->out_h);
        t->overlays.push_back(std::move(in));
    });
}

octvr_status octvr_debug_fill_poly(uint8_t* img, int w, int h, const int* pts, int npts, int val)
{
    return guard([&] {
        OB_CHECK(img && pts && w > 0 && h > 0 && npts >= 0, "bad argument");
        fill_poly_u8(img, w, h, pts, npts, (uint8_t)val);
    });
}

octvr_status octvr_template_create_masks(octvr_template* t)
{
    return guard([&] { OB_CHECK(t, "null argument"); t->seam_masks = distance_seam_masks(t->inputs, t->out_w, t->device); });
}

int octvr_debug_seam_backend(void) { return seam_backend(); }

octvr_status octvr_template_out_size(const octvr_template* t, int* w, int* h)
{
    return guard([&] { OB_CHECK(t && w && h, "null argument"); *w = t->out_w; *h = t->out_h; });
}
int octvr_template_num_inputs(const octvr_template* t) { return t ? (int)t->inputs.size() : 0; }
int octvr_template_num_overlays(const octvr_template* t) { return t ? (int)t->overlays.size() : 0; }

octvr_status octvr_template_input(const octvr_template* t, int index, int roi[4], const float** map1,
                                  const float** map2, const uint8_t** mask, const uint8_t** seam,
                                  const float** vignette, int vig_wh[2])
{
    return guard([&] {
        OB_CHECK(t, "null argument");
        int ni = (int)t->inputs.size(), no = (int)t->overlays.size();
        OB_CHECK(index >= 0 && index < ni + no, "input index out of range");
        const TInput& in = index < ni ? t->inputs[index] : t->overlays[index - ni];
        if (roi) { roi[0] = in.roi.x; roi[1] = in.roi.y; roi[2] = in.roi.w; roi[3] = in.roi.h; }
        if (map1) *map1 = in.map1.d.data();
        if (map2) *map2 = in.map2.d.data();
        if (mask) *mask = in.mask.d.data();
        if (seam) *seam = (index < ni && index < (int)t->seam_masks.size() && !t->seam_masks[index].empty()) ? t->seam_masks[index].d.data() : nullptr;
        if (vignette) *vignette = in.vignette.empty() ? nullptr : in.vignette.d.data();
        if (vig_wh) { vig_wh[0] = in.vignette.w; vig_wh[1] = in.vignette.h; }
    });
}

void octvr_template_destroy(octvr_template* t) { delete t; }

}  // extern "C"
